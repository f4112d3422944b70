// ws_zigzag.cu — ZigZag pivot -> feed series expansion.
//
// Pivot DETECTION is MetaQuotes' stock ZigZag indicator (iCustom, not in the reference tree); what
// the reference owns is the expansion of the indicator's buffers into the window fed to the FFT:
//   * WaveSpecZZ_1.1.0-gpuopt.mq5:393-451  ZigZagFeed::BuildFeed  — STEP / INTERP / MID
//   * Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:237-357 BuildZigZagPriceSeries
//     — ALTERNATING (== STEP) / CONTINUOUS (== INTERP, same arithmetic), other pivot rule and the
//       "fewer than two pivots -> skip the bar" guard.
// Every window restarts the expansion, but only through "the first / last pivot inside the
// window", so two per-series index arrays (previous pivot at or before a bar, next pivot at or
// after it) make every output sample an O(1) function of its bar: one thread per sample, no
// per-window rescans (the reference's INTERP mode is accidentally O(N^2) per bar, :416-443).
// Compiled with -fmad=false: `va + (vb - va) * t` is rounded as the reference rounds it, so the
// feed is bit-identical to the CPU statement.
#include "ws_common.cuh"
#include "ws_series.h"

namespace ws {

// rule 0 (1.1.0): pivot where main != 0.   rule 1 (Legacy): main, else high, else low, finite.
__device__ __forceinline__ double pivot_value(double m, double h, double l, int rule) {
    if (rule == 0) return m;
    double v = m;
    if (v == 0.0 || !isfinite(v)) {
        if (h != 0.0 && isfinite(h)) v = h;
        else if (l != 0.0 && isfinite(l)) v = l;
    }
    return (v != 0.0 && isfinite(v)) ? v : 0.0;
}

__global__ void zigzag_index_kernel(const double* __restrict__ zmain, const double* __restrict__ zhigh,
                                    const double* __restrict__ zlow, int32_t n_series, int32_t len, int rule,
                                    double* __restrict__ pv, int32_t* __restrict__ prev, int32_t* __restrict__ next) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_series) return;
    const int64_t o = (int64_t)s * len;
    int last = -1;
    for (int a = 0; a < len; a++) {
        const double v = pivot_value(zmain[o + a], zhigh[o + a], zlow[o + a], rule);
        pv[o + a] = v;
        if (v != 0.0) last = a;
        prev[o + a] = last;
    }
    int nx = len;
    for (int a = len - 1; a >= 0; a--) {
        if (pv[o + a] != 0.0) nx = a;
        next[o + a] = nx;
    }
}

__global__ void zigzag_expand_kernel(const double* __restrict__ pv, const int32_t* __restrict__ prev,
                                     const int32_t* __restrict__ next, const double* __restrict__ zhigh,
                                     const double* __restrict__ zlow, const double* __restrict__ fallback,
                                     int32_t len, int64_t nwin, int32_t N, int32_t hop, int mode, int min_pivots,
                                     double* __restrict__ lines, int32_t* __restrict__ valid) {
    const int s = blockIdx.y;
    const int64_t o = (int64_t)s * len;
    const int64_t total = nwin * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = i / N;
        const int j = (int)(i - w * N);
        const int ws = (int)(w * hop), we = ws + N - 1, a = ws + j;
        const int f = next[o + ws];                 // first pivot of the window (or beyond it)
        const int l = prev[o + we];                 // last pivot of the window (or before it)
        const bool any = f <= we;
        double v;
        if (mode == 2) {
            v = (zhigh[o + a] + zlow[o + a]) * 0.5;
        } else if (!any) {
            v = fallback[s];
        } else {
            const int pp = prev[o + a], nn = next[o + a];
            if (pp < ws) v = pv[o + f];              // before the first pivot: its value
            else if (mode == 0 || pp == a || nn > we) v = pv[o + pp];   // hold / on a pivot / after the last
            else {
                const double va = pv[o + pp], vb = pv[o + nn];
                const double t = (double)(a - pp) / (double)(nn - pp);
                v = va + (vb - va) * t;
            }
        }
        lines[((int64_t)s * nwin + w) * N + j] = v;
        if (j == 0 && valid) {
            int cnt = !any ? 0 : (l > f ? 2 : 1);
            valid[(int64_t)s * nwin + w] = cnt >= min_pivots ? 1 : 0;
        }
    }
}

cudaError_t launch_zigzag(const double* zmain, const double* zhigh, const double* zlow, const double* fallback,
                          int32_t n_series, int32_t len, int32_t N, int32_t hop, int rule, int mode, int min_pivots,
                          double* pv, int32_t* prev, int32_t* next, double* lines, int32_t* valid,
                          cudaStream_t stream) {
    const int64_t nwin = 1 + (int64_t)(len - N) / hop;
    zigzag_index_kernel<<<(n_series + 31) / 32, 32, 0, stream>>>(zmain, zhigh, zlow, n_series, len, rule, pv, prev, next);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int64_t blocks = (nwin * N + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    dim3 grid((unsigned)blocks, (unsigned)n_series);
    zigzag_expand_kernel<<<grid, 256, 0, stream>>>(pv, prev, next, zhigh, zlow, fallback, len, nwin, N, hop, mode,
                                                   min_pivots, lines, valid);
    return cudaGetLastError();
}

// ---- applied price (A1): Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3308-3316 --------------------
// One bar per thread, grid-stride; this file is compiled with -fmad=false, and the expressions
// keep the reference's operand order, so the series is bit-identical to the MQL5 loop.
__global__ void applied_price_kernel(const double* __restrict__ o, const double* __restrict__ h,
                                     const double* __restrict__ l, const double* __restrict__ c, int64_t n,
                                     int mode, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v;
        switch (mode) {
            case 1: v = c[i]; break;
            case 2: v = o[i]; break;
            case 3: v = h[i]; break;
            case 4: v = l[i]; break;
            case 5: v = (h[i] + l[i]) / 2.0; break;
            case 6: v = (h[i] + l[i] + c[i]) / 3.0; break;
            default: v = (h[i] + l[i] + 2 * c[i]) / 4.0; break;
        }
        out[i] = v;
    }
}

cudaError_t launch_applied_price(const double* o, const double* h, const double* l, const double* c, int64_t n,
                                 int mode, double* out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    applied_price_kernel<<<(unsigned)blocks, 256, 0, stream>>>(o, h, l, c, n, mode, out);
    return cudaGetLastError();
}

}  // namespace ws
