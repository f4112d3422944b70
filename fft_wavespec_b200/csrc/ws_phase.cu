// ws_phase.cu — phase / unwrapped phase / group delay of every bin of every window (A6), computed
// from a spectra plane (sm_100a).
//
// Reference: Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:1183-1263 (called per bar from
// Legacy/WaveSpecZZ_1.0.2.mq5:3104-3108): phi = atan2(im, re); numpy-style unwrap (one +-2 pi
// correction per step); central-difference -d(phi)/dk clamped to +-100.
//
// The chain only needs the window's spectrum, so it is a kernel of its own behind whichever FFT
// kernel produced the plane (the sliding kernels for plain hop-1 windows, the per-window kernels
// otherwise).  HBM bound by construction: 8 N bytes in, 12 N bytes out per window.
//
// One warp per window.  atan2 runs lane-strided over the bins; the unwrap is a prefix sum of
// increments that depend on neighbouring phases only (diff_i + corr_i), so every lane sums the
// increments of its M/32 consecutive bins, a warp scan adds the lanes' totals, and bin i gets
// phi_0 + (lanes before) + (local prefix) — the reference's serial sum up to re-association
// (a few ulp of a value of at most ~1e3 rad; the parity bar is 1e-9 relative).  The per-warp arrays
// are padded by one double per lane chunk so that both the lane-strided and the chunked accesses
// are free of bank conflicts.
#include "ws_common.cuh"
#include "ws_series.h"

namespace ws {

namespace {
constexpr double kPiPh = 3.14159265358979323846;
constexpr int kMaxPhaseWarps = 4;

__device__ __forceinline__ int pad_idx(int i, int lc) { return i + (i >> lc); }   // lc = log2(chunk)

// blockDim.x / 32 windows in flight per CTA (4 up to N = 4096, fewer above: two padded arrays of
// N/2 doubles per warp)
__global__ void __launch_bounds__(kMaxPhaseWarps * 32)
phase_chain_kernel(const double2* __restrict__ spec, int64_t spec_nwin, int64_t spec_w0, int n_series,
                   int64_t win_offset, int64_t chunk_nwin, int64_t nwin, int M, int log2M,
                   double* __restrict__ phase) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kPhaseWarps = blockDim.x >> 5;
    const int C = M >> 5;                       // bins per lane chunk (M >= 32)
    const int lc = log2M - 5;
    const int padded = M + 32;
    double* ph = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * padded;
    double* un = ph + padded;
    const int64_t total = (int64_t)n_series * chunk_nwin;
    for (int64_t item = (int64_t)blockIdx.x * kPhaseWarps + warp; item < total;
         item += (int64_t)gridDim.x * kPhaseWarps) {
        const int64_t s = item / chunk_nwin, j = item - s * chunk_nwin;
        const int64_t w = win_offset + j;
        const double2* X = spec + (s * spec_nwin + (w - spec_w0)) * M;
        // eight 16-byte loads in flight per lane before the first atan2 (its branches keep the
        // compiler from hoisting the loads of later iterations by itself)
        for (int k0 = 0; k0 < M; k0 += 256) {
            double2 xr[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int k = k0 + lane + 32 * u;
                xr[u] = k < M ? __ldcs(X + k) : make_double2(1.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int k = k0 + lane + 32 * u;
                if (k < M) ph[pad_idx(k, lc)] = atan2(xr[u].y, xr[u].x);
            }
        }
        __syncwarp();
        // local prefix of the increments of bins [lane C, lane C + C)
        const int i0 = lane * C;
        double prev = ph[pad_idx(i0 > 0 ? i0 - 1 : 0, lc)];
        double acc = 0.0;
        for (int t = 0; t < C; t++) {
            const int i = i0 + t;
            const double cur = ph[pad_idx(i, lc)];
            if (i > 0) {
                const double diff = cur - prev;
                double corr = 0.0;
                if (diff > kPiPh) corr = -2.0 * kPiPh;
                else if (diff < -kPiPh) corr = 2.0 * kPiPh;
                acc = acc + diff + corr;
            }
            un[pad_idx(i, lc)] = acc;
            prev = cur;
        }
        double tot = acc;                       // inclusive scan of the lane totals
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, tot, d);
            if (lane >= d) tot += v;
        }
        double before = __shfl_up_sync(0xffffffffu, tot, 1);
        if (lane == 0) before = 0.0;
        const double base = ph[0] + before;
        for (int t = 0; t < C; t++) {
            const int q = pad_idx(i0 + t, lc);
            un[q] = base + un[q];
        }
        __syncwarp();
        double* o = phase + (s * nwin + w) * (int64_t)(3 * M);
        for (int k = lane; k < M; k += 32) {
            const double u = un[pad_idx(k, lc)];
            double g;
            if (k == 0) g = -(un[pad_idx(1, lc)] - u);
            else if (k == M - 1) g = -(u - un[pad_idx(M - 2, lc)]);
            else g = -(un[pad_idx(k + 1, lc)] - un[pad_idx(k - 1, lc)]) / 2.0;
            if (g > 100.0) g = 100.0;
            if (g < -100.0) g = -100.0;
            __stcs(o + k, ph[pad_idx(k, lc)]);
            __stcs(o + M + k, u);
            __stcs(o + 2 * M + k, g);
        }
        __syncwarp();
    }
}
}  // namespace

bool phase_from_spectra_supported(int N) { return N >= 64 && N <= 8192; }      // M >= 32: one chunk per lane

cudaError_t launch_phase_from_spectra(const double* spectra, int64_t spec_nwin, int64_t spec_w0, int n_series,
                                      int64_t win_offset, int64_t chunk_nwin, int64_t nwin, int N,
                                      double* phase, cudaStream_t stream) {
    const int M = N / 2;
    int log2M = 0;
    while ((1 << log2M) < M) log2M++;
    const size_t per_warp = (size_t)2 * (M + 32) * 8;
    int warps = kMaxPhaseWarps;
    while (warps > 1 && warps * per_warp > 113 * 1024) warps >>= 1;      // two CTAs per SM where possible
    const size_t smem = warps * per_warp;
    static std::atomic<unsigned long long> attr_seen{0};
    if (first_launch_on_device(attr_seen)) {
        cudaError_t e = cudaFuncSetAttribute(phase_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
    }
    if (smem > 232448) return cudaErrorInvalidValue;
    const int64_t total = (int64_t)n_series * chunk_nwin;
    int64_t blocks = (total + warps - 1) / warps;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) return cudaSuccess;
    phase_chain_kernel<<<(unsigned)blocks, warps * 32, smem, stream>>>(
        reinterpret_cast<const double2*>(spectra), spec_nwin, spec_w0, n_series, win_offset, chunk_nwin, nwin, M,
        log2M, phase);
    return cudaGetLastError();
}

}  // namespace ws
