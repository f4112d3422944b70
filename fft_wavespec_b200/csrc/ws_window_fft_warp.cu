// ws_window_fft_warp.cu — per-window real FFT, ONE WARP PER WINDOW (sm_100a).
//
// Same job as ws_window_fft.cu (SURVEY.md section 8a rows A2a/A2b/A3 prologues, A4 transform, A5
// power, A7/A8 selection and rows) for the window lengths of BASELINE.json's configs
// (N = 512 .. 4096) when the windows come from the series itself (no PLA feed) and the phase
// chain is not requested.  The CTA stages a tile of consecutive windows once (and runs the tile-
// level trend IIR), then its warps take windows round-robin and never meet again: the transform
// (ws_warpfft_core.cuh) is in place in a warp-private shared array with __syncwarp between
// passes, the real-input split streams X[k] / X[M-k] straight to the spectra plane as contiguous
// 16-byte stores, and the selection reads the band powers the split left in shared memory.  The
// selected bins' complex values are recomputed from Z (two shared loads each), so no copy of X
// is kept.
//
// Versus the CTA-per-window-group kernel this removes every block barrier from the window loop
// (ncu on the old kernel: 23% of stall samples at barriers, 4x more integer than FP64
// instructions) and halves the shared memory per window in flight.
#include <cstdio>
#include <cstdlib>
#include "ws_common.cuh"
#include "ws_epilogue.cuh"
#include "ws_window_prologue.cuh"
#include "ws_warpfft_core.cuh"

namespace ws {


struct WarpLayout {
    int tile_doubles;     // staged samples (even)
    int nb_alloc;         // band powers per warp (even), 0 when no selection output is requested
    size_t z_off, pw_off, ord_off, total;
};

static WarpLayout warp_layout(const Params& p, int T, int kWarps) {
    WarpLayout L;
    const int M = p.N / 2;
    const int Lt = (T - 1) * p.hop + p.N;
    L.tile_doubles = (Lt + 1) & ~1;
    const bool want_sel = p.bins || p.rows || p.waves || p.contrib;
    int nb = want_sel && p.band_hi >= p.band_lo ? p.band_hi - p.band_lo + 1 : 0;
    L.nb_alloc = (nb + 1) & ~1;
    size_t off = (size_t)L.tile_doubles * 8 + (size_t)((T + 1) & ~1) * 8;
    L.z_off = off;
    size_t zbytes = (size_t)kWarps * M * 16;
    size_t iir = p.detrend == 1 ? (size_t)(L.tile_doubles + kWarps * 32) * 8 : 0;   // overlays Z
    off += zbytes > iir ? zbytes : iir;
    L.pw_off = off; off += (size_t)kWarps * L.nb_alloc * 8;
    L.ord_off = off; off += p.select == 1 ? (size_t)kWarps * L.nb_alloc * 4 : 0;
    L.total = off;
    return L;
}

// kWarps warps per CTA, at least MINB CTAs per SM (bounds the registers)
template <int LN, int kWarps, int MINB>
__global__ void __launch_bounds__(kWarps * 32, MINB)
window_fft_warp_kernel(const Params p, const WarpLayout L) {
    typedef ws_wf::Geo<LN> G;
    constexpr int N = G::N, M = G::M;
    constexpr int kThreads = kWarps * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.tile_windows;
    const int64_t w0 = p.win_offset + (int64_t)blockIdx.x * T;
    const int s = blockIdx.y;
    const int64_t nwin = p.nwin;
    const int64_t wend = p.win_offset + p.chunk_nwin;
    const int tw_count = (int)((w0 + T <= wend) ? T : (wend - w0));
    if (tw_count <= 0) return;
    const int Lt = (tw_count - 1) * p.hop + N;

    double* tile = reinterpret_cast<double*>(smem_raw);
    double* delta = tile + L.tile_doubles;
    double2* Z = reinterpret_cast<double2*>(smem_raw + L.z_off) + (size_t)warp * M;
    double* pwb = reinterpret_cast<double*>(smem_raw + L.pw_off) + (size_t)warp * L.nb_alloc;
    int* ord = reinterpret_cast<int*>(smem_raw + L.ord_off) + (size_t)warp * L.nb_alloc;

    // ---- tile prologue (whole CTA): stage the samples, run the trend IIR once per tile
    const double* src = p.series + (int64_t)s * p.series_stride + w0 * p.hop;
    stage_samples(src, Lt, Lt, tile, tid, kThreads);            // 128-bit loads, split at the 16-byte boundary
    __syncthreads();
    if (p.detrend == 1) {
        // Trend IIR (Legacy/...-kalman-fast.mq5:3367-3379) restarted per window:
        //   tr_w[j] = y[w+j] + alpha^j (2c x[w] - y[w])
        // for ANY y obeying y[a] = c (x[a] + x[a-1]) + alpha y[a-1] on the tile, so y is run once
        // per tile and each window only needs delta_w = 2c x[w] - y[w].
        double* y = reinterpret_cast<double*>(smem_raw + L.z_off);      // Z is not live yet
        double* carry = y + L.tile_doubles;
        const double c = p.iir_c;
        cta_trend_iir<kThreads>(tile, Lt, p.iir_alpha, c, y, carry);
        for (int t = tid; t < tw_count; t += kThreads) {
            const int a = t * p.hop;
            delta[t] = c * (tile[a] + tile[a]) - y[a];
        }
        __syncthreads();
        for (int i = tid; i < Lt; i += kThreads) tile[i] = tile[i] - y[i];
        __syncthreads();
    }

    const bool want_sel = p.bins || p.rows || p.waves || p.contrib;
    const int lo = p.band_lo, hi = p.band_hi;
    const int nband = hi - lo + 1;
    const double2* __restrict__ tw = p.tw;
    auto warp_sync = [] { __syncwarp(); };

    // ---- window loop: warps are independent from here on
    for (int t = warp; t < tw_count; t += kWarps) {
        const int off = t * p.hop;
        double sub = 0.0;
        if (p.detrend == 2) {
            // mean removal (Legacy/WaveSpecZZ_gpu_wip.mq5:940-943)
            double sum = 0.0;
            for (int n = lane; n < N; n += 32) sum += tile[off + n];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) sum += shfl_xor_d(sum, m);
            sub = sum / (double)N;
        } else if (p.detrend == 1) {
            sub = delta[t];
        }
        // pass 0 reads the tile through the prologue; one instance per (detrend, window) case so
        // the per-sample code carries no mode branches
        const double2* wt2 = reinterpret_cast<const double2*>(p.wtab);
        const double2* ap2 = reinterpret_cast<const double2*>(p.apow);
#define WS_PASS0(MODE, WIN)                                                                       \
        {                                                                                         \
            const ProloguePair<MODE, WIN> pr{tile, wt2, ap2, sub};                                \
            ws_wf::dif_pass<LN, G::radix(0), G::stride(0), true>(                                 \
                lane, [&](int m) { return pr(off, m); }, Z, tw);                                  \
        }
        if (p.has_window) {
            if (p.detrend == 1) WS_PASS0(1, true) else if (p.detrend == 2) WS_PASS0(2, true) else WS_PASS0(0, true)
        } else {
            if (p.detrend == 1) WS_PASS0(1, false) else if (p.detrend == 2) WS_PASS0(2, false) else WS_PASS0(0, false)
        }
#undef WS_PASS0
        ws_wf::later_passes<LN, 1>(lane, Z, tw, warp_sync);
        __syncwarp();

        const int64_t gw = (int64_t)s * nwin + w0 + t;
        double2* g = p.spectra
            ? reinterpret_cast<double2*>(p.spectra + ((int64_t)s * p.spec_nwin + (w0 - p.spec_w0) + t) * N) : nullptr;
        double2* bb = p.band_buf
            ? p.band_buf + ((int64_t)s * p.chunk_nwin + (w0 - p.win_offset) + t) * nband - lo : nullptr;
        double* pwa = pwb - lo;
        ws_wf::split_phase<LN>(lane, Z, tw, [&](int k, double2 X) {
            if (g) g[k] = X;
            const bool in_band = k >= lo && k <= hi;
            if (want_sel && in_band) pwa[k] = X.x * X.x + X.y * X.y;
            if (bb && in_band) bb[k] = X;
        });

        if (want_sel) {
            __syncwarp();
            warp_select_emit_x(p, pwa, [&](int b) { return ws_wf::split_bin<LN>(Z, tw, b); }, ord, gw);
        }
        __syncwarp();                   // Z and the band powers are reused by the next window
    }
}

static int warp_pick_tile(const Params& p, int kWarps) {
    // 64 windows per tile when the staged samples stay within 4096 + N doubles; strided batches
    // (hop ~ N) get at least one window per warp while the tile fits 12288 doubles.  N = 4096 keeps
    // its tile short: five 32 KB transform arrays leave ~50 KB for the staged samples.
    long budget = p.N >= 4096 ? 64 + p.N : 4096 + p.N;
    long t = (budget - p.N) / p.hop + 1;
    if (t < kWarps && (long)(kWarps - 1) * p.hop + p.N <= 12288) t = kWarps;
    const long cap = 64 - 64 % kWarps + (64 % kWarps ? kWarps : 0);   // multiple of kWarps: no warp idles on a full tile
    if (t > cap) t = cap;
    if (p.chunk_nwin < t) t = (long)p.chunk_nwin;
    return (int)t;
}

template <int LN, int kWarps, int MINB>
static cudaError_t launch_ln(Params p, cudaStream_t stream) {
    p.tile_windows = warp_pick_tile(p, kWarps);
    const WarpLayout L = warp_layout(p, p.tile_windows, kWarps);
    static std::atomic<unsigned long long> attr_seen{0};
    if (first_launch_on_device(attr_seen)) {
        cudaError_t e = cudaFuncSetAttribute(window_fft_warp_kernel<LN, kWarps, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
    }
    if (L.total > 232448) return cudaErrorInvalidValue;
    dim3 grid((unsigned)((p.chunk_nwin + p.tile_windows - 1) / p.tile_windows), (unsigned)p.n_series);
    window_fft_warp_kernel<LN, kWarps, MINB><<<grid, kWarps * 32, L.total, stream>>>(p, L);
    return cudaGetLastError();
}

// Warps per CTA, measured on B200 (profiles/README.md): 12 warps x 2 CTAs (80 registers) at
// N = 512, 8 warps x 2 CTAs (128 registers) at N = 1024, 12 warps x 1 CTA at N = 2048, 5 warps x 1 CTA
// at N = 4096 (a warp's transform array is 32 KB there).
// WAVESPEC_K1W=0 (tuning hook) routes everything to the CTA kernel of ws_window_fft.cu.
static bool warp_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("WAVESPEC_K1W");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
static bool fits(const Params& p, int W) {
    const int T = warp_pick_tile(p, W);
    if (T < W && p.chunk_nwin >= W) return false;                // tile too wide for a shared stage
    return warp_layout(p, T, W).total <= 232448;
}
// 0: not served.  N = 2048 drops to 8 warps when a wide band's powers do not fit beside 12 Z arrays.
static int warps_for(const Params& p) {
    if (p.N == 1024) return fits(p, 8) ? 8 : 0;
    if (p.N == 4096) return fits(p, 5) ? 5 : (fits(p, 4) ? 4 : 0);
    if (fits(p, 12)) return 12;
    return p.N == 2048 && fits(p, 8) ? 8 : 0;
}

// true when this kernel serves the request (the caller falls back to ws_window_fft.cu otherwise)
bool window_fft_warp_supported(const Params& p) {
    if (!warp_enabled() || p.feed || p.phase) return false;
    if (p.N < 512 || p.N > 4096) return false;    // N = 256: half the lanes idle in the passes, the CTA kernel wins
    if (p.chunk_nwin < 1) return false;
    return warps_for(p) != 0;
}

cudaError_t launch_window_fft_warp(Params p, cudaStream_t stream) {
    const int W = warps_for(p);
    switch (p.N) {
        case 512:  return launch_ln<9, 12, 2>(p, stream);
        case 1024: return launch_ln<10, 8, 2>(p, stream);
        case 2048: return W == 12 ? launch_ln<11, 12, 1>(p, stream) : launch_ln<11, 8, 1>(p, stream);
        case 4096: return W == 5 ? launch_ln<12, 5, 1>(p, stream) : launch_ln<12, 4, 1>(p, stream);
        default:   return cudaErrorInvalidValue;
    }
}

}  // namespace ws
