// ws_sliding.cu — shared-butterfly sliding real FFT for the plain hop-1 path (sm_100a).
//
// This is the headline kernel: the per-bar spectrum of every window of a series with no
// detrend and no window function (WaveSpecZZ_1.1.0-gpuopt.mq5:1239-1241 and
// Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:515-533), followed by the
// band-limited top-K selection and row pack (:537-568 of the latter).
//
// One CTA owns a tile of T consecutive windows of one series:
//   1. the T + N - 1 samples of the tile are staged in shared memory (each sample is read from
//      HBM once per tile);
//   2. the deepest decimation level is computed directly from the samples (2..16-point DFTs);
//   3. fused radix-8 passes (ws_sliding_core.cuh) build the level-3 vectors in shared memory,
//      sharing every sub-transform between the overlapping windows of the tile;
//   4. the top pass produces the N/2 complex bins of each window in registers and streams them
//      to HBM with 128-bit stores (a warp writes contiguous 512-byte runs); in-band bins are
//      also captured in shared memory;
//   5. one warp per window runs the top-K epilogue (ws_epilogue.cuh) on the captured band.
//
// Roofline: the only mandatory HBM traffic is 8*hop bytes in + 8*N bytes out per spectrum; the
// arithmetic is ~N/2 packed butterflies per window (~4.6 kflop-instr at N = 1024), far below the
// FP64 pipe limit, so the kernel is bound by the spectrum store.
#include "ws_common.cuh"
#include "ws_epilogue.cuh"
#include "ws_series.h"
#include "ws_sliding_core.cuh"

namespace ws {

using ws_slide::Plan;

constexpr int kSlideThreads = 256;

struct SlideLayout {
    int x_doubles;      // staged samples (even count)
    int arena_off;      // byte offsets inside dynamic shared memory
    int xb_off, pw_off, ord_off;
    int band;           // captured bins per window
    int total_bytes;
};

struct TopSink {
    double2* g;             // spectra of this series (nullptr: not requested)
    double2* xb;            // shared band capture (nullptr: no selection outputs)
    int N2;                 // N/2 slots per window
    int lo, hi, band;
    int64_t w0, nwin;
    __device__ __forceinline__ void put(int pos, int idx, double2 v) {
        const int64_t w = w0 + pos;
        if (w >= nwin) return;
        if (idx == 0) v.y = 0.0;     // slot 0 carries the Nyquist bin in .y: the contract drops it
        if (g) __stcs(g + w * N2 + idx, v);
        if (xb && idx >= lo && idx <= hi) xb[pos * band + (idx - lo)] = v;
    }
};

__global__ void __launch_bounds__(kSlideThreads, 2)
sliding_shared_kernel(const Params p, const Plan pl, const SlideLayout lay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* x = reinterpret_cast<double*>(smem_raw);
    double2* arena = reinterpret_cast<double2*>(smem_raw + lay.arena_off);
    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const int64_t w0 = p.win_offset + (int64_t)blockIdx.x * pl.T;
    const int64_t wend = p.win_offset + p.chunk_nwin;
    const double* src = p.series + (int64_t)s * p.series_stride;

    // 1. stage the tile (+ halo); beyond the series the samples only feed windows that are never stored
    for (int i = tid; i < pl.x_len; i += kSlideThreads) {
        int64_t a = w0 + i;
        x[i] = (a < p.series_len) ? src[a] : 0.0;
    }
    __syncthreads();
    // 2. deepest level straight from the samples
    ws_slide::bottom_level(tid, kSlideThreads, x, pl, p.tw, arena);
    __syncthreads();
    // 3. fused passes down to level 3
    for (int i = pl.nst; i >= 2; i--) {
        ws_slide::SmemSink sink{arena + pl.off[i - 1], pl.stride[i - 1]};
        ws_slide::fused_pass(tid, kSlideThreads, arena + pl.off[i], pl.stride[i], pl.Q[i], 1 << (3 * (i - 1)),
                             pl.P[i - 1], 1, p.tw, pl.N, 3 * (i - 1), sink);
        __syncthreads();
    }
    // 4. top pass: level 3 -> full spectra, streamed to HBM
    const bool want_sel = (p.bins || p.rows || p.waves || p.contrib) && lay.band > 0;
    TopSink top;
    top.g = p.spectra ? reinterpret_cast<double2*>(p.spectra) + (int64_t)s * p.nwin * (pl.N / 2) : nullptr;
    top.xb = want_sel ? reinterpret_cast<double2*>(smem_raw + lay.xb_off) : nullptr;
    top.N2 = pl.N / 2;
    top.lo = p.band_lo; top.hi = p.band_hi; top.band = lay.band;
    if (p.select == 1 && top.lo < 1) top.lo = 1;
    top.w0 = w0; top.nwin = wend;
    ws_slide::fused_pass(tid, kSlideThreads, arena + pl.off[1], pl.stride[1], pl.Q[1], 1, pl.T, pl.S, p.tw,
                         pl.N, 0, top);
    if (!(p.bins || p.rows || p.waves || p.contrib)) return;
    __syncthreads();
    // 5. warp-per-window selection + rows
    const int lane = tid & 31, warp = tid >> 5;
    int nvalid = (int)((wend - w0) < pl.T ? (wend - w0) : pl.T);
    if (lay.band > 0) {
        double* pw = reinterpret_cast<double*>(smem_raw + lay.pw_off) + warp * lay.band;
        int* ord = reinterpret_cast<int*>(smem_raw + lay.ord_off) + warp * lay.band;
        const int lo = top.lo;
        for (int wl = warp; wl < nvalid; wl += kSlideThreads / 32) {
            const double2* xb = top.xb + wl * lay.band;
            for (int b = lane; b < lay.band; b += 32) { double2 v = xb[b]; pw[b] = v.x * v.x + v.y * v.y; }
            __syncwarp();
            warp_select_emit(p, pw - lo, xb - lo, ord, (int64_t)s * p.nwin + w0 + wl);
            __syncwarp();
        }
    } else {
        // empty band: every slot is absent
        for (int wl = warp; wl < nvalid; wl += kSlideThreads / 32)
            warp_select_emit(p, nullptr, nullptr, nullptr, (int64_t)s * p.nwin + w0 + wl);
    }
}

static bool pick_plan(const Params& p, Plan& pl, SlideLayout& lay) {
    int T, S;
    switch (p.N) {
        case 256:  T = 128; S = 16; break;
        case 512:  T = 64;  S = 8;  break;
        case 1024: T = 32;  S = 4;  break;
        case 2048: T = 16;  S = 2;  break;
        case 4096: T = 8;   S = 1;  break;
        default: return false;
    }
    if (!ws_slide::plan_make(pl, p.N, T, S)) return false;
    int lo = p.band_lo, hi = p.band_hi;
    if (p.select == 1 && lo < 1) lo = 1;
    const bool sel = p.bins || p.rows || p.waves || p.contrib;
    lay.band = (sel && hi >= lo) ? hi - lo + 1 : 0;
    lay.x_doubles = (pl.x_len + 1) & ~1;
    lay.arena_off = lay.x_doubles * 8;
    const int below3 = lay.arena_off + pl.off[1] * 16;               // bytes below the level-3 array
    const int work_end = lay.arena_off + pl.arena_slots * 16;
    const int xb_bytes = pl.T * lay.band * 16;
    // the band capture is written while level 3 is still being read: it may only reuse what
    // lies below level 3; pw / ord are used after the barrier and may overlay anything
    lay.xb_off = (xb_bytes <= below3) ? 0 : work_end;
    lay.pw_off = lay.xb_off + xb_bytes;
    const int warps = kSlideThreads / 32;
    lay.ord_off = lay.pw_off + warps * lay.band * 8;
    int end = lay.ord_off + warps * lay.band * 4;
    lay.total_bytes = end > work_end ? end : work_end;
    lay.total_bytes = (lay.total_bytes + 15) & ~15;
    return lay.total_bytes <= 232448;
}

bool sliding_shared_supported(const Params& p) {
    if (p.hop != 1 || p.detrend != 0 || p.has_window || p.feed || p.phase) return false;
    Plan pl; SlideLayout lay;
    return pick_plan(p, pl, lay);
}

cudaError_t launch_sliding_shared(Params p, cudaStream_t stream) {
    Plan pl; SlideLayout lay;
    if (!pick_plan(p, pl, lay)) return cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(sliding_shared_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    p.tile_windows = pl.T;
    dim3 grid((unsigned)((p.chunk_nwin + pl.T - 1) / pl.T), (unsigned)p.n_series);
    sliding_shared_kernel<<<grid, kSlideThreads, lay.total_bytes, stream>>>(p, pl, lay);
    return cudaGetLastError();
}

}  // namespace ws
