// ws_pla.cu — piecewise-linear approximation feed:
// FitPlaSegment / ComputePlaSegmentError / PlaSplit / BuildPlaPriceSeries,
// Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:387-502.
//
// The recursion is data dependent and is re-run on every window, and the integer segment bounds
// (pivots) must match the CPU statement exactly.  So: one THREAD per window walks the recursion
// with an explicit stack and accumulates each least-squares sum in ascending index order exactly
// as the reference does (this file is compiled with -fmad=false, every operation separately
// rounded) — parallelism comes from the windows.
//
// The 32 windows of a warp are neighbours (window w+1 starts one sample later), so their recursion
// trees are nearly the same — but only nearly, and a warp whose lanes drift apart inside a
// data-dependent `while` executes them one after the other (measured on the first version of this
// kernel: 176 cycles per loop iteration, 9.8 M windows/s).  The walk is therefore written in warp
// LOCKSTEP: every trip of the outer loop pops one node per lane and the warp reconverges after each of
// the fit / error / render loops of the trip (lanes with shorter segments drop out of a loop early, so
// a trip costs its longest segment).  Each lane's arithmetic and its order are untouched.
// (Measured and rejected: giving a trip only to the lanes whose next segment is the same one in
// absolute bar coordinates, plus the lanes with shorter segments — 1.74x instead of 2.07x a lane's own
// work in simulation, but the extra trips cost it back on the device: 16.3 vs 17.1 M windows/s.)
// Leaves are rendered by the lane that owns them, in append order (later segments overwrite earlier
// ones as in :487-495); no shared memory, so occupancy is bounded by registers only.
//
// Reference quirk kept on purpose: when the worst sample is the first of a segment, PlaSplit
// recurses on [s,s] and on the same [s,e] again (:462-467), appending single-point segments
// until the budget `count + 2 <= max_segments` is exhausted.
#include <cstdlib>
#include "ws_common.cuh"
#include "ws_series.h"

namespace ws {

constexpr int kPlaStack = 128;
constexpr int kPlaThreads = 128;

__global__ void __launch_bounds__(kPlaThreads)
pla_kernel(const double* __restrict__ series, int64_t series_stride, int64_t nwin, int32_t N,
           int32_t hop, int32_t max_segments, double max_error, double* __restrict__ lines,
           int32_t* __restrict__ seg_bounds, int32_t* __restrict__ seg_counts, int32_t bounds_cap,
           int32_t* __restrict__ overflow) {
    constexpr unsigned kFull = 0xffffffffu;
    const int sidx = blockIdx.y;
    const int64_t w = (int64_t)blockIdx.x * kPlaThreads + threadIdx.x;
    const bool active = w < nwin;
    const double* y = series + (int64_t)sidx * series_stride + (active ? w : 0) * hop;
    const int64_t gw = (int64_t)sidx * nwin + (active ? w : 0);
    double* line_w = lines ? lines + gw * N : nullptr;
    int32_t* bounds_w = seg_bounds ? seg_bounds + gw * bounds_cap * 2 : nullptr;

    int stk_s[kPlaStack], stk_e[kPlaStack];
    int sp = 0;
    if (active) { stk_s[0] = 0; stk_e[0] = N - 1; sp = 1; }
    int count = 0;
    const int maxseg = max_segments < 1 ? 1 : max_segments;
    const double maxerr = max_error < 1e-8 ? 1e-8 : max_error;

    while (__any_sync(kFull, sp > 0)) {
        const bool on = sp > 0;
        int s = 0, e = -1;
        if (on) { --sp; s = stk_s[sp]; e = stk_e[sp]; }
        const bool fit = on && s < e;
        const int len = fit ? e - s + 1 : 0;

        // FitPlaSegment (:387-417).  sy and sxy are accumulated in ascending index order exactly as the
        // reference does.  sx = sum i and sx2 = sum i*i are sums of integers far below 2^53: every
        // partial sum of the reference's loop is exact, so the closed forms below are bit-identical to
        // it.  x runs as a double incremented by 1.0 (exact) instead of a conversion per sample.
        // The loops have per-lane bounds: lanes with shorter segments drop out and the warp reconverges
        // at the __syncwarp — the trip costs its longest segment, without per-iteration predicate math.
        double sy = 0.0, sxy = 0.0;
        {
            double x = (double)s;
            const double* yp = y + s;
            for (int t = 0; t < len; ++t) {
                const double v = yp[t];
                sy += v; sxy += x * v;
                x += 1.0;
            }
        }
        __syncwarp();
        const long long ls = s, le = e, ln = len;
        const double sx = (double)((ls + le) * ln / 2);
        auto sq = [](long long n) { return n * (n + 1) * (2 * n + 1) / 6; };          // sum_{i=0}^{n} i^2
        const double sx2 = (double)(sq(le) - (ls > 0 ? sq(ls - 1) : 0));
        double slope = 0.0, icpt = 0.0;
        bool leaf = on && e >= s;                       // AppendPlaSegment ignores end < start
        if (on && s == e) icpt = y[s];
        if (fit) {
            const double denom = (double)len * sx2 - sx * sx;
            if (fabs(denom) < 1e-9) { slope = 0.0; icpt = sy / (double)len; }
            else {
                slope = ((double)len * sxy - sx * sy) / denom;
                icpt = (sy - slope * sx) / (double)len;
            }
        }
        // ComputePlaSegmentError (:419-440)
        double mx = 0.0;
        int worst = s;
        {
            double x = (double)s;
            const double* yp = y + s;
            for (int t = 0; t < len; ++t) {
                const double approx = slope * x + icpt;
                const double err = fabs(yp[t] - approx);
                if (err > mx) { mx = err; worst = s + t; }
                x += 1.0;
            }
        }
        __syncwarp();
        if (fit) {
            const bool can_split = (count + 2) <= maxseg && (e - s) > 1;
            if (can_split && mx > maxerr) {
                const int left_end = s > worst - 1 ? s : worst - 1;
                const int right_start = e < worst ? e : worst;
                if (sp + 2 > kPlaStack) { atomicExch(overflow, 1); sp = 0; }
                else {
                    stk_s[sp] = right_start; stk_e[sp] = e; ++sp;      // processed after the left part
                    stk_s[sp] = s; stk_e[sp] = left_end; ++sp;
                }
                leaf = false;
            }
        }
        // leaves: rendered by their own lane, in append order
        int rlen = 0;
        if (leaf) {
            const int last = e < N ? e : N - 1;
            rlen = last - s + 1;
            if (bounds_w && count < bounds_cap) { bounds_w[2 * count] = s; bounds_w[2 * count + 1] = e; }
            ++count;
        }
        if (line_w) {
            double x = (double)s;
            double* lp = line_w + s;
            for (int t = 0; t < rlen; ++t) { lp[t] = slope * x + icpt; x += 1.0; }
            __syncwarp();
        }
    }
    if (active) {
        if (seg_counts) seg_counts[gw] = count;
        if (bounds_w)
            for (int q = count; q < bounds_cap; ++q) { bounds_w[2 * q] = -1; bounds_w[2 * q + 1] = -1; }
    }
}

// `overflow` is a device int the caller has zeroed on `stream`; it is set when a window's recursion
// outgrew the on-chip stack (the caller reads it back before it trusts the feed).
cudaError_t launch_pla(const double* series, int64_t series_stride, int32_t n_series, int64_t nwin,
                       int32_t N, int32_t hop, int32_t max_segments, double max_error, double* lines,
                       int32_t* seg_bounds, int32_t* seg_counts, int32_t bounds_cap, int32_t* overflow,
                       cudaStream_t stream) {
    dim3 grid((unsigned)((nwin + kPlaThreads - 1) / kPlaThreads), (unsigned)n_series);
    // The walk is bound by the FP64 pipe, which eight warps per scheduler already fill.  An (unused)
    // dynamic shared-memory request of 27 KB caps the residency at 8 CTAs per SM: a launch then runs
    // in twice as many waves and its last, partly filled wave costs half as much.
    static int smem_kb = -1;
    if (smem_kb < 0) { const char* e = getenv("WAVESPEC_PLA_SMEM_KB"); smem_kb = e ? atoi(e) : 27; }
    pla_kernel<<<grid, kPlaThreads, smem_kb * 1024, stream>>>(series, series_stride, nwin, N, hop, max_segments, max_error,
                                                 lines, seg_bounds, seg_counts, bounds_cap, overflow);
    return cudaGetLastError();
}

// z[s][wa + w] = feed[s][w][N-1]: the newest sample of every PLA line is the Kalman4D measurement
// (Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3354-3360)
__global__ void gather_last_kernel(const double* __restrict__ feed, int64_t cn, int32_t N, double* __restrict__ z,
                                   int64_t z_stride, int64_t wa) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (w < cn) z[(int64_t)s * z_stride + wa + w] = feed[((int64_t)s * cn + w) * N + (N - 1)];
}

cudaError_t launch_gather_last(const double* feed, int32_t n_series, int64_t cn, int32_t N, double* z,
                               int64_t z_stride, int64_t wa, cudaStream_t stream) {
    dim3 grid((unsigned)((cn + 255) / 256), (unsigned)n_series);
    gather_last_kernel<<<grid, 256, 0, stream>>>(feed, cn, N, z, z_stride, wa);
    return cudaGetLastError();
}

}  // namespace ws
