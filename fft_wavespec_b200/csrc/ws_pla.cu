// ws_pla.cu — piecewise-linear approximation feed:
// FitPlaSegment / ComputePlaSegmentError / PlaSplit / BuildPlaPriceSeries,
// Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:387-502.
//
// The recursion is data dependent and is re-run on every window, and the integer segment bounds
// (pivots) must match the CPU statement exactly.  So: one THREAD per window walks the recursion
// with an explicit stack and accumulates each least-squares sum in ascending index order exactly
// as the reference does (this file is compiled with -fmad=false, every operation separately
// rounded) — parallelism comes from the windows.  Adjacent lanes read adjacent samples
// (window w+1 starts one sample later), so the global loads coalesce.  The line is rendered
// warp-cooperatively so the stores coalesce too.
//
// Reference quirk kept on purpose: when the worst sample is the first of a segment, PlaSplit
// recurses on [s,s] and on the same [s,e] again (:462-467), appending single-point segments
// until the budget `count + 2 <= max_segments` is exhausted.
#include "ws_common.cuh"
#include "ws_series.h"

namespace ws {

constexpr int kPlaStack = 128;
constexpr int kPlaSegCap = 60;    // segments kept per window in shared memory for rendering

struct PlaSeg { int s, e; double slope, icpt; };

__device__ __forceinline__ void render_direct(const PlaSeg sg, double* line, int32_t* bounds, int q, int bounds_cap,
                                              int N) {
    if (line)
        for (int i = sg.s; i <= sg.e && i < N; ++i) line[i] = sg.slope * (double)i + sg.icpt;
    if (bounds && q < bounds_cap) { bounds[2 * q] = sg.s; bounds[2 * q + 1] = sg.e; }
}

__global__ void __launch_bounds__(32)
pla_kernel(const double* __restrict__ series, int64_t series_stride, int64_t nwin, int32_t N,
           int32_t hop, int32_t max_segments, double max_error, double* __restrict__ lines,
           int32_t* __restrict__ seg_bounds, int32_t* __restrict__ seg_counts, int32_t bounds_cap,
           int32_t* __restrict__ overflow) {
    __shared__ PlaSeg segs[32][kPlaSegCap];
    __shared__ int seg_n[32];
    const int lane = threadIdx.x;
    const int sidx = blockIdx.y;
    const int64_t w = (int64_t)blockIdx.x * 32 + lane;
    const bool active = w < nwin;
    const double* y = series + (int64_t)sidx * series_stride + (active ? w : 0) * hop;

    int count = 0;
    bool direct = false;
    const int64_t gw_me = (int64_t)sidx * nwin + (active ? w : 0);
    double* line_w = lines ? lines + gw_me * N : nullptr;
    int32_t* bounds_w = seg_bounds ? seg_bounds + gw_me * bounds_cap * 2 : nullptr;
    if (active) {
        int stk_s[kPlaStack], stk_e[kPlaStack];
        int sp = 0;
        stk_s[0] = 0; stk_e[0] = N - 1; sp = 1;
        const int maxseg = max_segments < 1 ? 1 : max_segments;
        const double maxerr = max_error < 1e-8 ? 1e-8 : max_error;
        while (sp > 0) {
            --sp;
            const int s = stk_s[sp], e = stk_e[sp];
            double slope = 0.0, icpt = 0.0;
            bool leaf = true;
            if (s >= e) {
                icpt = y[s];
                if (e < s) continue;                    // AppendPlaSegment ignores end < start
            } else {
                // FitPlaSegment (:387-417)
                const int n = e - s + 1;
                double sx = 0.0, sy = 0.0, sx2 = 0.0, sxy = 0.0;
                for (int i = s; i <= e; ++i) {
                    const double x = (double)i, v = y[i];
                    sx += x; sy += v; sx2 += x * x; sxy += x * v;
                }
                const double denom = (double)n * sx2 - sx * sx;
                if (fabs(denom) < 1e-9) { slope = 0.0; icpt = sy / (double)n; }
                else {
                    slope = ((double)n * sxy - sx * sy) / denom;
                    icpt = (sy - slope * sx) / (double)n;
                }
                // ComputePlaSegmentError (:419-440)
                double mx = 0.0;
                int worst = s;
                for (int i = s; i <= e; ++i) {
                    const double approx = slope * (double)i + icpt;
                    const double err = fabs(y[i] - approx);
                    if (err > mx) { mx = err; worst = i; }
                }
                const bool can_split = (count + 2) <= maxseg && (e - s) > 1;
                if (can_split && mx > maxerr) {
                    const int left_end = s > worst - 1 ? s : worst - 1;
                    const int right_start = e < worst ? e : worst;
                    if (sp + 2 > kPlaStack) { atomicExch(overflow, 1); sp = 0; break; }
                    stk_s[sp] = right_start; stk_e[sp] = e; ++sp;      // processed after the left part
                    stk_s[sp] = s; stk_e[sp] = left_end; ++sp;
                    leaf = false;
                }
            }
            if (leaf) {
                if (count < kPlaSegCap && !direct) {
                    segs[lane][count] = PlaSeg{s, e, slope, icpt};
                } else {
                    // more segments than the shared staging holds (deep left spines are not bounded
                    // by max_segments): this window renders by itself, in append order
                    if (!direct) {
                        direct = true;
                        for (int q = 0; q < count; q++) render_direct(segs[lane][q], line_w, bounds_w, q, bounds_cap, N);
                    }
                    render_direct(PlaSeg{s, e, slope, icpt}, line_w, bounds_w, count, bounds_cap, N);
                }
                ++count;
            }
        }
    }
    seg_n[lane] = direct ? -count : count;      // negative: already rendered by its own thread
    __syncwarp();

    // render + pivots, one window at a time, whole warp cooperating
    for (int wl = 0; wl < 32; wl++) {
        const int64_t ww = (int64_t)blockIdx.x * 32 + wl;
        if (ww >= nwin) break;
        const int cn_raw = seg_n[wl];
        const int64_t gw = (int64_t)sidx * nwin + ww;
        if (cn_raw < 0) {                           // direct-rendered window: only the count is left to write
            if (seg_counts && lane == 0) seg_counts[gw] = -cn_raw;
            if (seg_bounds)
                for (int q = lane; q < bounds_cap; q += 32)
                    if (q >= -cn_raw) { seg_bounds[(gw * bounds_cap + q) * 2] = -1; seg_bounds[(gw * bounds_cap + q) * 2 + 1] = -1; }
            continue;
        }
        const int cn = cn_raw;
        if (lines) {
            double* line = lines + gw * N;
            for (int q = 0; q < cn; q++) {
                const PlaSeg sg = segs[wl][q];
                for (int i = sg.s + lane; i <= sg.e && i < N; i += 32)
                    line[i] = sg.slope * (double)i + sg.icpt;
                __syncwarp();   // later segments overwrite earlier ones, as in :487-495
            }
        }
        if (seg_counts && lane == 0) seg_counts[gw] = cn;
        if (seg_bounds) {
            for (int q = lane; q < bounds_cap; q += 32) {
                int32_t* b = seg_bounds + (gw * bounds_cap + q) * 2;
                if (q < cn) { b[0] = segs[wl][q].s; b[1] = segs[wl][q].e; }
                else { b[0] = -1; b[1] = -1; }
            }
        }
    }
}

// `overflow` is a device int the caller has zeroed on `stream`; it is set when a window's recursion
// outgrew the on-chip stack (the caller reads it back before it trusts the feed).
cudaError_t launch_pla(const double* series, int64_t series_stride, int32_t n_series, int64_t nwin,
                       int32_t N, int32_t hop, int32_t max_segments, double max_error, double* lines,
                       int32_t* seg_bounds, int32_t* seg_counts, int32_t bounds_cap, int32_t* overflow,
                       cudaStream_t stream) {
    dim3 grid((unsigned)((nwin + 31) / 32), (unsigned)n_series);
    pla_kernel<<<grid, 32, 0, stream>>>(series, series_stride, nwin, N, hop, max_segments, max_error,
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
