// ws_rows.cu — top-K cycle rows from the in-band bins of each window (sm_100a).
//
// Second half of the plain hop-1 path.  The sliding kernel (ws_sliding.cu) streams the spectra to
// HBM and hands the in-band bins of every window to this kernel through a compact band buffer
// [window][band] (16 * band bytes per window: 816 B next to the 8 KiB spectrum at N = 1024, band
// 18-200).  Splitting here keeps the sliding kernel a pure streaming writer (its register budget
// allows only two CTAs per SM, so an in-kernel epilogue phase stalls the store stream) and lets
// the selection run at full occupancy: one THREAD per window runs the reference's insertion
// top-K (Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:537-554: first slot with
// p > top[s], shift the rest down) in registers — no shuffles, no dependent cross-lane steps —
// then packs rows / bins / waves / contributions (:559-568; WaveSpecZZ_1.1.0-gpuopt.mq5:329).
// The band buffer is read with coalesced 128-bit loads and transposed through shared memory so
// the per-window scan reads consecutive words.  K <= 8.
#include "ws_common.cuh"
#include "ws_epilogue.cuh"
#include "ws_series.h"

namespace ws {

constexpr int kRowsThreads = 128;   // windows per CTA, one per thread
constexpr int kRowsChunk = 32;      // band bins staged per pass
constexpr int kRowsPitch = kRowsThreads + 1;

__global__ void __launch_bounds__(kRowsThreads)
rows_from_band_kernel(const Params p) {
    // pwT[bin][window]: powers of one band chunk, transposed so that the per-window scan below
    // reads consecutive words (the global reads above it are coalesced along the bins)
    __shared__ double pwT[kRowsChunk * kRowsPitch];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t total = (int64_t)p.n_series * p.chunk_nwin;
    const int64_t g0 = (int64_t)blockIdx.x * kRowsThreads;
    const int nloc = (int)((total - g0) < kRowsThreads ? (total - g0) : kRowsThreads);
    const int N = p.N, K = p.K;
    int64_t gw = 0;
    if (tid < nloc) {
        const int64_t gid = g0 + tid;
        const int64_t s = gid / p.chunk_nwin;
        gw = s * p.nwin + p.win_offset + (gid - s * p.chunk_nwin);
    }
    const int lo = p.band_lo, hi = p.band_hi;
    const int band = hi - lo + 1;
    const double2* __restrict__ Ball = p.band_buf + g0 * band;          // this CTA's windows, contiguous
    const double2* __restrict__ B = Ball + (int64_t)tid * band - lo;    // B[bin] for this thread's window

    double tv[8]; int tp[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { tv[i] = -1.0; tp[i] = -1; }
    double bsum = 0.0;
    for (int c0 = lo; c0 <= hi; c0 += kRowsChunk) {
        const int cn = (hi + 1 - c0) < kRowsChunk ? (hi + 1 - c0) : kRowsChunk;
        // cooperative, coalesced load of the chunk: a warp takes one window at a time, lanes run
        // along the bins (512-byte contiguous reads)
        // (eight windows' loads are issued before the first use so that each lane keeps eight
        // 128-bit reads in flight; a single dependent load per trip leaves HBM latency exposed)
        constexpr int kWarps = kRowsThreads / 32, kMlp = 8;
        for (int wl0 = warp; wl0 < nloc; wl0 += kWarps * kMlp) {
            for (int e = lane; e < cn; e += 32) {
                double2 x[kMlp];
#pragma unroll
                for (int u = 0; u < kMlp; u++) {
                    const int wl = wl0 + u * kWarps;
                    x[u] = (wl < nloc) ? __ldcs(Ball + (int64_t)wl * band + (c0 - lo) + e) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < kMlp; u++) {
                    const int wl = wl0 + u * kWarps;
                    if (wl < nloc) pwT[e * kRowsPitch + wl] = x[u].x * x[u].x + x[u].y * x[u].y;
                }
            }
        }
        __syncthreads();
        if (tid < nloc) {
            for (int e = 0; e < cn; e++) {
                double v = pwT[e * kRowsPitch + tid];
                bsum += v;
                int q = c0 + e;
                bool carried = false;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const bool sw = carried || (v > tv[i]);      // insert at the first p > top[i] ...
                    const double ov = tv[i]; const int oq = tp[i];
                    tv[i] = sw ? v : ov; tp[i] = sw ? q : oq;    // ... and shift the rest down
                    v = sw ? ov : v; q = sw ? oq : q;
                    carried = sw;
                }
            }
        }
        __syncthreads();
    }
    if (tid >= nloc) return;
    const double nn = (double)(N - 1);
    const int rs = p.row_stride;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (r >= K) break;
        const int bin = tp[r];
        const double pw = tv[r];
        const int64_t slot = gw * K + r;
        double re = 0.0, im = 0.0;
        if (bin >= 0) { const double2 x = __ldg(B + bin); re = x.x; im = x.y; }
        if (p.bins) p.bins[slot] = bin;
        if (p.waves) {
            double wv = 0.0;
            if (bin > 0) {
                double mag = sqrt(pw);
                double ph = atan2(im, re);
                wv = (mag / (double)N) * cos(ph + 2.0 * kPi * (double)bin * nn / (double)N);
            }
            p.waves[slot] = wv;
        }
        if (p.contrib) {
            double cv = 0.0;
            if (bin >= 0) {
                double sn, cs;
                sincos(2.0 * kPi * bin * nn / N, &sn, &cs);
                cv = (2.0 / N) * (re * cs - im * sn);
            }
            p.contrib[slot] = cv;
        }
        if (p.rows) {
            double f[kRowFields];
#pragma unroll
            for (int i = 0; i < kRowFields; i++) f[i] = 0.0;
            if (bin > 0) {
                f[0] = 2.0 * sqrt(pw) / (double)N;
                f[1] = (double)bin / (double)N;
                f[2] = (double)N / (double)bin;
                double ph = atan2(im, re) + (0.5 * kPi - 2.0 * kPi * (double)bin / (double)N);
                if (ph > kPi) ph -= 2.0 * kPi;
                if (ph < -kPi) ph += 2.0 * kPi;
                f[3] = ph;
                double d = 0.5 * kPi - ph;
                if (d < 0.0) d += kPi;
                if (d >= kPi) d -= kPi;
                f[4] = d / (2.0 * kPi * f[1]);
                f[5] = f[4] * p.sample_rate_seconds;
                f[6] = bsum > 0.0 ? pw / bsum : 0.0;
            }
            double* row = p.rows + slot * (int64_t)rs;
#pragma unroll
            for (int i = 0; i < kRowFields; i++) if (i < rs) __stcs(row + i, f[i]);
            for (int i = kRowFields; i < rs; i++) row[i] = 0.0;
        }
    }
}

bool rows_from_band_supported(const Params& p) {
    return p.select == 0 && p.K <= 8 && p.band_hi >= p.band_lo && (p.bins || p.rows || p.waves || p.contrib);
}

cudaError_t launch_rows_from_band(const Params& p, cudaStream_t stream) {
    const int64_t total = (int64_t)p.n_series * p.chunk_nwin;
    const unsigned blocks = (unsigned)((total + kRowsThreads - 1) / kRowsThreads);
    rows_from_band_kernel<<<blocks, kRowsThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace ws
