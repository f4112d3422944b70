// ws_epilogue.cuh — per-window epilogue run by ONE warp: band-limited top-K selection with the
// reference's tie rules, then result rows / last-sample reconstruction for the selected bins.
//
// Reference rows of SURVEY.md section 8(a): A7a (insertion top-K,
// Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:537-554), A7b (swap selection sort,
// Legacy/WaveSpecZZ_1.0.4-kalman.mq5:143-180), A8a (:559-568 of the former), A8b (:182-192 of the
// latter), A14 row layout (WaveSpecZZ_1.1.0-gpuopt.mq5:329).
#pragma once
#include "ws_common.cuh"

namespace ws {

constexpr double kPi = 3.14159265358979323846;

// warp-wide argmax of (p, pos) under `better`; every lane returns the winner.
__device__ __forceinline__ void warp_argbest(double& p, int& pos) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        double op = shfl_xor_d(p, m);
        int opos = __shfl_xor_sync(0xffffffffu, pos, m);
        if (better(op, opos, p, pos)) { p = op; pos = opos; }
    }
}

// pw  : shared, N/2 powers of this window (destroyed: selected entries are overwritten)
// X   : shared, N/2 complex bins of this window
// ord : shared int scratch of >= band entries (only used by the SORT rule)
// gw  : global window index = series * nwin + window
__device__ __forceinline__ void warp_select_emit(const Params& p, double* pw, const double2* X,
                                                 int* ord, int64_t gw) {
    const int lane = threadIdx.x & 31;
    const int N = p.N;
    const int K = p.K;
    int lo = p.band_lo, hi = p.band_hi;
    if (p.select == 1 && lo < 1) lo = 1;           // 1.0.4-kalman.mq5:149  k = max(1, min_idx)

    // band energy (row field 6); summation order differs from a serial loop by rounding only
    double bsum = 0.0;
    for (int b = lo + lane; b <= hi; b += 32) bsum += pw[b];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) bsum += shfl_xor_d(bsum, m);

    int my_bin = -1;          // lane r < K ends up owning slot r
    double my_pow = -1.0;

    if (p.select == 1) {
        const int band = hi - lo + 1;
        for (int i = lane; i < band; i += 32) ord[i] = lo + i;
        __syncwarp();
        const int rounds = K < band ? K : band;
        for (int r = 0; r < rounds; r++) {
            double bp = -1.0; int bpos = 0x7fffffff;
            for (int i = r + lane; i < band; i += 32) {
                double v = pw[ord[i]];
                // selection sort keeps cycles[r] unless something is strictly greater, and the
                // first strictly-greatest later element wins: (power desc, position asc)
                if (better(v, i, bp, bpos)) { bp = v; bpos = i; }
            }
            warp_argbest(bp, bpos);
            if (bpos == 0x7fffffff) { bpos = r; bp = pw[ord[r]]; }   // only NaNs left: keep order
            int sel = ord[bpos];
            __syncwarp();
            if (lane == 0) { int t = ord[r]; ord[r] = sel; ord[bpos] = t; }
            __syncwarp();
            if (lane == r) { my_bin = sel; my_pow = bp; }
        }
    } else {
        for (int r = 0; r < K; r++) {
            double bp = -1.0; int bpos = 0x7fffffff;
            for (int b = lo + lane; b <= hi; b += 32) {
                double v = pw[b];
                if (better(v, b, bp, bpos)) { bp = v; bpos = b; }
            }
            warp_argbest(bp, bpos);
            if (bpos == 0x7fffffff) break;              // band exhausted: remaining slots stay -1
            if (lane == 0) pw[bpos] = -2.0;             // excluded from later rounds (-2 < -1)
            __syncwarp();
            if (lane == r) { my_bin = bpos; my_pow = bp; }
        }
    }

    if (lane < K) {
        const int64_t slot = gw * K + lane;
        if (p.bins) p.bins[slot] = my_bin;
        double re = 0.0, im = 0.0;
        if (my_bin >= 0) { double2 x = X[my_bin]; re = x.x; im = x.y; }
        const double nn = (double)(N - 1);
        if (p.waves) {
            double wv = 0.0;
            if (my_bin > 0) {
                double mag = sqrt(my_pow);
                double ph = atan2(im, re);
                wv = (mag / (double)N) * cos(ph + 2.0 * kPi * (double)my_bin * nn / (double)N);
            }
            p.waves[slot] = wv;
        }
        if (p.contrib) {
            double cv = 0.0;
            if (my_bin >= 0) {
                double s, c;
                sincos(2.0 * kPi * my_bin * nn / N, &s, &c);
                cv = (2.0 / N) * (re * c - im * s);
            }
            p.contrib[slot] = cv;
        }
        if (p.rows) {
            double f[kRowFields];
#pragma unroll
            for (int i = 0; i < kRowFields; i++) f[i] = 0.0;
            if (my_bin > 0) {
                f[0] = 2.0 * sqrt(my_pow) / (double)N;
                f[1] = (double)my_bin / (double)N;
                f[2] = (double)N / (double)my_bin;
                double ph = atan2(im, re) + 2.0 * kPi * (double)my_bin * nn / (double)N + 0.5 * kPi;
                ph = remainder(ph, 2.0 * kPi);
                f[3] = ph;
                double d = fmod(0.5 * kPi - ph, kPi);
                if (d < 0.0) d += kPi;
                f[4] = d / (2.0 * kPi * f[1]);
                f[5] = f[4] * p.sample_rate_seconds;
                f[6] = bsum > 0.0 ? my_pow / bsum : 0.0;
            }
            double* row = p.rows + slot * (int64_t)p.row_stride;
            const int m = p.row_stride < kRowFields ? p.row_stride : kRowFields;
#pragma unroll
            for (int i = 0; i < kRowFields; i++) if (i < m) row[i] = f[i];
            for (int i = kRowFields; i < p.row_stride; i++) row[i] = 0.0;
        }
    }
}

}  // namespace ws
