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

// |x|^2 of a captured bin, as two explicitly rounded operations: the batched epilogue forms it at two
// places and both must give the same bits.
__device__ __forceinline__ double band_power(double2 x) { return __fma_rn(x.x, x.x, __dmul_rn(x.y, x.y)); }

// warp-wide argmax of (p, pos) under `better`; every lane returns the winner.
__device__ __forceinline__ void warp_argbest(double& p, int& pos) {
    // Fast path: the high word of a non-negative double orders like the double.  When exactly one
    // lane holds the maximal high word that lane is the strict winner: one REDUX, one VOTE and
    // three shuffles instead of the 5-stage ladder.  Ties in the top 32 bits (or no candidate at
    // all) take the full comparison below; both conditions are warp-uniform.
    const int hi = (p >= 0.0) ? __double2hiint(p) + 1 : 0;          // 0: no candidate (-1, -2, NaN)
    const int mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned cand = __ballot_sync(0xffffffffu, hi == mh);
    if (mh != 0 && (cand & (cand - 1)) == 0) {
        const int src = __ffs(cand) - 1;
        p = __shfl_sync(0xffffffffu, p, src);
        pos = __shfl_sync(0xffffffffu, pos, src);
        return;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        double op = shfl_xor_d(p, m);
        int opos = __shfl_xor_sync(0xffffffffu, pos, m);
        if (better(op, opos, p, pos)) { p = op; pos = opos; }
    }
}

// Top-K of a wide band (more entries than the lanes can hold in registers) in ONE pass over shared
// memory: every lane walks its entries pw[lane], pw[lane + 32], .. in ascending order and keeps its C
// best in a sorted register list (strict '>' on insertion: an equal power met later never gets ahead
// of the earlier, lower bin); then K rounds of a warp argmax over the list heads, the winning lane
// popping its head.  A lane can own at most K of the K winners, so C >= K makes the result exactly what
// K scans of the whole band give.  NaN powers never enter a list.  Lane r returns winner r.
template <int C>
__device__ __forceinline__ void warp_wide_topk(const double* pw, int n, int K, int lane, int& my_pos, double& my_pow) {
    double v[C];
    int e[C];
#pragma unroll
    for (int j = 0; j < C; j++) { v[j] = -1.0; e[j] = 0x7fffffff; }
    for (int i = lane; i < n; i += 32) {
        const double q = pw[i];
        if (q > v[C - 1]) {
            bool placed = false;
#pragma unroll
            for (int j = C - 1; j >= 1; j--) {
                if (!placed) {
                    if (q > v[j - 1]) { v[j] = v[j - 1]; e[j] = e[j - 1]; }
                    else { v[j] = q; e[j] = i; placed = true; }
                }
            }
            if (!placed) { v[0] = q; e[0] = i; }
        }
    }
    for (int r = 0; r < K; r++) {
        double bp = v[0]; int bpos = e[0];
        if (!(bp >= 0.0)) { bp = -1.0; bpos = 0x7fffffff; }
        warp_argbest(bp, bpos);
        if (bpos == 0x7fffffff) break;                  // band exhausted: remaining slots stay absent
        if (bpos == e[0]) {
#pragma unroll
            for (int j = 0; j + 1 < C; j++) { v[j] = v[j + 1]; e[j] = e[j + 1]; }
            v[C - 1] = -1.0; e[C - 1] = 0x7fffffff;
        }
        if (lane == r) { my_pos = bpos; my_pow = bp; }
    }
}

// pw   : shared, powers of this window indexed by bin; only [band_lo, band_hi] is touched
//        (destroyed: selected entries are overwritten)
// getX : bin -> complex value of this window (called for the K selected bins only)
// ord  : shared int scratch of >= band entries (only used by the SORT rule)
// gw   : global window index = series * nwin + window
template <class GetX>
__device__ __forceinline__ void warp_select_emit_x(const Params& p, double* pw, GetX getX,
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
    } else if (hi - lo < 128) {
        // bands up to 128 bins (the usual cases: 51 bins at N = 1024 / 18-200, 74 at N = 2048 / 18-52):
        // each lane keeps its up to four candidates in registers, so a round is three compares, the warp
        // argmax and a conditional retire — no shared-memory traffic.  Within a lane the candidates are
        // in ascending bin order and the compares are strict, so an equal power never displaces the
        // lower bin.  NaN powers never win (as in the scan below): they are retired up front.
        int bb[4];
        double v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            bb[i] = lo + 32 * i + lane;
            v[i] = -2.0;
            if (bb[i] <= hi) { v[i] = pw[bb[i]]; if (!(v[i] >= 0.0)) v[i] = -2.0; }
        }
        for (int r = 0; r < K; r++) {
            double bp = v[0]; int bpos = bb[0];
#pragma unroll
            for (int i = 1; i < 4; i++) if (v[i] > bp) { bp = v[i]; bpos = bb[i]; }
            if (bp < 0.0) { bp = -1.0; bpos = 0x7fffffff; }
            warp_argbest(bp, bpos);
            if (bpos == 0x7fffffff) break;              // band exhausted: remaining slots stay -1
#pragma unroll
            for (int i = 0; i < 4; i++) if (bpos == bb[i]) v[i] = -2.0;
            if (lane == r) { my_bin = bpos; my_pow = bp; }
        }
    } else if (K <= 8) {
        // wide bands (config 5: 435 bins at N = 4096 / 9-200): one pass, per-lane register lists
        int pos = -1;
        if (K <= 4) warp_wide_topk<4>(pw + lo, hi - lo + 1, K, lane, pos, my_pow);
        else warp_wide_topk<8>(pw + lo, hi - lo + 1, K, lane, pos, my_pow);
        if (pos >= 0) my_bin = lo + pos;
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
        if (my_bin >= 0) { double2 x = getX(my_bin); re = x.x; im = x.y; }
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


// X : shared, N/2 complex bins of this window
__device__ __forceinline__ void warp_select_emit(const Params& p, double* pw, const double2* X,
                                                 int* ord, int64_t gw) {
    warp_select_emit_x(p, pw, [X](int b) { return X[b]; }, ord, gw);
}

// ---- batched form of the insertion rule (A7a): several windows per warp ------------------------
// The warp is split into groups of Lg lanes (Lg a power of two >= K); each group owns one window,
// so the selection shuffles of 32/Lg windows issue together and the row arithmetic (sqrt, atan2)
// runs on all 32 lanes instead of K.  Rows are staged in shared memory and leave as contiguous
// 128-bit stores (rows of consecutive windows are contiguous in the output plane).
//
// pw    : shared, [wpb][band] powers of the wpb windows of this batch (destroyed)
// xb    : shared, [wpb][band] complex bins (band-relative index)
// lo    : first bin of the band;  nvalid : windows of the batch that exist (<= wpb)
// gw0   : global window index of the batch's first window (series * nwin + window)
// stage : shared, per-warp scratch of 32 * 16 doubles (used when row_stride <= 16)
// LG: compile-time group width (8: the common case K <= 8, band <= 64 — scans and shuffle ladders
// fully unrolled, no loop or address arithmetic left); 0: run-time width Lg_rt.
template <int LG>
__device__ __forceinline__ void warp_select_emit_batch(const Params& p, double* pw, const double2* xb,
                                                       int band, int lo, int Lg_rt, int nvalid, int64_t gw0,
                                                       double* stage, bool fast = true) {
    const int lane = threadIdx.x & 31;
    const int Lg = LG ? LG : Lg_rt;
    const int g = lane / Lg, l = lane - g * Lg;
    const int N = p.N, K = p.K;
    const double2* xbb = xb + g * band;
    const bool live = g < nvalid;

    double bsum = 0.0;
    int my_pos = -1;
    double my_pow = -1.0;
    if (LG == 8) {
        // K <= 8, band <= 64: a sorting network instead of K dependent argmax rounds.  Lane l of
        // the group owns band entries l, l+8, .., l+56 (powers straight from the captured bins;
        // entries past the band, NaNs and windows that do not exist are "absent" = -2), sorts
        // them (Batcher, 19 compare-exchanges), then three bitonic merges with the lanes at
        // distance 1, 2, 4 keep the best eight of each union.  Every compare-exchange uses the
        // full (power desc, bin asc) order, so the result is exactly what the reference's
        // ascending insertion scan leaves in its top-K list; all eight lanes end with the same
        // list and lane l takes entry l.  No shared-memory traffic and an 18-stage dependency
        // chain instead of K x (scan + ladder).
        double v[8];
        int e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            e[i] = l + 8 * i;
            double q = -2.0;
            if (live && e[i] < band) q = band_power(xbb[e[i]]);
            if (!(q >= 0.0)) q = -2.0;
            v[i] = q;
            bsum += (q >= 0.0) ? q : 0.0;
        }
#pragma unroll
        for (int m = 4; m >= 1; m >>= 1) bsum += shfl_xor_d(bsum, m);
        // Fast path: the same network on ONE 32-bit key per entry — the power rounded towards zero to
        // float with its low 7 mantissa bits replaced by 64 - entry, so that a larger key is the better
        // entry under (power desc, bin asc) and a compare-exchange is a max and a min (2 instructions
        // instead of ~10 on a (double, int) pair; the exact network below is 980 instructions per batch
        // of four windows, 48 % of everything the headline kernel issues).  Rounding is monotonic, so
        // keys whose power fields differ order the exact powers strictly.  The result is therefore
        // exact whenever the power fields of neighbouring selected entries differ and the best
        // DISCARDED entry's field differs from the last selected one's; any warp that sees a batch where
        // they do not (powers equal to 16 bits: flat markets, quantised prices) runs the exact network.
        bool exact = true;
        if (fast) {
            unsigned key[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
                key[i] = v[i] >= 0.0 ? ((__float_as_uint(__double2float_rz(v[i])) & ~127u) | (unsigned)(64 - e[i])) : 0u;
#define WS_KE(i, j) { const unsigned a = key[i], b = key[j]; key[i] = max(a, b); key[j] = min(a, b); }
            WS_KE(0, 1) WS_KE(2, 3) WS_KE(4, 5) WS_KE(6, 7)
            WS_KE(0, 2) WS_KE(1, 3) WS_KE(4, 6) WS_KE(5, 7)
            WS_KE(1, 2) WS_KE(5, 6)
            WS_KE(0, 4) WS_KE(1, 5) WS_KE(2, 6) WS_KE(3, 7)
            WS_KE(2, 4) WS_KE(3, 5)
            WS_KE(1, 2) WS_KE(3, 4) WS_KE(5, 6)
            unsigned lost = 0u;                         // best key this lane dropped in a merge
#pragma unroll
            for (int d = 1; d <= 4; d <<= 1) {
                unsigned pk[8];
#pragma unroll
                for (int i = 0; i < 8; i++) pk[i] = __shfl_xor_sync(0xffffffffu, key[7 - i], d);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    lost = max(lost, min(key[i], pk[i]));
                    key[i] = max(key[i], pk[i]);
                }
                WS_KE(0, 4) WS_KE(1, 5) WS_KE(2, 6) WS_KE(3, 7)
                WS_KE(0, 2) WS_KE(1, 3) WS_KE(4, 6) WS_KE(5, 7)
                WS_KE(0, 1) WS_KE(2, 3) WS_KE(4, 5) WS_KE(6, 7)
            }
#undef WS_KE
#pragma unroll
            for (int m = 4; m >= 1; m >>= 1) lost = max(lost, __shfl_xor_sync(0xffffffffu, lost, m));
            // entries 0 .. K-1 are taken: each must beat its successor by the power field, the last one
            // the best entry that is not taken (entry K of the list, or the best dropped one at K = 8)
            bool bad = K == 8 && lost != 0u && (lost >> 7) == (key[7] >> 7);
#pragma unroll
            for (int i = 0; i < 7; i++) bad = bad || (i < K && key[i + 1] != 0u && (key[i] >> 7) == (key[i + 1] >> 7));
            exact = __any_sync(0xffffffffu, bad);
            if (!exact) {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (l == i && i < K && key[i] != 0u) my_pos = 64 - (int)(key[i] & 127u);
                if (my_pos >= 0) my_pow = band_power(xbb[my_pos]);      // the same two operations as above
            }
        }
        if (exact) {
#define WS_CE(i, j)                                                                     \
        {                                                                               \
            const bool sw = better(v[j], e[j], v[i], e[i]);                             \
            const double a = v[i], b = v[j];                                            \
            const int ea = e[i], eb = e[j];                                             \
            v[i] = sw ? b : a; v[j] = sw ? a : b;                                       \
            e[i] = sw ? eb : ea; e[j] = sw ? ea : eb;                                   \
        }
        WS_CE(0, 1) WS_CE(2, 3) WS_CE(4, 5) WS_CE(6, 7)
        WS_CE(0, 2) WS_CE(1, 3) WS_CE(4, 6) WS_CE(5, 7)
        WS_CE(1, 2) WS_CE(5, 6)
        WS_CE(0, 4) WS_CE(1, 5) WS_CE(2, 6) WS_CE(3, 7)
        WS_CE(2, 4) WS_CE(3, 5)
        WS_CE(1, 2) WS_CE(3, 4) WS_CE(5, 6)
        // the three merge rounds share one copy of the code (the epilogue runs once per group of
        // windows: straight-line code this long is fetched from L2 every time)
#pragma unroll 1
        for (int d = 1; d <= 4; d <<= 1) {
            double pv[8];
            int pe[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                pv[i] = shfl_xor_d(v[7 - i], d);
                pe[i] = __shfl_xor_sync(0xffffffffu, e[7 - i], d);
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (better(pv[i], pe[i], v[i], e[i])) { v[i] = pv[i]; e[i] = pe[i]; }
            WS_CE(0, 4) WS_CE(1, 5) WS_CE(2, 6) WS_CE(3, 7)
            WS_CE(0, 2) WS_CE(1, 3) WS_CE(4, 6) WS_CE(5, 7)
            WS_CE(0, 1) WS_CE(2, 3) WS_CE(4, 5) WS_CE(6, 7)
        }
#undef WS_CE
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (l == i && i < K && v[i] >= 0.0) { my_pos = e[i]; my_pow = v[i]; }
        }
    } else {
        double* pwb = pw + g * band;
        if (live) for (int e = l; e < band; e += Lg) bsum += pwb[e];
        for (int m = Lg >> 1; m >= 1; m >>= 1) bsum += shfl_xor_d(bsum, m);
        if (Lg == 32 && K <= 8) {
            // one window per warp, wide band: one pass with per-lane register lists
            if (K <= 4) warp_wide_topk<4>(pwb, live ? band : 0, K, lane, my_pos, my_pow);
            else warp_wide_topk<8>(pwb, live ? band : 0, K, lane, my_pos, my_pow);
        } else
        for (int r = 0; r < K; r++) {
            double bp = -1.0; int bpos = 0x7fffffff;
            // ascending scan inside a lane: an equal power met later never displaces the earlier
            // (lower) bin, so strict '>' alone implements the tie rule here
            if (live) {
                for (int e = l; e < band; e += Lg) {
                    double v = pwb[e];
                    if (v > bp) { bp = v; bpos = e; }
                }
            }
            {
                // Cross-lane argmax of (power desc, position asc) inside the group.  Fast path: the
                // high word of a non-negative double orders like the double; when exactly one lane
                // of every group holds the maximal high word that lane is the winner and is simply
                // broadcast.  Otherwise (ties in the top 32 bits) fall back to the full comparison.
                const int hi = (bp >= 0.0) ? __double2hiint(bp) + 1 : 0;      // 0 = no candidate
                int mh = hi;
                for (int m = Lg >> 1; m >= 1; m >>= 1) mh = max(mh, __shfl_xor_sync(0xffffffffu, mh, m));
                const unsigned cand = __ballot_sync(0xffffffffu, hi == mh && mh != 0);
                const unsigned gmask = (Lg == 32 ? 0xffffffffu : ((1u << Lg) - 1u)) << (g * Lg);
                const unsigned mine = cand & gmask;
                const bool unique = (mine & (mine - 1)) == 0;                 // 0 or 1 candidate
                if (__all_sync(0xffffffffu, unique)) {
                    const int srcl = mine ? (__ffs(mine) - 1) : lane;
                    bp = __shfl_sync(0xffffffffu, bp, srcl);
                    bpos = __shfl_sync(0xffffffffu, bpos, srcl);
                    if (!mine) { bp = -1.0; bpos = 0x7fffffff; }
                } else {
                    for (int m = Lg >> 1; m >= 1; m >>= 1) {
                        double op = shfl_xor_d(bp, m);
                        int opos = __shfl_xor_sync(0xffffffffu, bpos, m);
                        if (better(op, opos, bp, bpos)) { bp = op; bpos = opos; }
                    }
                }
            }
            if (bpos != 0x7fffffff) {
                if ((bpos & (Lg - 1)) == l) pwb[bpos] = -2.0;      // the owning lane retires it (-2 < -1)
                if (l == r) { my_pos = bpos; my_pow = bp; }
            }
        }
    }
    const bool has_row = live && l < K;
    const int my_bin = my_pos >= 0 ? lo + my_pos : -1;
    double re = 0.0, im = 0.0;
    if (has_row && my_pos >= 0) { const double2 x = xbb[my_pos]; re = x.x; im = x.y; }
    const int64_t slot = (gw0 + g) * K + l;
    if (has_row) {
        if (p.bins) p.bins[slot] = my_bin;
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
                double sn, cs;
                sincos(2.0 * kPi * my_bin * nn / N, &sn, &cs);
                cv = (2.0 / N) * (re * cs - im * sn);
            }
            p.contrib[slot] = cv;
        }
    }
    if (p.rows) {
        double f[kRowFields];
#pragma unroll
        for (int i = 0; i < kRowFields; i++) f[i] = 0.0;
        if (has_row && my_bin > 0) {
            // N is a power of two: the divisions by N are exact scalings.  The period N/k and the factor
            // 1/(2 pi freq) come from the per-bin table when the caller provides one (same correctly
            // rounded period; the eta fields differ from the divided form by an ulp, inside their 1e-9 bar)
            const double invN = 1.0 / (double)N;
            f[0] = 2.0 * sqrt(my_pow) * invN;
            f[1] = (double)my_bin * invN;
            double bars_per_rad;
            if (p.rowtab) { const double2 t = __ldg(p.rowtab + my_bin); f[2] = t.x; bars_per_rad = t.y; }
            else { f[2] = (double)N / (double)my_bin; bars_per_rad = 1.0 / (2.0 * kPi * f[1]); }
            // phase at the newest sample: atan2 + 2 pi k (N-1)/N + pi/2, wrapped to [-pi, pi].
            // 2 pi k (N-1)/N == -2 pi k/N (mod 2 pi); the small angle keeps the wrap to one step.
            double ph = atan2(im, re) + (0.5 * kPi - 2.0 * kPi * (double)my_bin * invN);
            if (ph > kPi) ph -= 2.0 * kPi;
            if (ph < -kPi) ph += 2.0 * kPi;
            f[3] = ph;
            double d = 0.5 * kPi - ph;                  // bars to the next extremum of amp*sin
            if (d < 0.0) d += kPi;
            if (d >= kPi) d -= kPi;
            f[4] = d * bars_per_rad;
            f[5] = f[4] * p.sample_rate_seconds;
            f[6] = bsum > 0.0 ? my_pow / bsum : 0.0;
        }
        const int rs = p.row_stride;
        if (rs <= 16) {
            // stage, then stream out contiguously
            if (has_row) {
                double* srow = stage + (g * K + l) * rs;
#pragma unroll
                for (int i = 0; i < kRowFields; i++) if (i < rs) srow[i] = f[i];
                if (rs > kRowFields) srow[kRowFields] = 0.0;
            }
            __syncwarp();
            const int total = nvalid * K * rs;                       // doubles of this batch
            double* dst = p.rows + gw0 * K * (int64_t)rs;
            if ((((gw0 * K * (int64_t)rs) | total) & 1) == 0) {
                const double2* s2 = reinterpret_cast<const double2*>(stage);
                double2* d2 = reinterpret_cast<double2*>(dst);
                for (int i = lane; i < (total >> 1); i += 32) __stcs(d2 + i, s2[i]);
            } else {
                for (int i = lane; i < total; i += 32) __stcs(dst + i, stage[i]);
            }
            __syncwarp();
        } else if (has_row) {
            double* row = p.rows + slot * (int64_t)rs;
#pragma unroll
            for (int i = 0; i < kRowFields; i++) if (i < rs) row[i] = f[i];
            for (int i = kRowFields; i < rs; i++) row[i] = 0.0;
        }
    }
}

}  // namespace ws
