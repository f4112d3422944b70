// ws_warpfft_core.cuh — arithmetic of the warp-per-window real FFT (ws_window_fft_warp.cu),
// written as host/device functions of an explicit lane index so the same code runs in the CUDA
// kernel and, lane by lane and phase by phase, in the CPU emulation test (tests/emu/emu_warpfft.cpp).
//
// One warp transforms one window of N = 2^LN real samples as an M = N/2-point complex FFT of
// z[m] = v[2m] + i v[2m+1], in place in a warp-private shared array Z[M]:
//
//   * decimation in frequency, radix 8 while three bits remain, then one radix 4 or 2 pass; the
//     last pass has stride 1 and therefore no twiddles.  A butterfly reads R elements and writes
//     the same R positions, so lanes never touch each other's data inside a pass and the only
//     synchronisation is one __syncwarp between passes — no CTA barrier in the window loop.
//   * pass 0 reads the staged samples straight from the tile (through the prologue: detrend and
//     window function), so a window costs log8(M) shared-memory round trips.
//   * the result sits in digit-reversed order; the real-input split gathers Z[rev k] and
//     Z[rev (M-k)], forms X[k] and X[M-k] together and hands them to a sink (global spectra
//     plane, band powers for the selection, band hand-off buffer).
//   * bank conflicts: element x lives at x ^ d1 ^ d2 ^ d3 (d_i = i-th octal digit of x).  Every
//     access pattern of the passes and of the digit-reversed gather varies one octal digit (or
//     the bits of one stride-R group) across a quarter warp, so the low three bits — the 16-byte
//     bank group — take eight distinct values.  The map is GF(2)-linear: swz(base + r S) =
//     swz(base) ^ swz(r S) whenever the bits are disjoint, one LOP3 per access.
//
// Twiddles are exact table values W_N^t (no recurrences): results agree with the oracle's
// radix-2 FFT to ~1e-14 relative, inside the 1e-9 bar of the parity tests.
#pragma once

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WF_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#ifndef WS_HOST_DOUBLE2
#define WS_HOST_DOUBLE2
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif
#define WF_HD inline
#endif

namespace ws_wf {

WF_HD double2 c_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
WF_HD double2 c_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
WF_HD double2 c_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

WF_HD double2 ld_tw(const double2* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// pass plan of an M = 2^(LN-1) point transform
template <int LN>
struct Geo {
    static constexpr int N = 1 << LN, M = N / 2, LM = LN - 1;
    static constexpr int P8 = LM / 3, REM = LM % 3, P = P8 + (REM ? 1 : 0);
    static WF_HD constexpr int radix(int p) { return p < P8 ? 8 : (1 << REM); }
    static WF_HD constexpr int stride(int p) {              // distance between the R inputs of a pass-p butterfly
        int s = M;
        for (int q = 0; q <= p; q++) s /= radix(q);
        return s;
    }
    // The twiddles of the radix-8 passes live BEHIND the N-entry table exp(-2 pi i m / N), one block per
    // twiddled pass (stride S > 1) in pass order, each laid out [r - 1][i], r = 1..7, i < S, holding
    // W_{8S}^{i r}: the lanes of a warp have consecutive i, so each of the seven loads of a butterfly
    // is one contiguous run (4 L1 wavefronts per warp at S >= 32, a single line at S = 8) — gathering
    // W^t, W^2t, W^4t from the N-entry table cost 56 wavefronts per warp in pass 0 at N = 1024 plus
    // four complex products per butterfly, on a kernel whose L1 data pipe is 86 % busy (ncu).
    static WF_HD constexpr int tw_off(int S) {             // offset of the block of the pass with stride S
        int off = 0, s = M;
        for (int q = 0; q < P8; q++) { s /= 8; if (s > S) off += 7 * s; }
        return off;
    }
    // position of output bin k after the last pass (digit reversal over the pass radices)
    static WF_HD int rev(int k) {
        int pos = 0, rem = k;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int R = radix(p);
            pos += (rem & (R - 1)) * stride(p);
            rem /= R;
        }
        return pos;
    }
};

// Host side: number of pass-twiddle entries behind the N-entry table of an N = 2^LN transform, and
// their values, taken from the table itself (entry [r - 1][i] of the pass with stride S is
// tw[(N / 8S) i r]).  Mirrors Geo<LN>::tw_off for a run-time LN.
inline int pass_twiddle_count(int LN) {
    const int LM = LN - 1, P8 = LM / 3;
    int s = 1 << LM, total = 0;
    for (int q = 0; q < P8; q++) { s /= 8; if (s > 1) total += 7 * s; }
    return total;
}
inline void fill_pass_twiddles(int LN, const double2* tw, double2* out) {
    const int N = 1 << LN, LM = LN - 1, P8 = LM / 3;
    int s = 1 << LM;
    for (int q = 0; q < P8; q++) {
        s /= 8;
        if (s <= 1) continue;
        for (int r = 1; r < 8; r++)
            for (int i = 0; i < s; i++) *out++ = tw[(N / (8 * s)) * i * r];
    }
}

// conflict-free placement (see header)
WF_HD constexpr int swz(int x) { return x ^ ((x >> 3) & 7) ^ ((x >> 6) & 7) ^ ((x >> 9) & 7); }

WF_HD void bfly8(double2* a) {
    // forward 8-point DFT, inputs and outputs in natural order
    const double h = 0.70710678118654752440;
    double2 b0 = c_add(a[0], a[4]), b1 = c_add(a[1], a[5]), b2 = c_add(a[2], a[6]), b3 = c_add(a[3], a[7]);
    double2 d0 = c_sub(a[0], a[4]), t1 = c_sub(a[1], a[5]), t2 = c_sub(a[2], a[6]), t3 = c_sub(a[3], a[7]);
    double2 d1 = make_double2((t1.x + t1.y) * h, (t1.y - t1.x) * h);        // * W8
    double2 d2 = make_double2(t2.y, -t2.x);                                   // * -i
    double2 d3 = make_double2((t3.y - t3.x) * h, -(t3.x + t3.y) * h);       // * W8^3
    double2 c0 = c_add(b0, b2), c1 = c_add(b1, b3), c2 = c_sub(b0, b2), u = c_sub(b1, b3);
    double2 c3 = make_double2(u.y, -u.x);
    double2 e0 = c_add(d0, d2), e1 = c_add(d1, d3), e2 = c_sub(d0, d2), v = c_sub(d1, d3);
    double2 e3 = make_double2(v.y, -v.x);
    a[0] = c_add(c0, c1); a[4] = c_sub(c0, c1); a[2] = c_add(c2, c3); a[6] = c_sub(c2, c3);
    a[1] = c_add(e0, e1); a[5] = c_sub(e0, e1); a[3] = c_add(e2, e3); a[7] = c_sub(e2, e3);
}

WF_HD void bfly4(double2* a) {
    double2 b0 = c_add(a[0], a[2]), b1 = c_sub(a[0], a[2]), b2 = c_add(a[1], a[3]);
    double2 d = c_sub(a[1], a[3]);
    double2 b3 = make_double2(d.y, -d.x);   // -i (a1 - a3)
    a[0] = c_add(b0, b2); a[2] = c_sub(b0, b2); a[1] = c_add(b1, b3); a[3] = c_sub(b1, b3);
}

// One in-place DIF pass of radix R and input stride S over the M points of one window.
// LOAD(x) returns element x of the pass input (pass 0: the prologue applied to the tile; later
// passes: Z[swz(x)]).  Butterfly j = B S + i works on x = B R S + i + r S, r < R, and multiplies
// output r by W_{RS}^{i r} = W_N^{(N / RS) i r}.
template <int LN, int R, int S, bool FIRST, class Load>
WF_HD void dif_pass(int lane, Load load, double2* Z, const double2* tw) {
    typedef Geo<LN> G;
    constexpr int NB = G::M / R;                  // butterflies per window
    constexpr int IT = (NB + 31) / 32;
#pragma unroll
    for (int b = 0; b < IT; b++) {
        const int j = lane + 32 * b;
        if (NB < 32 * IT && j >= NB) break;
        const int i = j & (S - 1);
        const int base = (j - i) * R + i;
        const int sb = swz(base);
        double2 a[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            if constexpr (FIRST) a[r] = load(base + r * S);
            else a[r] = Z[sb ^ swz(r * S)];
        }
        if (R == 8) bfly8(a);
        else if (R == 4) bfly4(a);
        else { double2 x0 = a[0], x1 = a[1]; a[0] = c_add(x0, x1); a[1] = c_sub(x0, x1); }
        if (S > 1) {
            // only radix-8 passes carry twiddles (the short pass is last): seven exact table values
            // from the pass's own block (see Geo::tw_off)
            const double2* tp = tw + G::N + G::tw_off(S) + i;
#pragma unroll
            for (int r = 1; r < R; r++) a[r] = c_mul(a[r], ld_tw(tp + (r - 1) * S));
        }
#pragma unroll
        for (int r = 0; r < R; r++) Z[sb ^ swz(r * S)] = a[r];
    }
}

// passes 1 .. P-1 (pass 0 is issued by the caller, which owns the prologue); SYNC() separates them
template <int LN, int PASS, class Sync>
WF_HD void later_passes(int lane, double2* Z, const double2* tw, Sync sync) {
    typedef Geo<LN> G;
    if constexpr (PASS < G::P) {
        sync();
        dif_pass<LN, G::radix(PASS), G::stride(PASS), false>(lane, 0, Z, tw);
        later_passes<LN, PASS + 1>(lane, Z, tw, sync);
    }
}

// X[k] and X[M-k] of the real transform from the digit-reversed complex result, 0 < k < M/2;
// also valid for the self-paired bins k = 0 and k = M/2 (then xm is a duplicate of xk).
template <int LN>
WF_HD void split_pair(const double2* Z, const double2* tw, int k, double2& xk, double2& xm) {
    typedef Geo<LN> G;
    const double2 zk = Z[swz(G::rev(k))];
    const double2 zr = Z[swz(G::rev((G::M - k) & (G::M - 1)))];
    const double2 E = make_double2(0.5 * (zk.x + zr.x), 0.5 * (zk.y - zr.y));
    const double2 D = make_double2(zk.x - zr.x, zk.y + zr.y);      // zk - conj(zr)
    const double2 O = make_double2(0.5 * D.y, -0.5 * D.x);         // -i/2 D
    const double2 T = c_mul(ld_tw(tw + k), O);
    xk = c_add(E, T);
    const double2 m = c_sub(E, T);
    xm = make_double2(m.x, -m.y);
}

// bin b (0 <= b < M) recomputed exactly as split_phase produced it
template <int LN>
WF_HD double2 split_bin(const double2* Z, const double2* tw, int b) {
    typedef Geo<LN> G;
    const int k = b <= G::M / 2 ? b : G::M - b;
    double2 xk, xm;
    split_pair<LN>(Z, tw, k, xk, xm);
    return b == k ? xk : xm;
}

// real-input split over the whole window: SINK(k, X[k]) for every bin 0 <= k < M exactly once
template <int LN, class Sink>
WF_HD void split_phase(int lane, const double2* Z, const double2* tw, Sink sink) {
    typedef Geo<LN> G;
    constexpr int H = G::M / 2;
    constexpr int IT = (H + 31) / 32;
#pragma unroll
    for (int it = 0; it < IT; it++) {
        const int k = lane + 32 * it;
        if (H < 32 * IT && k >= H) break;
        double2 xk, xm;
        split_pair<LN>(Z, tw, k, xk, xm);
        sink(k, xk);
        if (k > 0) sink(G::M - k, xm);
        else if (H > 0) {                         // lane 0 also owns the self-paired bin M/2
            double2 yk, ym;
            split_pair<LN>(Z, tw, H, yk, ym);
            sink(H, yk);
        }
    }
}

// ---- inverse real transform on the same passes ---------------------------------------------------
// With Z = E + iO the M-point spectrum of z[m] = x[2m] + i x[2m+1],
//     E[k] = (X[k] + conj X[M-k]) / 2,   O[k] = (X[k] - conj X[M-k]) / 2 * e^{+2 pi i k / N},
// and z = conj(FFT(conj Z)) / M.  inverse_unpack writes conj Z[k] and conj Z[M-k] (swizzled) from
// the pair xa = X[k], xb = X[M-k] (the caller passes xb = 0 for k = 0: the Nyquist bin is not part
// of the contract; k = 0 and k = M/2 are self-paired); after the forward passes inverse_pair reads
// (x[2m], x[2m+1]).
template <int LN>
WF_HD void inverse_unpack(int k, double2 xa, double2 xb, const double2* tw, double2* Z) {
    typedef Geo<LN> G;
    const int km = (G::M - k) & (G::M - 1);
    const double2 wa = ld_tw(tw + k);                                       // e^{-2 pi i k / N}
    {
        const double2 E = make_double2(0.5 * (xa.x + xb.x), 0.5 * (xa.y - xb.y));
        const double2 D = make_double2(0.5 * (xa.x - xb.x), 0.5 * (xa.y + xb.y));
        const double2 O = make_double2(D.x * wa.x + D.y * wa.y, D.y * wa.x - D.x * wa.y);   // D * conj(wa)
        Z[swz(k)] = make_double2(E.x - O.y, -(E.y + O.x));                                  // conj(E + iO)
    }
    if (k != 0 && k != G::M / 2) {
        const double2 wb = ld_tw(tw + km);
        const double2 E = make_double2(0.5 * (xb.x + xa.x), 0.5 * (xb.y - xa.y));
        const double2 D = make_double2(0.5 * (xb.x - xa.x), 0.5 * (xb.y + xa.y));
        const double2 O = make_double2(D.x * wb.x + D.y * wb.y, D.y * wb.x - D.x * wb.y);
        Z[swz(km)] = make_double2(E.x - O.y, -(E.y + O.x));
    }
}

template <int LN>
WF_HD double2 inverse_pair(const double2* Z, int m) {
    typedef Geo<LN> G;
    const double sc = 1.0 / (double)G::M;
    const double2 z = Z[swz(G::rev(m))];
    return make_double2(z.x * sc, -z.y * sc);
}

}  // namespace ws_wf
