// ws_sliding_core.cuh — the arithmetic of the shared-butterfly sliding real FFT, written as
// host/device functions of an explicit (thread id, thread count) so the same code runs inside
// the CUDA kernel (ws_sliding.cu) and, thread by thread, in the CPU emulation test
// (tests/emu_sliding.cpp).
//
// Idea.  For hop 1 the radix-2 decimation-in-time sub-transforms of neighbouring windows are the
// SAME numbers: with F(s, m)[k] = sum_j x[m + 2^s j] W_{L}^{jk}, L = N / 2^s (the L-point DFT of
// the stride-2^s samples starting at absolute offset m),
//        F(s, m)[k] = F(s+1, m)[k] + W_L^k F(s+1, m + 2^s)[k]
// and the "odd" half of window m is the "even" half of window m + 2^s.  So a tile of T
// consecutive windows needs each F(s, m) once: ~N/2 real butterflies per window instead of
// (N/2) log2 N — the same exact-twiddle radix-2 arithmetic as a per-window FFT (no recursive
// sliding-DFT error growth), with the work shared between overlapping windows.
//
// Real input: every F(s, m) is Hermitian, kept PACKED in L/2 complex slots: slot 0 = (F[0], F[L/2])
// (both real), slot k = F[k] for 0 < k < L/2.  One packed butterfly on slot k of two vectors
// A = F(s+1, m), B = F(s+1, m + 2^s) (Q = L/4 slots each) with t = W_L^k B[k] gives
//        F(s,m)[k] = A[k] + t          F(s,m)[2Q - k] = conj(A[k] - t)          0 < k < Q
//        slot 0 = (A0 + B0, A0 - B0)   F(s,m)[Q] = (A[Q], -B[Q])                (A[Q], B[Q] real)
//
// Three levels are fused per pass (radix-8 in registers).  The top pass (levels 3 -> 0), which is
// ~85% of the arithmetic and produces the output, runs as register CHAINS: a thread owns slot k of
// level 3, walks consecutive windows m, m+1, ... keeping the intermediate levels in registers,
// reads ONE new slot and emits EIGHT bins per window with 7 packed butterflies.  The lower
// passes are chain-free (12 butterflies per item) so that the whole CTA can work on them.
#pragma once

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WS_HD __host__ __device__ __forceinline__
#else
#include <cmath>
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#define WS_HD inline
#endif

namespace ws_slide {

// Geometry of one tile.  Level indices: 0 = full window transform.  The top pass keeps `top`
// levels in registers (3: radix-8 chain reading level 3; 2: radix-4 chain reading level 2, half the
// registers per thread); below it chain-free radix-8 passes connect the stored levels
// lev[i] = top + 3 (i - 1), i = 1..nst; the deepest one, sb = lev[nst], is computed directly from
// the samples (Lb-point DFT, Lb <= 16).
struct Plan {
    int N, log2N;
    int T;            // windows per tile
    int S;            // sub-chains of the top pass (T % S == 0, (T/S) % 4 == 0)
    int top;          // levels fused in the top pass: 3 or 2
    int nst;          // stored levels
    int sb;           // bottom level = lev[nst]
    int Lb;           // bottom DFT length = N >> sb  (2..16)
    int lev[4];       // lev[i] = level index of stored level i (lev[0] = 0)
    int Q[4];         // slots per vector at stored level i : Q[i] = N >> (lev[i] + 1)   (Q[0] = N/2)
    int P[4];         // positions held at stored level i   : P[i] = T + (1 << lev[i]) - 1
    int stride[4];    // slots between consecutive positions in shared memory (padded)
    int off[4];       // slot offset of stored level i inside the vector arena (i >= 1)
    int x_len;        // staged samples: T + N - 1
    int arena_slots;  // total double2 slots of the vector arena
};

// odd, so that the lanes of a warp that hold the same slot of consecutive positions (direct_pass, the
// bottom level) never share a bank group
WS_HD int top_stride(int q) { return q + 1; }

WS_HD bool plan_make(Plan& pl, int N, int T, int S, int top = 3) {
    pl.N = N;
    int l = 0;
    while ((1 << l) < N) l++;
    pl.log2N = l;
    if ((1 << l) != N || N < 256 || N > 4096) return false;
    if (top != 2 && top != 3) return false;
    pl.T = T; pl.S = S; pl.top = top;
    if (S < 1 || T % S || (T / S) % 4) return false;
    // deepest stored level: the smallest lev = top + 3 j whose direct DFT length N >> lev is <= 16
    pl.nst = 1;
    while ((N >> (top + 3 * (pl.nst - 1))) > 16) pl.nst++;
    if (pl.nst > 3) return false;
    pl.lev[0] = 0;
    for (int i = 1; i <= pl.nst; i++) pl.lev[i] = top + 3 * (i - 1);
    pl.sb = pl.lev[pl.nst];
    pl.Lb = N >> pl.sb;
    if (pl.Lb < 2) return false;
    for (int i = 0; i <= pl.nst; i++) {
        pl.Q[i] = N >> (pl.lev[i] + 1);
        pl.P[i] = T + (1 << pl.lev[i]) - 1;
        pl.stride[i] = top_stride(pl.Q[i]);
    }
    // arena order: deepest level first, the level read by the top pass last, so that the epilogue
    // buffers can reuse everything below it while the top pass still reads it
    int off = 0;
    for (int i = pl.nst; i >= 1; i--) { pl.off[i] = off; off += pl.P[i] * pl.stride[i]; }
    pl.off[0] = 0;
    pl.arena_slots = off;
    pl.x_len = T + N - 1;
    return true;
}

WS_HD double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// packed butterfly on a general slot: P = A + wB, R = conj(A - wB)
WS_HD void bfly(double2 A, double2 B, double2 w, double2& P, double2& R) {
    double2 t = cmul(w, B);
    P = make_double2(A.x + t.x, A.y + t.y);
    R = make_double2(A.x - t.x, t.y - A.y);
}

// ---- bottom level: Lb-point packed real DFT of x[p + j * step], computed per position --------
// Recursive halving with the same packed butterfly; Lb in {2,4,8,16}.  out has Lb/2 slots.
template <int L>
struct SmallRdft {
    // tw: N-entry table exp(-2 pi i m / N); W_L^k = tw[(N/L) k]
    static WS_HD void run(const double* x, int step, const double2* tw, int N, double2* out) {
        double2 a[L / 4 > 0 ? L / 4 : 1], b[L / 4 > 0 ? L / 4 : 1];
        SmallRdft<L / 2>::run(x, 2 * step, tw, N, a);
        SmallRdft<L / 2>::run(x + step, 2 * step, tw, N, b);
        const int Q = L / 4;                      // slots of the halves
        out[0] = make_double2(a[0].x + b[0].x, a[0].x - b[0].x);
        out[Q] = make_double2(a[0].y, -b[0].y);
        for (int k = 1; k < Q; k++) {
            double2 P, R;
            bfly(a[k], b[k], tw[(N / L) * k], P, R);
            out[k] = P;
            out[2 * Q - k] = R;
        }
    }
};
template <>
struct SmallRdft<2> {
    static WS_HD void run(const double* x, int step, const double2*, int, double2* out) {
        out[0] = make_double2(x[0] + x[step], x[0] - x[step]);
    }
};

// ---- radix-8 packed butterflies (levels s+3 -> s) -------------------------------------------
// Output slot order of a general slot k (Q = slots of the input level):
//   o[0]=k  o[1]=8Q-k  o[2]=4Q-k  o[3]=4Q+k  o[4]=2Q-k  o[5]=6Q+k  o[6]=2Q+k  o[7]=6Q-k
struct Tw7 { double2 w2, w1a, w1b, w0a, w0b, w0c, w0d; };

WS_HD Tw7 load_tw7(const double2* tw, int N, int L /* DFT length of the output level = 16Q */, int Q, int k) {
    const int u2 = N / (L / 4), u1 = N / (L / 2), u0 = N / L;
    Tw7 t;
    t.w2 = tw[u2 * k];
    t.w1a = tw[u1 * k];           t.w1b = tw[u1 * (2 * Q - k)];
    t.w0a = tw[u0 * k];           t.w0b = tw[u0 * (4 * Q - k)];
    t.w0c = tw[u0 * (2 * Q - k)]; t.w0d = tw[u0 * (2 * Q + k)];
    return t;
}

WS_HD void slot_index8(int Q, int k, int idx[8]) {
    idx[0] = k;         idx[1] = 8 * Q - k; idx[2] = 4 * Q - k; idx[3] = 4 * Q + k;
    idx[4] = 2 * Q - k; idx[5] = 6 * Q + k; idx[6] = 2 * Q + k; idx[7] = 6 * Q - k;
}

// last level of the radix-8 group: F(s,m) = F1(m) (+) F1(m+D)
WS_HD void level0_general(const double2 f1[4], const double2 g[4], const Tw7& t, double2 o[8]) {
    bfly(f1[0], g[0], t.w0a, o[0], o[1]);
    bfly(f1[1], g[1], t.w0b, o[2], o[3]);
    bfly(f1[2], g[2], t.w0c, o[4], o[5]);
    bfly(f1[3], g[3], t.w0d, o[6], o[7]);
}

// chain-free form: the eight inputs I[j] = In(m + j D)[k] -> the eight output slots (12 butterflies)
WS_HD void radix8_general(const double2 I[8], const Tw7& t, double2 o[8]) {
    double2 aP, aR, bP, bR, cP, cR, dP, dR;
    bfly(I[0], I[4], t.w2, aP, aR);     // F2(m)
    bfly(I[1], I[5], t.w2, bP, bR);     // F2(m+D)
    bfly(I[2], I[6], t.w2, cP, cR);     // F2(m+2D)
    bfly(I[3], I[7], t.w2, dP, dR);     // F2(m+3D)
    double2 f1[4], g[4];
    bfly(aP, cP, t.w1a, f1[0], f1[1]);  // F1(m)
    bfly(aR, cR, t.w1b, f1[2], f1[3]);
    bfly(bP, dP, t.w1a, g[0], g[1]);    // F1(m+D)
    bfly(bR, dR, t.w1b, g[2], g[3]);
    level0_general(f1, g, t, o);
}

// packed slot 0 of the input level: (F[0], F[Q]) both real.  Output slots, in this order:
//   o[0]=slot 0 packed (F[0], F[8Q])  o[1]=4Q  o[2]=2Q  o[3]=6Q  o[4]=Q  o[5]=7Q  o[6]=3Q  o[7]=5Q
struct Tw4s { double2 w8, wa, wb, wc; };
WS_HD Tw4s load_tw4s(const double2* tw, int N, int L, int Q) {
    const int u1 = N / (L / 2), u0 = N / L;
    Tw4s t;
    t.w8 = tw[u1 * Q];          // W_{8Q}^{Q}   = W_8
    t.wa = tw[u0 * 2 * Q];      // W_{16Q}^{2Q} = W_8
    t.wb = tw[u0 * Q];          // W_16
    t.wc = tw[u0 * 3 * Q];      // W_16^3
    return t;
}
WS_HD void slot_index8_special(int Q, int idx[8]) {
    idx[0] = 0; idx[1] = 4 * Q; idx[2] = 2 * Q; idx[3] = 6 * Q; idx[4] = Q; idx[5] = 7 * Q; idx[6] = 3 * Q; idx[7] = 5 * Q;
}
WS_HD void radix8_special(const double2 I[8], const Tw4s& t, double2 o[8]) {
    // level s+2 for positions m, m+D, m+2D, m+3D: idx 0, 2Q (real) and Q (complex)
    double r0[4], r2[4]; double2 cQ[4];
    for (int j = 0; j < 4; j++) {
        r0[j] = I[j].x + I[j + 4].x; r2[j] = I[j].x - I[j + 4].x;
        cQ[j] = make_double2(I[j].y, -I[j + 4].y);
    }
    // level s+1 for positions m (from j = 0, 2) and m+D (from j = 1, 3): idx 0, 4Q real; 2Q, Q, 3Q complex
    double s0[2], s4[2]; double2 c2[2], c1[2], c3[2];
    for (int j = 0; j < 2; j++) {
        s0[j] = r0[j] + r0[j + 2]; s4[j] = r0[j] - r0[j + 2];
        c2[j] = make_double2(r2[j], -r2[j + 2]);
        bfly(cQ[j], cQ[j + 2], t.w8, c1[j], c3[j]);
    }
    o[0] = make_double2(s0[0] + s0[1], s0[0] - s0[1]);
    o[1] = make_double2(s4[0], -s4[1]);
    bfly(c2[0], c2[1], t.wa, o[2], o[3]);
    bfly(c1[0], c1[1], t.wb, o[4], o[5]);
    bfly(c3[0], c3[1], t.wc, o[6], o[7]);
}

// ---- chain-free pass (lower levels): one work item per (slot, position) -------------------------
// Sink: put(pos, idx, value).  12 butterflies per general item instead of the chain's 7, but every
// item is independent, so the whole CTA works on it (the lower levels are ~12% of the arithmetic).
// Items are numbered position-fastest: the lanes of a warp hold the SAME slot of consecutive
// positions, so with the odd position strides every shared-memory access of the pass is free of
// bank conflicts and the twiddles are warp-uniform (broadcast) loads; the packed slot 0 is one more
// row of items, so that the pass is a single sweep over n_out * Q items.
template <class Sink>
WS_HD void direct_pass(int tid, int nthreads, const double2* in, int in_stride, int Q, int D, int n_out,
                       const double2* tw, int N, int s_level, Sink& sink) {
    const int L = N >> s_level;
    for (int item = tid; item < n_out * Q; item += nthreads) {
        const int k = item / n_out, m = item - k * n_out;
        double2 I[8], o[8];
        int idx[8];
        for (int j = 0; j < 8; j++) I[j] = in[(m + j * D) * in_stride + k];
        if (k) {
            const Tw7 t = load_tw7(tw, N, L, Q, k);
            radix8_general(I, t, o);
            slot_index8(Q, k, idx);
        } else {
            const Tw4s ts = load_tw4s(tw, N, L, Q);
            radix8_special(I, ts, o);
            slot_index8_special(Q, idx);
        }
        for (int j = 0; j < 8; j++) sink.put(m, idx[j], o[j]);
    }
}

// ---- top pass (levels 3 -> 0, D = 1): register chains over consecutive windows ------------------
// A thread owns one general slot k of level 3 and one sub-chain of T/S consecutive windows; per
// window it reads ONE new slot and emits EIGHT bins with 7 packed butterflies.  The packed slot 0
// has no chain: its eight bins are produced per window by radix8_special (one thread per window).
//
// Compile-time N: the eight output bins of slot k are  +-k + C_J  with constant C_J, so a sink can
// address them as two moving pointers plus immediates.  Only four twiddles are kept in registers;
// the other three follow from  W^{L/4 - j} = -i conj(W^j)  (bfly_alt).  The loop is unrolled by 4 so
// the three register queues (inputs: 4 deep, level 2: 2 deep, level 1: 2 deep) rotate by renaming.
//
// Sink: bind(k, m0) once per chain (the windows follow consecutively from m0); begin(m) once per window; put<J>(value), J = slot_index8 order;
//       put0(m, idx, value) for the bins that come from the packed slot 0.
template <int J> struct SlotOfs;    // C_J in units of Q, and the sign of k
template <> struct SlotOfs<0> { static constexpr int c = 0, sgn = +1; };
template <> struct SlotOfs<1> { static constexpr int c = 8, sgn = -1; };
template <> struct SlotOfs<2> { static constexpr int c = 4, sgn = -1; };
template <> struct SlotOfs<3> { static constexpr int c = 4, sgn = +1; };
template <> struct SlotOfs<4> { static constexpr int c = 2, sgn = -1; };
template <> struct SlotOfs<5> { static constexpr int c = 6, sgn = +1; };
template <> struct SlotOfs<6> { static constexpr int c = 2, sgn = +1; };
template <> struct SlotOfs<7> { static constexpr int c = 6, sgn = -1; };

// packed butterfly with the twiddle w' = -i conj(w) = (-w.y, -w.x)
WS_HD void bfly_alt(double2 A, double2 B, double2 w, double2& P, double2& R) {
    double2 t = make_double2(w.x * B.y - w.y * B.x, -(w.x * B.x + w.y * B.y));
    P = make_double2(A.x + t.x, A.y + t.y);
    R = make_double2(A.x - t.x, t.y - A.y);
}

template <int N, int TOP = 3> struct TopGeomT {
    static constexpr int Q = N >> (TOP + 1);                          // slots of a vector at level TOP
    static constexpr int stride = Q + 1;                             // = Plan::stride[1] = top_stride(Q)
};
template <int N> using TopGeom = TopGeomT<N, 3>;

template <int N, class Sink>
WS_HD void chain_step(const double2*& nxt, double2& qold, double2& f2oP, double2& f2oR, const double2* f1,
                      double2* g, double2 w2, double2 w1a, double2 w0a, double2 w0c, int m, Sink& sink) {
    const double2 In = *nxt;
    nxt += TopGeom<N>::stride;
    double2 nP, nR;
    bfly(qold, In, w2, nP, nR);                 // F2(m+3)
    bfly(f2oP, nP, w1a, g[0], g[1]);            // F1(m+1): idx k, 4Q-k
    bfly_alt(f2oR, nR, w1a, g[2], g[3]);        //          idx 2Q-k, 2Q+k   (twiddle W^{2Q-k} = -i conj W^k)
    qold = In; f2oP = nP; f2oR = nR;            // the consumed queue heads are replaced by the newest
    double2 P, R;
    sink.begin(m);
    bfly(f1[0], g[0], w0a, P, R);     sink.template put<0>(P); sink.template put<1>(R);
    bfly_alt(f1[1], g[1], w0a, P, R); sink.template put<2>(P); sink.template put<3>(R);   // W^{4Q-k}
    bfly(f1[2], g[2], w0c, P, R);     sink.template put<4>(P); sink.template put<5>(R);   // W^{2Q-k}
    bfly_alt(f1[3], g[3], w0c, P, R); sink.template put<6>(P); sink.template put<7>(R);   // W^{2Q+k}
}

// One chain: slot k of level 3, windows m0 .. m0 + per - 1 (per a multiple of 4).  HOOK(it) runs
// after every unrolled group of four windows — also for threads with active == false, which skip
// the arithmetic but keep the control flow of their warp uniform (the overlapped kernel arrives on
// a named barrier there).
template <int N, class Sink, class Hook>
WS_HD void chain_single(bool active, int k, int m0, int per, const double2* in, const double2* tw, Sink& sink,
                        Hook hook) {
    constexpr int Q = TopGeom<N>::Q;
    constexpr int stride = TopGeom<N>::stride;
    const double2 w2 = tw[4 * k];            // W_{N/4}^k (level 2 has DFT length N/4)
    const double2 w1a = tw[2 * k];           // W_{N/2}^k
    const double2 w0a = tw[k];               // W_N^k
    const double2 w0c = tw[2 * Q - k];       // W_N^{2Q-k}
    const double2* src = in + k;
    sink.bind(k, m0);
    double2 qa, qb, qc, qd;                  // inputs I_{m+3..m+6}
    double2 eP, eR, oP, oR;                  // F2(m+1), F2(m+2)
    double2 ha[4], hb[4];                    // F1(m) / F1(m+1), alternating
    {
        const double2 I0 = src[m0 * stride], I1 = src[(m0 + 1) * stride], I2 = src[(m0 + 2) * stride];
        qa = src[(m0 + 3) * stride]; qb = src[(m0 + 4) * stride];
        qc = src[(m0 + 5) * stride]; qd = src[(m0 + 6) * stride];
        double2 zP, zR;
        bfly(I0, qb, w2, zP, zR);            // F2(m0)
        bfly(I1, qc, w2, eP, eR);            // F2(m0+1)
        bfly(I2, qd, w2, oP, oR);            // F2(m0+2)
        bfly(zP, oP, w1a, ha[0], ha[1]);     // F1(m0)
        bfly_alt(zR, oR, w1a, ha[2], ha[3]);
    }
    const double2* nxt = src + (m0 + 7) * stride;
    int it = 0;
    for (int m = m0; m < m0 + per; m += 4, it++) {
        if (active) {
            chain_step<N>(nxt, qa, eP, eR, ha, hb, w2, w1a, w0a, w0c, m, sink);
            chain_step<N>(nxt, qb, oP, oR, hb, ha, w2, w1a, w0a, w0c, m + 1, sink);
            chain_step<N>(nxt, qc, eP, eR, ha, hb, w2, w1a, w0a, w0c, m + 2, sink);
            chain_step<N>(nxt, qd, oP, oR, hb, ha, w2, w1a, w0a, w0c, m + 3, sink);
        }
        hook(it);
    }
}

// Same chain with a per-WINDOW hand-over: sink.end(m) runs after every window for every thread of the
// warp, active or not.  The staged kernel (spectra rows collected in shared memory and drained by
// bulk async stores) synchronises the warps of a segment there.
template <int N, class Sink, class Hook>
WS_HD void chain_single_stepwise(bool active, int k, int m0, int per, const double2* in, const double2* tw,
                                 Sink& sink, Hook hook) {
    constexpr int Q = TopGeom<N>::Q;
    constexpr int stride = TopGeom<N>::stride;
    const double2 w2 = tw[4 * k];
    const double2 w1a = tw[2 * k];
    const double2 w0a = tw[k];
    const double2 w0c = tw[2 * Q - k];
    const double2* src = in + k;
    sink.bind(k, m0);
    double2 qa, qb, qc, qd;
    double2 eP, eR, oP, oR;
    double2 ha[4], hb[4];
    {
        const double2 I0 = src[m0 * stride], I1 = src[(m0 + 1) * stride], I2 = src[(m0 + 2) * stride];
        qa = src[(m0 + 3) * stride]; qb = src[(m0 + 4) * stride];
        qc = src[(m0 + 5) * stride]; qd = src[(m0 + 6) * stride];
        double2 zP, zR;
        bfly(I0, qb, w2, zP, zR);
        bfly(I1, qc, w2, eP, eR);
        bfly(I2, qd, w2, oP, oR);
        bfly(zP, oP, w1a, ha[0], ha[1]);
        bfly_alt(zR, oR, w1a, ha[2], ha[3]);
    }
    const double2* nxt = src + (m0 + 7) * stride;
    int it = 0;
    for (int m = m0; m < m0 + per; m += 4, it++) {
        if (active) chain_step<N>(nxt, qa, eP, eR, ha, hb, w2, w1a, w0a, w0c, m, sink);
        sink.end(m);
        if (active) chain_step<N>(nxt, qb, oP, oR, hb, ha, w2, w1a, w0a, w0c, m + 1, sink);
        sink.end(m + 1);
        if (active) chain_step<N>(nxt, qc, eP, eR, ha, hb, w2, w1a, w0a, w0c, m + 2, sink);
        sink.end(m + 2);
        if (active) chain_step<N>(nxt, qd, oP, oR, hb, ha, w2, w1a, w0a, w0c, m + 3, sink);
        sink.end(m + 3);
        hook(it);
    }
}

// bins that come from the packed slot 0 of level 3 (multiples of Q): one window per thread
template <int N, class Sink>
WS_HD void special_pass(int tid, int nthreads, const double2* in, int T, const double2* tw, Sink& sink) {
    constexpr int Q = TopGeom<N>::Q;
    constexpr int stride = TopGeom<N>::stride;
    const Tw4s ts = load_tw4s(tw, N, N, Q);
    for (int m = tid; m < T; m += nthreads) {
        double2 I[8], o[8];
        int idx[8];
        for (int j = 0; j < 8; j++) I[j] = in[(m + j) * stride];
        radix8_special(I, ts, o);
        slot_index8_special(Q, idx);
        for (int j = 0; j < 8; j++) sink.put0(m, idx[j], o[j]);
    }
}

template <int N, class Sink>
WS_HD void chain_pass(int tid, int nthreads, const double2* in, int T, int S, const double2* tw, Sink& sink) {
    constexpr int gen = TopGeom<N>::Q - 1;
    const int per = T / S;                       // multiple of 4 (plan_make)
    for (int c = tid; c < S * gen; c += nthreads) {
        const int sub = c / gen, k = 1 + c - sub * gen;
        chain_single<N>(true, k, sub * per, per, in, tw, sink, [](int) {});
    }
    special_pass<N>(tid, nthreads, in, T, tw, sink);
}

// ---- radix-4 top pass (levels 2 -> 0): half the register state of the radix-8 chain -------------
// A thread owns one general slot k of level 2 (Q = N/8 slots) and a sub-chain of consecutive windows:
// per window ONE new slot in, FOUR bins out (k, 4Q-k, 2Q-k, 2Q+k) with 3 packed butterflies; two
// twiddles in registers (W_{N/2}^k, W_N^k; W_N^{2Q-k} = -i conj W_N^k).  Queues: inputs 2 deep,
// level 1 one deep -> period 2, unrolled by 4.
// Sink: as for chain_pass, with put<J>, J = 0..3 in the order k, 4Q-k, 2Q-k, 2Q+k; the packed slot 0
// yields bins 0, Q, 2Q, 3Q per window through put0.
template <int J> struct SlotOfs4;
template <> struct SlotOfs4<0> { static constexpr int c = 0, sgn = +1; };
template <> struct SlotOfs4<1> { static constexpr int c = 4, sgn = -1; };
template <> struct SlotOfs4<2> { static constexpr int c = 2, sgn = -1; };
template <> struct SlotOfs4<3> { static constexpr int c = 2, sgn = +1; };

template <int N, class Sink>
WS_HD void chain4_step(const double2*& nxt, double2& qold, const double2* f1, double2* g, double2 w1,
                       double2 w0, int m, Sink& sink) {
    const double2 In = *nxt;
    nxt += TopGeomT<N, 2>::stride;
    bfly(qold, In, w1, g[0], g[1]);              // F1(m+1): idx k, 2Q-k
    qold = In;
    double2 P, R;
    sink.begin(m);
    bfly(f1[0], g[0], w0, P, R);     sink.template put<0>(P); sink.template put<1>(R);
    bfly_alt(f1[1], g[1], w0, P, R); sink.template put<2>(P); sink.template put<3>(R);   // W^{2Q-k}
}

template <int N, class Sink>
WS_HD void chain_pass4(int tid, int nthreads, const double2* in, int T, int S, const double2* tw, Sink& sink) {
    constexpr int Q = TopGeomT<N, 2>::Q;
    constexpr int stride = TopGeomT<N, 2>::stride;
    constexpr int gen = Q - 1;
    const int per = T / S;
    for (int c = tid; c < S * gen; c += nthreads) {
        const int sub = c / gen, k = 1 + c - sub * gen;
        const double2 w1 = tw[2 * k];            // W_{N/2}^k
        const double2 w0 = tw[k];                // W_N^k
        const double2* src = in + k;
        const int m0 = sub * per;
        sink.bind(k, m0);
        double2 qa, qb;                          // F2(m+1), F2(m+2)
        double2 ha[2], hb[2];                    // F1(m) / F1(m+1), alternating
        {
            const double2 I0 = src[m0 * stride];
            qa = src[(m0 + 1) * stride]; qb = src[(m0 + 2) * stride];
            bfly(I0, qb, w1, ha[0], ha[1]);      // F1(m0)
        }
        const double2* nxt = src + (m0 + 3) * stride;
        for (int m = m0; m < m0 + per; m += 4) {
            chain4_step<N>(nxt, qa, ha, hb, w1, w0, m, sink);
            chain4_step<N>(nxt, qb, hb, ha, w1, w0, m + 1, sink);
            chain4_step<N>(nxt, qa, ha, hb, w1, w0, m + 2, sink);
            chain4_step<N>(nxt, qb, hb, ha, w1, w0, m + 3, sink);
        }
    }
    // packed slot 0 of level 2: (F2[0], F2[Q]) real -> bins 0, Q, 2Q, 3Q of every window
    const double2 w8 = tw[Q];                    // W_N^Q = W_8
    for (int m = tid; m < T; m += nthreads) {
        const double2 I0 = in[m * stride], I1 = in[(m + 1) * stride], I2 = in[(m + 2) * stride], I3 = in[(m + 3) * stride];
        // level 1 at positions m (I0, I2) and m+1 (I1, I3): idx 0, 2Q real; idx Q complex
        const double a0 = I0.x + I2.x, a2 = I0.x - I2.x; const double2 aQ = make_double2(I0.y, -I2.y);
        const double b0 = I1.x + I3.x, b2 = I1.x - I3.x; const double2 bQ = make_double2(I1.y, -I3.y);
        sink.put0(m, 0, make_double2(a0 + b0, a0 - b0));      // (F[0], F[N/2]); the sink drops the Nyquist
        sink.put0(m, 2 * Q, make_double2(a2, -b2));
        double2 P, R;
        bfly(aQ, bQ, w8, P, R);
        sink.put0(m, Q, P);
        sink.put0(m, 3 * Q, R);
    }
}

// bottom level for all positions: one position per thread step.  A kernel templated on N passes
// the stored-level count and the DFT length at compile time (NST, LB): no run-time indexing of the
// Plan arrays, one SmallRdft instance instead of four.
template <int NST = 0, int LB = 0>
WS_HD void bottom_level(int tid, int nthreads, const double* x, const Plan& pl, const double2* tw,
                        double2* arena) {
    const int i = NST ? NST : pl.nst;
    const int step = 1 << pl.lev[i];
    double2* out = arena + pl.off[i];
    for (int p = tid; p < pl.P[i]; p += nthreads) {
        double2* o = out + (size_t)p * pl.stride[i];
        if (LB) { SmallRdft<(LB ? LB : 2)>::run(x + p, step, tw, pl.N, o); continue; }
        switch (pl.Lb) {
            case 2: SmallRdft<2>::run(x + p, step, tw, pl.N, o); break;
            case 4: SmallRdft<4>::run(x + p, step, tw, pl.N, o); break;
            case 8: SmallRdft<8>::run(x + p, step, tw, pl.N, o); break;
            default: SmallRdft<16>::run(x + p, step, tw, pl.N, o); break;
        }
    }
}

// stored levels and bottom DFT length of a window length, as plan_make derives them for top = 3
template <int N, int TOP = 3> struct LevelsOf {
    static constexpr int nst = (N >> TOP) <= 16 ? 1 : ((N >> (TOP + 3)) <= 16 ? 2 : 3);
    static constexpr int Lb = N >> (TOP + 3 * (nst - 1));
};

struct SmemSink {
    double2* base; int stride;
    WS_HD void put(int pos, int idx, double2 v) { base[(size_t)pos * stride + idx] = v; }
};

}  // namespace ws_slide
