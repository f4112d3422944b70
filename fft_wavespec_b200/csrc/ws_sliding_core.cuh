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
// Three levels are fused per pass (radix-8 in registers): a thread owns slot k of level s+3 and
// one residue r of the positions modulo D = 2^s, walks the chain m = r, r+D, r+2D, ... and keeps
// the intermediate levels in registers, reading ONE new slot and producing the EIGHT slots of
// F(s, m) per step: 7 packed butterflies per step for a general slot, 4 for the packed slot 0.
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

// Geometry of one tile.  Level indices: 0 = full window transform; `nst` fused passes cover levels
// 3*nst -> 0; the bottom level sb = 3*nst is computed directly from the samples (Lb-point DFT).
struct Plan {
    int N, log2N;
    int T;            // windows per tile
    int S;            // sub-chains of the top pass (T % S == 0)
    int nst;          // fused passes: 2 (N <= 1024) or 3
    int sb;           // bottom level = 3*nst
    int Lb;           // bottom DFT length = N >> sb  (2..16)
    int Q[4];         // slots per vector at level 3*i : Q[i] = N >> (3*i + 1)   (Q[0] = N/2)
    int P[4];         // positions held at level 3*i   : P[i] = T + (1 << 3*i) - 1
    int stride[4];    // slots between consecutive positions in shared memory (padded)
    int off[4];       // slot offset of level 3*i's array inside the vector arena (i >= 1)
    int x_len;        // staged samples: T + N - 1
    int arena_slots;  // total double2 slots of the vector arena
};

WS_HD bool plan_make(Plan& pl, int N, int T, int S) {
    pl.N = N;
    int l = 0;
    while ((1 << l) < N) l++;
    pl.log2N = l;
    if ((1 << l) != N || N < 256 || N > 4096) return false;
    pl.T = T; pl.S = S;
    if (S < 1 || T % S) return false;
    pl.nst = (N <= 1024) ? 2 : 3;
    pl.sb = 3 * pl.nst;
    pl.Lb = N >> pl.sb;
    for (int i = 0; i <= pl.nst; i++) {
        pl.Q[i] = N >> (3 * i + 1);
        pl.P[i] = T + (1 << (3 * i)) - 1;
        // pad so that the 8 lanes of a quarter warp never share a bank group when they write
        // different positions (a writer pass has Q_in = q/8 slots per residue)
        int q = pl.Q[i];
        pl.stride[i] = (i == pl.nst) ? q + 1 : (q >= 64 ? q : q + (q >> 3 ? (q >> 3) : 1));
    }
    // arena order: deepest level first, level 3 (read by the top pass) last, so that the
    // epilogue buffers can reuse everything below level 3 while the top pass still reads it
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

// ---- one fused pass (levels s+3 -> s) ---------------------------------------------------------
// Sink: receives slot `idx` of F(s, position pos).
//
// in       : level s+3 vectors, `in_stride` slots apart, Q slots each
// D        : 2^s (chain stride in positions)
// n_out    : number of level-s positions to produce (positions 0 .. n_out-1, tile relative)
// S        : sub-chains per residue (top pass only; 1 otherwise).  Sub-chain c covers the
//            ceil-split range of chain steps.
template <class Sink>
WS_HD void fused_pass(int tid, int nthreads, const double2* in, int in_stride, int Q, int D, int n_out,
                      int S, const double2* tw, int N, int s_level, Sink& sink) {
    const int L = N >> s_level;                   // DFT length at the output level (= 16 Q)
    const int chains = S * D * Q;
    for (int c = tid; c < chains; c += nthreads) {
        const int k = c % Q;
        const int r = (c / Q) % D;
        const int sub = c / (Q * D);
        // chain steps for residue r: positions r, r+D, ... < n_out
        const int steps_total = (n_out - r + D - 1) / D;
        if (steps_total <= 0) continue;
        const int per = (steps_total + S - 1) / S;
        const int i0 = sub * per;
        int i1 = i0 + per;
        if (i1 > steps_total) i1 = steps_total;
        if (i0 >= i1) continue;
        const int m0 = r + i0 * D;
        const double2* src = in + k;
#define WS_IN(pos) src[(size_t)(pos) * in_stride]
        if (k != 0) {
            // twiddles: level s+2 has length L/4, level s+1 L/2, level s L
            const int u2 = N / (L / 4), u1 = N / (L / 2), u0 = N / L;
            const double2 w2 = tw[u2 * k];
            const double2 w1a = tw[u1 * k], w1b = tw[u1 * (2 * Q - k)];
            const double2 w0a = tw[u0 * k], w0b = tw[u0 * (4 * Q - k)], w0c = tw[u0 * (2 * Q - k)],
                          w0d = tw[u0 * (2 * Q + k)];
            double2 q0, q1, q2, q3;               // I_{3..6} relative to the current position
            double2 f2aP, f2aR, f2bP, f2bR;       // F2(m+D), F2(m+2D)
            double2 f1[4];                        // F1(m): idx k, 4Q-k, 2Q-k, 2Q+k
            {
                double2 I0 = WS_IN(m0), I1 = WS_IN(m0 + D), I2 = WS_IN(m0 + 2 * D);
                q0 = WS_IN(m0 + 3 * D);
                q1 = WS_IN(m0 + 4 * D); q2 = WS_IN(m0 + 5 * D); q3 = WS_IN(m0 + 6 * D);
                double2 f20P, f20R;
                bfly(I0, q1, w2, f20P, f20R);     // F2(m0)
                bfly(I1, q2, w2, f2aP, f2aR);     // F2(m0+D)
                bfly(I2, q3, w2, f2bP, f2bR);     // F2(m0+2D)
                bfly(f20P, f2bP, w1a, f1[0], f1[1]);
                bfly(f20R, f2bR, w1b, f1[2], f1[3]);
            }
            for (int i = i0; i < i1; i++) {
                const int m = r + i * D;
                double2 In = WS_IN(m + 7 * D);
                double2 nP, nR;
                bfly(q0, In, w2, nP, nR);         // F2(m+3D)
                double2 g[4];
                bfly(f2aP, nP, w1a, g[0], g[1]);  // F1(m+D)
                bfly(f2aR, nR, w1b, g[2], g[3]);
                double2 P, R;
                bfly(f1[0], g[0], w0a, P, R); sink.put(m, k, P);         sink.put(m, 8 * Q - k, R);
                bfly(f1[1], g[1], w0b, P, R); sink.put(m, 4 * Q - k, P); sink.put(m, 4 * Q + k, R);
                bfly(f1[2], g[2], w0c, P, R); sink.put(m, 2 * Q - k, P); sink.put(m, 6 * Q + k, R);
                bfly(f1[3], g[3], w0d, P, R); sink.put(m, 2 * Q + k, P); sink.put(m, 6 * Q - k, R);
                f1[0] = g[0]; f1[1] = g[1]; f1[2] = g[2]; f1[3] = g[3];
                f2aP = f2bP; f2aR = f2bR; f2bP = nP; f2bR = nR;
                q0 = q1; q1 = q2; q2 = q3; q3 = In;
            }
        } else {
            // packed slot 0: (F[0], F[Q']) both real, Q' = Q slots of the input level
            const int u1 = N / (L / 2), u0 = N / L;
            const double2 w8 = tw[u1 * Q];                          // W_{8Q}^{Q}  = W_8
            const double2 wa = tw[u0 * 2 * Q];                      // W_{16Q}^{2Q} = W_8
            const double2 wb = tw[u0 * Q], wc = tw[u0 * 3 * Q];     // W_16, W_16^3
            struct F2s { double r0, r2; double2 cQ; };              // idx 0, 2Q (real), Q (complex)
            struct F1s { double s0, s4; double2 c2, c1, c3; };      // idx 0, 4Q (real), 2Q, Q, 3Q
            auto mk2 = [](double2 A, double2 B) {
                F2s f; f.r0 = A.x + B.x; f.r2 = A.x - B.x; f.cQ = make_double2(A.y, -B.y); return f;
            };
            auto mk1 = [&](const F2s& A, const F2s& B) {
                F1s f; f.s0 = A.r0 + B.r0; f.s4 = A.r0 - B.r0; f.c2 = make_double2(A.r2, -B.r2);
                bfly(A.cQ, B.cQ, w8, f.c1, f.c3);
                return f;
            };
            double2 q0, q1, q2, q3;
            F2s f2a, f2b;
            F1s f1;
            {
                double2 I0 = WS_IN(m0), I1 = WS_IN(m0 + D), I2 = WS_IN(m0 + 2 * D);
                q0 = WS_IN(m0 + 3 * D);
                q1 = WS_IN(m0 + 4 * D); q2 = WS_IN(m0 + 5 * D); q3 = WS_IN(m0 + 6 * D);
                F2s f20 = mk2(I0, q1);
                f2a = mk2(I1, q2);
                f2b = mk2(I2, q3);
                f1 = mk1(f20, f2b);
            }
            for (int i = i0; i < i1; i++) {
                const int m = r + i * D;
                double2 In = WS_IN(m + 7 * D);
                F2s n2 = mk2(q0, In);
                F1s g = mk1(f2a, n2);
                sink.put(m, 0, make_double2(f1.s0 + g.s0, f1.s0 - g.s0));
                sink.put(m, 4 * Q, make_double2(f1.s4, -g.s4));
                double2 P, R;
                bfly(f1.c2, g.c2, wa, P, R); sink.put(m, 2 * Q, P); sink.put(m, 6 * Q, R);
                bfly(f1.c1, g.c1, wb, P, R); sink.put(m, Q, P);     sink.put(m, 7 * Q, R);
                bfly(f1.c3, g.c3, wc, P, R); sink.put(m, 3 * Q, P); sink.put(m, 5 * Q, R);
                f1 = g; f2a = f2b; f2b = n2;
                q0 = q1; q1 = q2; q2 = q3; q3 = In;
            }
        }
#undef WS_IN
    }
}

// bottom level for all positions: one position per thread step
WS_HD void bottom_level(int tid, int nthreads, const double* x, const Plan& pl, const double2* tw,
                        double2* arena) {
    const int i = pl.nst;
    const int step = 1 << pl.sb;
    double2* out = arena + pl.off[i];
    for (int p = tid; p < pl.P[i]; p += nthreads) {
        double2* o = out + (size_t)p * pl.stride[i];
        switch (pl.Lb) {
            case 2: SmallRdft<2>::run(x + p, step, tw, pl.N, o); break;
            case 4: SmallRdft<4>::run(x + p, step, tw, pl.N, o); break;
            case 8: SmallRdft<8>::run(x + p, step, tw, pl.N, o); break;
            default: SmallRdft<16>::run(x + p, step, tw, pl.N, o); break;
        }
    }
}

struct SmemSink {
    double2* base; int stride;
    WS_HD void put(int pos, int idx, double2 v) { base[(size_t)pos * stride + idx] = v; }
};

}  // namespace ws_slide
