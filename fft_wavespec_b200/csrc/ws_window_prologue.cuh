// ws_window_prologue.cuh — per-window prologue pieces shared by the two per-window FFT kernels
// (ws_window_fft.cu: one CTA per group of windows; ws_window_fft_warp.cu: one warp per window).
#pragma once
#include "ws_common.cuh"

namespace ws {

// prologue value of sample n of a window starting at tile offset `off`
struct Prologue {
    const double* tile;     // shared: x (or x - y for the IIR detrend)
    const double* wtab;     // global window table or nullptr
    const double* apow;     // global alpha^j or nullptr
    double sub;             // mean (DETREND_MEAN) or delta_w (DETREND_IIR)
    int mode;
    __device__ __forceinline__ double operator()(int off, int n) const {
        double v = tile[off + n];
        if (mode == 2) v = v - sub;
        else if (mode == 1) v = v - __ldg(apow + n) * sub;
        if (wtab) v = v * __ldg(wtab + n);
        return v;
    }
};

// Same arithmetic with the detrend mode and the presence of a window table fixed at compile time,
// on the sample PAIR (2m, 2m+1) that forms one complex input: the tables are read as one 16-byte
// load per pair (they are cudaMalloc'ed, so element 2m is 16-byte aligned).
template <int MODE, bool WIN>
struct ProloguePair {
    const double* tile;
    const double2* wtab2;
    const double2* apow2;
    double sub;
    __device__ __forceinline__ double2 operator()(int off, int m) const {
        double2 v = make_double2(tile[off + 2 * m], tile[off + 2 * m + 1]);
        if (MODE == 2) { v.x = v.x - sub; v.y = v.y - sub; }
        else if (MODE == 1) { const double2 a = __ldg(apow2 + m); v.x = v.x - a.x * sub; v.y = v.y - a.y * sub; }
        if (WIN) { const double2 w = __ldg(wtab2 + m); v.x = v.x * w.x; v.y = v.y * w.y; }
        return v;
    }
};

// Trend IIR of Legacy/...-kalman-fast.mq5:3367-3379 over x[0..L): y[0] = c (x0 + x0),
// y[a] = c (x[a] + x[a-1]) + alpha y[a-1].  Blocked over the CTA: local recurrences from zero,
// a serial carry pass over the kThreads chunk ends, then the alpha^k fix-up.  Differs from the
// serial loop by rounding only (a few ulp of y).
template <int kThreads>
__device__ void cta_trend_iir(const double* x, int L, double al, double c, double* y, double* carry) {
    const int tid = threadIdx.x;
    const int chunk = (L + kThreads - 1) / kThreads;
    const int a0 = tid * chunk;
    double acc = 0.0;
    for (int a = a0; a < a0 + chunk && a < L; a++) {
        double u = (a == 0) ? c * (x[0] + x[0]) : c * (x[a] + x[a - 1]);
        acc = u + al * acc;
        y[a] = acc;
    }
    carry[tid] = acc;
    __syncthreads();
    if (tid == 0) {
        double ac = pow(al, (double)chunk);
        double run = 0.0;
        for (int t = 0; t < kThreads; t++) {
            double e = carry[t];
            carry[t] = run;                 // carry-in of chunk t
            run = e + ac * run;
        }
    }
    __syncthreads();
    double cin = carry[tid];
    double f = al;
    for (int a = a0; a < a0 + chunk && a < L; a++) { y[a] = y[a] + f * cin; f *= al; }
    __syncthreads();
}

}  // namespace ws
