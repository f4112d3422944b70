// ws_window_fft.cu — generic per-window real FFT with fused prologue and epilogue (sm_100a).
//
// Used whenever a window needs its own transform: any detrend / window function / PLA feed
// (SURVEY.md section 8a rows A2a, A2b, A3, A11), hop > 1, single-window calls
// (gpu_fft_real_forward, gpu_extract_cycles) and contiguous batches.  The plain hop-1
// rectangular case is served by the shared-butterfly kernel in ws_sliding.cu instead.
//
// One CTA owns a tile of consecutive windows of one series.  The tile's samples are staged in
// shared memory once (each sample is read from HBM once per tile, not once per window).  Each
// window becomes an N/2-point complex Stockham FFT (radix-8 passes in registers + one radix-4 or
// radix-2 pass for the remaining bits) of z[m] = v[2m] + i v[2m+1] exchanged through shared memory, followed by the
// real-input split, the power spectrum and a warp-per-window top-K epilogue.
//
// The transform uses exact table twiddles, not the reference's multiplicative recurrence
// (Legacy/WaveSpecZZ_1.0.2.mq5:955-970): results agree to ~1e-14 relative, inside the 1e-9 bar.
#include <cstdio>
#include <cstdlib>
#include "ws_common.cuh"
#include "ws_epilogue.cuh"
#include "ws_window_prologue.cuh"
#include "ws_series.h"

namespace ws {

// CTA size: 256 threads; up to 8 windows in flight per CTA below N = 1024, 4 from there on

__device__ __forceinline__ void r4_butterfly(double2& a0, double2& a1, double2& a2, double2& a3) {
    double2 b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3);
    double2 d = csub(a1, a3);
    double2 b3 = make_double2(d.y, -d.x);   // -i * (a1 - a3)
    a0 = cadd(b0, b2); a2 = csub(b0, b2); a1 = cadd(b1, b3); a3 = csub(b1, b3);
}

// forward 8-point DFT in registers (decimation in frequency; outputs in natural order)
__device__ __forceinline__ void r8_butterfly(double2* a) {
    const double h = 0.70710678118654752440;
    double2 b0 = cadd(a[0], a[4]), b1 = cadd(a[1], a[5]), b2 = cadd(a[2], a[6]), b3 = cadd(a[3], a[7]);
    double2 d0 = csub(a[0], a[4]), t1 = csub(a[1], a[5]), t2 = csub(a[2], a[6]), t3 = csub(a[3], a[7]);
    double2 d1 = make_double2((t1.x + t1.y) * h, (t1.y - t1.x) * h);        // * W8
    double2 d2 = make_double2(t2.y, -t2.x);                                   // * -i
    double2 d3 = make_double2((t3.y - t3.x) * h, -(t3.x + t3.y) * h);       // * W8^3
    double2 c0 = cadd(b0, b2), c1 = cadd(b1, b3), c2 = csub(b0, b2), u = csub(b1, b3);
    double2 c3 = make_double2(u.y, -u.x);
    double2 e0 = cadd(d0, d2), e1 = cadd(d1, d3), e2 = csub(d0, d2), v = csub(d1, d3);
    double2 e3 = make_double2(v.y, -v.x);
    a[0] = cadd(c0, c1); a[4] = csub(c0, c1); a[2] = cadd(c2, c3); a[6] = csub(c2, c3);
    a[1] = cadd(e0, e1); a[5] = csub(e0, e1); a[3] = cadd(e2, e3); a[7] = csub(e2, e3);
}

// One Stockham pass of radix R over nw windows of M points: thread item j reads in[j + r M/R],
// multiplies by W^{r k} (k = j mod Ns), does the R-point DFT and writes out[(j-k) R + k + r Ns].
// LOAD(wl, m) supplies element m of window wl (the first pass reads the staged samples through
// the prologue, the others the previous pass's buffer).
template <int R, int kThreads, class Load>
__device__ __forceinline__ void stockham_pass(Load load, double2* out, int nw, int M, int N, int Ns,
                                              const double2* __restrict__ tw) {
    const int q = M / R;
    const int tstep = N / (R * Ns);
    for (int idx = threadIdx.x; idx < nw * q; idx += kThreads) {
        const int wl = idx / q, j = idx - wl * q;
        const int k = j & (Ns - 1);
        double2 a[R];
#pragma unroll
        for (int r = 0; r < R; r++) a[r] = load(wl, j + r * q);
        if (Ns > 1) {
            const int t = tstep * k;
            if (R == 8) {
                // three table loads; the other four twiddles are products (FP64 is idle here, the
                // loads are what stalls: ncu long-scoreboard)
                const double2 w1 = __ldg(tw + t), w2 = __ldg(tw + 2 * t), w4 = __ldg(tw + 4 * t);
                const double2 w3 = cmul(w1, w2), w5 = cmul(w1, w4), w6 = cmul(w2, w4), w7 = cmul(w3, w4);
                a[1] = cmul(a[1], w1); a[2] = cmul(a[2], w2); a[3] = cmul(a[3], w3); a[4] = cmul(a[4], w4);
                a[5] = cmul(a[5], w5); a[6] = cmul(a[6], w6); a[7] = cmul(a[7], w7);
            } else {
#pragma unroll
                for (int r = 1; r < R; r++) a[r] = cmul(a[r], __ldg(tw + r * t));
            }
        }
        if (R == 8) r8_butterfly(a);
        else if (R == 4) r4_butterfly(a[0], a[1], a[2], a[3]);
        else { double2 x0 = a[0], x1 = a[1]; a[0] = cadd(x0, x1); a[1] = csub(x0, x1); }
        double2* o = out + (size_t)wl * M + (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; r++) o[r * Ns] = a[r];
    }
}

// LN = log2 of the window length baked in at compile time (0: run-time N); CAP likewise for the
// concurrent-window cap.  With LN fixed every index split below is a shift and the pass loop
// unrolls; the run-time variant pays an integer division per element (ncu: 4x the FP64 work).
template <int kThreads, int LN, int CAP>
__global__ void __launch_bounds__(kThreads)
window_fft_kernel(const Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = LN ? (1 << LN) : p.N, M = N >> 1;
    const int log2N = LN ? LN : p.log2N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = p.tile_windows;
    const int64_t w0 = p.win_offset + (int64_t)blockIdx.x * T;
    const int s = blockIdx.y;
    const int64_t nwin = p.nwin;
    const int64_t wend = p.win_offset + p.chunk_nwin;
    const int tw_count = (int)((w0 + T <= wend) ? T : (wend - w0));
    if (tw_count <= 0) return;
    const int Lt = (tw_count - 1) * p.hop + N;

    // windows processed concurrently by the CTA: as many as keep every thread on one radix-8
    // butterfly per pass (M/8 butterflies per window), at most 4
    const int q = M >> 3;                          // butterflies per window per radix-8 pass
    int wpc = q > 0 ? kThreads / q : kThreads;
    if (wpc < 1) wpc = 1;
    const int cap = CAP ? CAP : p.k1_wpc_cap;
    if (wpc > cap) wpc = cap;

    // shared layout: tile[Lt_max] | delta[T] | bufA[wpc*M] | bufB[wpc*M] | ord[wpc*M] ints
    const bool from_feed = p.feed != nullptr;
    const int Lt_max = from_feed ? wpc * N : (T - 1) * p.hop + N;
    double* tile = reinterpret_cast<double*>(smem_raw);
    double* delta = tile + ((Lt_max + 1) & ~1);
    double2* bufA = reinterpret_cast<double2*>(delta + ((T + 1) & ~1));
    double2* bufB = bufA + (size_t)wpc * M;
    int* ord = reinterpret_cast<int*>(bufB + (size_t)wpc * M);
    double* scr = reinterpret_cast<double*>(ord + (p.select == 1 ? (((size_t)wpc * M + 1) & ~(size_t)1) : 0));  // IIR scratch

    const int pro_mode = (from_feed && p.detrend == 1) ? 0 : p.detrend;
    const double* src = p.series + (int64_t)s * p.series_stride + w0 * p.hop;

    if (!from_feed) {
        for (int i = tid; i < Lt; i += kThreads) tile[i] = src[i];
        __syncthreads();
        if (p.detrend == 1) {
            // Trend IIR (Legacy/...-kalman-fast.mq5:3367-3379) restarted per window:
            //   tr_w[j] = y[w+j] + alpha^j * (2c x[w] - y[w])
            // for ANY y obeying y[a] = c (x[a] + x[a-1]) + alpha y[a-1] on the tile, so y is run
            // once per tile (blocked: local recurrences, serial carry pass, fix-up) and each
            // window only needs delta_w = 2c x[w] - y[w].
            double* y = scr;
            double* carry = y + Lt_max;
            const double c = p.iir_c;
            cta_trend_iir<kThreads>(tile, Lt, p.iir_alpha, c, y, carry);
            for (int t = tid; t < tw_count; t += kThreads) {
                int a = t * p.hop;
                delta[t] = c * (tile[a] + tile[a]) - y[a];
            }
            __syncthreads();
            for (int i = tid; i < Lt; i += kThreads) tile[i] = tile[i] - y[i];
            __syncthreads();
        }
    }


    for (int wb = 0; wb < tw_count; wb += wpc) {
        const int nw = (tw_count - wb) < wpc ? (tw_count - wb) : wpc;

        if (from_feed) {
            // pre-built feed (PLA lines): one window per row, no overlap to exploit
            for (int i = tid; i < nw * N; i += kThreads) {
                int wl = i / N, n = i - wl * N;
                tile[wl * N + n] = p.feed[((int64_t)s * p.chunk_nwin + (w0 - p.win_offset) + wb + wl) * N + n];
            }
            __syncthreads();
            if (p.detrend == 1) {
                // per-window restart on a pre-built feed: no tile-level sharing possible
                double* y = scr;
                double* carry = y + Lt_max;
                for (int wl = 0; wl < nw; wl++) {
                    cta_trend_iir<kThreads>(tile + wl * N, N, p.iir_alpha, p.iir_c, y, carry);
                    for (int i = tid; i < N; i += kThreads) tile[wl * N + i] = tile[wl * N + i] - y[i];
                    __syncthreads();
                }
            }
        }
        if (p.detrend == 2) {
            // mean removal (Legacy/WaveSpecZZ_gpu_wip.mq5:940-943); one warp per window
            for (int wl = warp; wl < nw; wl += kThreads / 32) {
                int off = from_feed ? wl * N : (wb + wl) * p.hop;
                double sum = 0.0;
                for (int n = lane; n < N; n += 32) sum += tile[off + n];
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) sum += shfl_xor_d(sum, m);
                if (lane == 0) delta[wb + wl] = sum / (double)N;
            }
            __syncthreads();
        }

        // ---- mixed-radix Stockham passes: radix 8 while three bits remain, then one radix 4 / 2
        double2* in = bufA;
        double2* out = bufB;
        {
            auto from_tile = [&](int wl, int m) {
                const int off = from_feed ? wl * N : (wb + wl) * p.hop;
                Prologue pr{tile, p.has_window ? p.wtab : nullptr, p.apow,
                            pro_mode ? delta[wb + wl] : 0.0, pro_mode};
                return make_double2(pr(off, 2 * m), pr(off, 2 * m + 1));
            };
            int rem = log2N - 1;                     // log2 M
            int Ns = 1;
            bool first = true;
            if (rem == 0) {                          // M = 1 (N = 2): nothing to transform
                for (int wl = tid; wl < nw; wl += kThreads) out[wl] = from_tile(wl, 0);
                __syncthreads();
                double2* t = in; in = out; out = t;
            }
            constexpr int kUnroll = LN ? 5 : 1;      // 5 passes cover log2 M <= 13
#pragma unroll kUnroll
            for (int pass = 0; pass < 5; pass++) {
                if (rem <= 0) break;
                const int lr = rem >= 3 ? 3 : rem;
                const double2* src = in;
                auto from_buf = [&](int wl, int m) { return src[(size_t)wl * M + m]; };
                if (first) {
                    if (lr == 3) stockham_pass<8, kThreads>(from_tile, out, nw, M, N, Ns, p.tw);
                    else if (lr == 2) stockham_pass<4, kThreads>(from_tile, out, nw, M, N, Ns, p.tw);
                    else stockham_pass<2, kThreads>(from_tile, out, nw, M, N, Ns, p.tw);
                } else {
                    if (lr == 3) stockham_pass<8, kThreads>(from_buf, out, nw, M, N, Ns, p.tw);
                    else if (lr == 2) stockham_pass<4, kThreads>(from_buf, out, nw, M, N, Ns, p.tw);
                    else stockham_pass<2, kThreads>(from_buf, out, nw, M, N, Ns, p.tw);
                }
                __syncthreads();
                double2* t = in; in = out; out = t;
                Ns <<= lr; rem -= lr; first = false;
            }
        }

        // ---- real-input split: X[k] = E + W_N^k O ; power ; spectra store
        // `in` holds Z; X goes to `out`, powers reuse the (dead) Z buffer afterwards.
        for (int idx = tid; idx < nw * M; idx += kThreads) {
            int wl = idx / M, k = idx - wl * M;
            const double2* Z = in + (size_t)wl * M;
            double2 zk = Z[k];
            double2 zm = cconj(Z[(M - k) & (M - 1)]);
            double2 E = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y + zm.y));
            double2 D = csub(zk, zm);
            double2 O = make_double2(0.5 * D.y, -0.5 * D.x);       // -i/2 * D
            double2 X = cadd(E, cmul(__ldg(p.tw + k), O));
            out[(size_t)wl * M + k] = X;
            if (p.spectra) {
                double2* g = reinterpret_cast<double2*>(
                    p.spectra + (((int64_t)s * p.spec_nwin + (w0 - p.spec_w0) + wb + wl)) * N);
                g[k] = X;
            }
            if (p.band_buf && k >= p.band_lo && k <= p.band_hi)      // hand-off to the tracker / rows kernels
                p.band_buf[((int64_t)s * p.chunk_nwin + (w0 - p.win_offset) + wb + wl) * (p.band_hi - p.band_lo + 1) +
                           (k - p.band_lo)] = X;
        }
        __syncthreads();
        double* pw = reinterpret_cast<double*>(in);    // Z is dead: wpc*M double2 = room for M doubles per window
        for (int idx = tid; idx < nw * M; idx += kThreads) {
            int wl = idx / M, k = idx - wl * M;
            double2 X = out[(size_t)wl * M + k];
            pw[(size_t)wl * M + k] = X.x * X.x + X.y * X.y;
        }
        __syncthreads();

        const bool want_sel = p.bins || p.rows || p.waves || p.contrib;
        if (want_sel) {
            for (int wl = warp; wl < nw; wl += kThreads / 32)
                warp_select_emit(p, pw + (size_t)wl * M, out + (size_t)wl * M, ord + (size_t)wl * M,
                                 (int64_t)s * nwin + w0 + wb + wl);
        }
        if (p.phase) {
            // A6 phase chain (Legacy/...-kalman-fast.mq5:1183-1263) over bins 0..M-1.
            // phase: parallel atan2; unwrap: serial prefix per window (one lane), as the
            // reference accumulates it; group delay: parallel differences.
            double* ph = pw;                      // overwrite powers (selection is done below the barrier)
            __syncthreads();
            for (int idx = tid; idx < nw * M; idx += kThreads) {
                int wl = idx / M, k = idx - wl * M;
                double2 X = out[(size_t)wl * M + k];
                ph[(size_t)wl * M + k] = atan2(X.y, X.x);
            }
            __syncthreads();
            double* un = reinterpret_cast<double*>(out);    // X is dead after atan2
            for (int wl = tid; wl < nw; wl += kThreads) {
                const double* f = ph + (size_t)wl * M;
                double* u = un + (size_t)wl * M;
                u[0] = f[0];
                for (int i = 1; i < M; i++) {
                    double diff = f[i] - f[i - 1];
                    double corr = 0.0;
                    if (diff > kPi) corr = -2.0 * kPi;
                    else if (diff < -kPi) corr = 2.0 * kPi;
                    u[i] = u[i - 1] + diff + corr;
                }
            }
            __syncthreads();
            for (int idx = tid; idx < nw * M; idx += kThreads) {
                int wl = idx / M, k = idx - wl * M;
                const double* u = un + (size_t)wl * M;
                double g;
                if (M < 3) g = 0.0;
                else if (k == 0) g = -(u[1] - u[0]);
                else if (k == M - 1) g = -(u[M - 1] - u[M - 2]);
                else g = -(u[k + 1] - u[k - 1]) / 2.0;
                if (g > 100.0) g = 100.0;
                if (g < -100.0) g = -100.0;
                double* o = p.phase + (((int64_t)s * nwin + w0 + wb + wl)) * (int64_t)(3 * M);
                o[k] = ph[(size_t)wl * M + k];
                o[M + k] = u[k];
                o[2 * M + k] = g;
            }
        }
        __syncthreads();
    }
}

// Host-side launcher.  Returns the dynamic shared memory it used (0 on error).
size_t window_fft_smem_bytes(const Params& p, int tile_windows, int kThreads) {
    const int N = p.N, M = N / 2;
    int q = M / 8;
    int wpc = q > 0 ? kThreads / q : kThreads;
    if (wpc < 1) wpc = 1;
    if (wpc > p.k1_wpc_cap) wpc = p.k1_wpc_cap;
    size_t Lt = p.feed ? (size_t)wpc * N : (size_t)(tile_windows - 1) * p.hop + N;
    size_t doubles = ((Lt + 1) & ~(size_t)1) + ((tile_windows + 1) & ~1);
    size_t bytes = doubles * 8 + 2 * (size_t)wpc * M * 16;
    if (p.select == 1) bytes += (((size_t)wpc * M + 1) & ~(size_t)1) * 4;
    if (p.detrend == 1) bytes += (Lt + kThreads) * 8;
    return bytes;
}

int window_fft_pick_tile(const Params& p, int kThreads) {
    // tile sized so that the staged samples stay near 16 KB and, for the IIR detrend, fit the
    // scratch carved out of bufA (Lt <= 2*wpc*M doubles)
    const int N = p.N, M = N / 2;
    int q = M / 8;
    int wpc = q > 0 ? kThreads / q : kThreads;
    if (wpc < 1) wpc = 1;
    if (wpc > p.k1_wpc_cap) wpc = p.k1_wpc_cap;
    if (p.feed) return wpc;
    long budget = N <= 1024 ? 4096 : N + 63;      // doubles (large N: keep shared memory for a third CTA)
    if (N >= 8192) budget = N;                    // shared memory is full: one window per tile
    long t = (budget - N) / p.hop + 1;
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    if (p.chunk_nwin < t) t = (long)p.chunk_nwin;
    // keep tiles a multiple of wpc so no iteration runs half empty
    if (t > wpc) t -= t % wpc;
    return (int)t;
}

template <int NT, int LN, int CAP>
static cudaError_t launch_nt(Params p, cudaStream_t stream) {
    p.tile_windows = window_fft_pick_tile(p, NT);
    size_t smem = window_fft_smem_bytes(p, p.tile_windows, NT);
    static std::atomic<unsigned long long> attr_seen{0};
    if (first_launch_on_device(attr_seen)) {
        cudaError_t e = cudaFuncSetAttribute(window_fft_kernel<NT, LN, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
    }
    if (smem > 232448) return cudaErrorInvalidValue;
    dim3 grid((unsigned)((p.chunk_nwin + p.tile_windows - 1) / p.tile_windows), (unsigned)p.n_series);
    window_fft_kernel<NT, LN, CAP><<<grid, NT, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_window_fft(Params p, cudaStream_t stream, const char** which) {
    if (which) *which = "window_fft";
    // WAVESPEC_K1="threads,wpc" selects the run-time-N kernel with that CTA size and
    // concurrent-window cap (tuning hook)
    static int env_nt = -1, env_wpc = 0;
    if (env_nt < 0) {
        env_nt = 0;
        if (const char* e = getenv("WAVESPEC_K1")) sscanf(e, "%d,%d", &env_nt, &env_wpc);
    }
    if (env_nt > 0) {
        p.k1_wpc_cap = env_wpc > 0 ? env_wpc : 4;
        if (env_nt >= 512) return launch_nt<512, 0, 0>(p, stream);
        return env_nt >= 256 ? launch_nt<256, 0, 0>(p, stream) : launch_nt<128, 0, 0>(p, stream);
    }
    if (window_fft_warp_supported(p)) {
        if (which) *which = "window_fft_warp";
        return launch_window_fft_warp(p, stream);
    }
    // measured (profiles/README.md): 256 threads everywhere; 8 concurrent windows below N = 1024.
    // The window lengths of BASELINE.json's configs get compile-time-N instances.
    p.k1_wpc_cap = p.N >= 1024 ? 4 : 8;
    switch (p.N) {
        case 256:  return launch_nt<256, 8, 8>(p, stream);
        case 512:  return launch_nt<256, 9, 8>(p, stream);
        case 1024: return launch_nt<256, 10, 4>(p, stream);
        case 2048: return launch_nt<256, 11, 4>(p, stream);
        case 4096: return launch_nt<256, 12, 4>(p, stream);
        default:   return launch_nt<256, 0, 0>(p, stream);
    }
}

}  // namespace ws
