// ws_common.cuh — shared device-side definitions for the wavespec kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

namespace ws {

constexpr int kMaxTopK = 32;
constexpr int kRowFields = 15;

// Everything a kernel needs to process one batch of equally long series.
struct Params {
    const double* series;      // [n_series][series_len], device
    int64_t series_stride;     // elements between series
    int32_t n_series;
    int32_t series_len;
    int32_t N;                 // window length (power of two)
    int32_t log2N;
    int32_t hop;
    int32_t K;                 // top_k
    int32_t row_stride;
    int64_t nwin;              // windows per series (also the per-series stride of every output plane)
    int64_t win_offset;        // first window this launch covers
    int64_t chunk_nwin;        // number of windows this launch covers (feed plane holds exactly these)
    int32_t band_lo, band_hi;  // inclusive bin band, already clipped; lo > hi means empty
    int32_t detrend;           // WAVESPEC_DETREND_*
    int32_t has_window;        // window table present
    int32_t select;            // WAVESPEC_SELECT_*
    double iir_alpha, iir_c;   // trend IIR coefficients (A2a)
    double sample_rate_seconds;
    const double2* tw;         // exp(-2 pi i m / N), m = 0..N-1
    const double* wtab;        // window coefficients w[i], i = 0..N-1 (or nullptr)
    const double* apow;        // iir_alpha^j, j = 0..N-1 (or nullptr)
    const double2* rowtab;     // per bin k < N/2: (N/k, N/(2 pi k)) for the result rows (or nullptr)
    const double* feed;        // optional pre-built per-window feed [n_series][nwin][N] (PLA), or nullptr
    double* spectra;           // [n_series][spec_nwin][N] interleaved, or nullptr; window w of series s is row
                               // s * spec_nwin + (w - spec_w0)  (spec_nwin = nwin, spec_w0 = 0 for a caller's
                               // plane; a window-range scratch of the phase path uses its own stride)
    int64_t spec_nwin, spec_w0;
    double* rows;              // [n_series][nwin][K][row_stride], or nullptr
    int32_t* bins;             // [n_series][nwin][K], or nullptr
    double* waves;             // [n_series][nwin][K] A8a, or nullptr
    double* contrib;           // [n_series][nwin][K] A8b (input of the weight-Kalman scan), or nullptr
    double* phase;             // [n_series][nwin][3][N/2], or nullptr
    double2* band_buf;         // scratch [n_series][chunk_nwin][band]: in-band bins handed from the
                               // sliding kernel to the rows kernel (ws_rows.cu), or nullptr
    long long* dbg;            // optional [1024][4] clock64 phase stamps of the first CTAs (WAVESPEC_TIMING_FILE), or nullptr
    int32_t tile_windows;      // windows per CTA tile
    int32_t k1_wpc_cap;        // per-window FFT: max windows transformed concurrently by a CTA
    int32_t prefetch_tiles;    // sliding kernels: a CTA pulls the new samples of the tile this many tiles ahead into L2 (0: off)
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// 64-bit warp shuffles
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// (power, bin) ordering used by every top-K rule of the reference: larger power first; on equal
// power the candidate met first in the ascending bin scan stays ahead (strict '>' insertion,
// Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:546-553).  NaN never wins.
__device__ __forceinline__ bool better(double p, int pos, double bp, int bpos) {
    return (p > bp) || (p == bp && pos < bpos);
}

// Stages samples [0, count) of `g` into shared x[0, count) with 128-bit global loads: the run is split
// at its first 16-byte boundary (a window tile starts at an arbitrary sample, so `g` is 0 or 8 mod
// 16), the aligned middle moves as double2, and the odd sample at either end moves alone.  Only the
// first `valid` samples exist (a tile may run past the end of its series): the rest read as 0.
__device__ __forceinline__ void stage_samples(const double* __restrict__ g, int count, int valid, double* x,
                                              int tid, int nthreads) {
    const int n = valid < 0 ? 0 : (valid < count ? valid : count);
    const int head = (int)((reinterpret_cast<unsigned long long>(g) >> 3) & 1ull);
    const int pairs = n > head ? (n - head) >> 1 : 0;
    const double2* g2 = reinterpret_cast<const double2*>(g + head);
    // All of a thread's loads are issued before its first store: the tile is staged while the other
    // CTAs of the SM keep the memory system full of spectrum stores, and a load then takes thousands
    // of cycles — a load / store / load loop would pay that latency once per trip.  The odd samples at
    // either end are fetched by the last warp alongside its pairs.
    const int last = head + 2 * pairs;           // first sample after the aligned middle
    const bool odd_head = head && tid == nthreads - 1;
    const bool odd_tail = tid == nthreads - 2 && last < n;
    double e = 0.0;
    if (odd_head && n > 0) e = g[0];
    if (odd_tail) e = g[last];
    constexpr int kInFlight = 4;
    for (int j0 = tid; j0 < pairs; j0 += kInFlight * nthreads) {
        double2 v[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; u++) {
            const int j = j0 + u * nthreads;
            v[u] = j < pairs ? g2[j] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < kInFlight; u++) {
            const int j = j0 + u * nthreads;
            if (j < pairs) { x[head + 2 * j] = v[u].x; x[head + 2 * j + 1] = v[u].y; }
        }
    }
    if (odd_head) x[0] = e;
    if (odd_tail) x[last] = e;
    for (int i = (last < n ? last + 1 : last) + tid; i < count; i += nthreads) x[i] = 0.0;   // beyond the series
}

// The tiles of a series overlap in all but their newest T * hop samples, and those have been read by
// nobody when the tile starts: its staging would wait for a DRAM read that queues behind the spectrum
// stores of the whole chip (several thousand cycles).  So every CTA first asks L2 for the new samples
// of the tile `ahead` tiles further on, which some SM will start about a wave later.
__device__ __forceinline__ void prefetch_future_tile(const double* __restrict__ series, int64_t series_len,
                                                     int64_t first_new, int count, int tid) {
    const int64_t i = first_new + (int64_t)tid * 16;                 // one 128-byte line per thread
    if (tid * 16 < count + 16 && i < series_len)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(series + i));
}

// Host side: true the first time a kernel instantiation is launched on the current device (the
// opt-in dynamic shared-memory size is a per-device function attribute).  `seen` is the caller's
// static bit mask, one bit per device; launches come from several host threads (one worker per
// device), hence the atomic.
inline bool first_launch_on_device(std::atomic<unsigned long long>& seen) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return (seen.fetch_or(bit) & bit) == 0;
}

}  // namespace ws
