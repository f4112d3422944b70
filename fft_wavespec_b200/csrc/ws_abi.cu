// ws_abi.cu — C ABI of libwavespec.so: session, coefficient tables, job table, dispatch.
// Entry points and the reference interfaces they replace are documented in
// include/wavespec_abi.h (imports.mqh:5-21 and the two Legacy declarations).
//
// No CPU fallback: when no CUDA device can be opened every compute call returns
// WAVESPEC_BACKEND_UNAVAILABLE and says why through gpu_get_last_error_w.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/wavespec_abi.h"
#include "ws_common.cuh"
#include "ws_series.h"

namespace {

using ws::Params;

thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};
const char* g_last_kernel = "none";

int fail(int code, const std::string& msg) { t_last_error = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    int code = (e == cudaErrorMemoryAllocation) ? WAVESPEC_NO_MEM
             : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
                   ? WAVESPEC_BACKEND_UNAVAILABLE : WAVESPEC_INTERNAL_ERROR;
    return fail(code, m);
}
#define WS_CUDA(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(e__, what); } while (0)

bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

// ---- coefficient tables ------------------------------------------------------------------------
const double kPi = 3.14159265358979323846;   // MQL5 M_PI

// Window coefficients with the reference's own expressions and evaluation order
// (Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:1126-1156; type 5: Legacy/WaveSpecZZ_gpu_wip.mq5:954).
// Evaluated once per (N, type) on the host in IEEE double, i.e. the same values the MQL5 loop
// recomputes for every bar.
void build_window(int n, int type, std::vector<double>& w) {
    w.resize(n);
    for (int i = 0; i < n; i++) {
        double v = 1.0;
        switch (type) {
            case WAVESPEC_WINDOW_HANN:     v = 0.5 * (1.0 - std::cos(2.0 * kPi * i / (n - 1))); break;
            case WAVESPEC_WINDOW_HAMMING:  v = 0.54 - 0.46 * std::cos(2.0 * kPi * i / (n - 1)); break;
            case WAVESPEC_WINDOW_BLACKMAN: v = 0.42 - 0.5 * std::cos(2.0 * kPi * i / (n - 1))
                                               + 0.08 * std::cos(4.0 * kPi * i / (n - 1)); break;
            case WAVESPEC_WINDOW_BARTLETT: v = 1.0 - std::fabs((2.0 * i - n + 1) / (n - 1)); break;
            case WAVESPEC_WINDOW_HANN_WIP: v = 0.5 - 0.5 * std::cos((2.0 * kPi * i) / (double)(n - 1)); break;
            default: break;
        }
        w[i] = v;
    }
}

struct DeviceBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DeviceBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t b) {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = b;
        return b ? cudaMalloc(&p, b) : cudaSuccess;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// Size-keyed cache of device buffers for the job table: a sliding batch job needs ~1 GB of row
// storage, and cudaMalloc/cudaFree of that size per job (cudaFree also synchronises the device)
// costs more than the kernel.  Buffers return to the pool in gpu_free_job and die in gpu_shutdown.
struct DevicePool {
    std::mutex mu;
    std::multimap<size_t, void*> free_list;
    size_t cached_bytes = 0;
    static constexpr size_t kMaxCached = (size_t)24 << 30;
    cudaError_t get(size_t bytes, void** out) {
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_list.lower_bound(bytes);
            if (it != free_list.end() && it->first <= bytes + bytes / 8) {
                *out = it->second; cached_bytes -= it->first; free_list.erase(it);
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(out, bytes);
        if (e == cudaErrorMemoryAllocation) { trim(); cudaGetLastError(); e = cudaMalloc(out, bytes); }
        return e;
    }
    void put(void* p, size_t bytes) {
        std::lock_guard<std::mutex> lk(mu);
        if (cached_bytes + bytes > kMaxCached) { cudaFree(p); return; }
        free_list.emplace(bytes, p); cached_bytes += bytes;
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& kv : free_list) cudaFree(kv.second);
        free_list.clear(); cached_bytes = 0;
    }
};
DevicePool g_pool;

struct PooledBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~PooledBuf() { if (p) g_pool.put(p, bytes); }
    cudaError_t alloc(size_t b) { bytes = b; return g_pool.get(b, &p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct Job {
    std::mutex mu;
    int kind = 0;                 // 0 single window, 1 sliding batch
    PooledBuf d_series, d_rows;
    cudaEvent_t done = nullptr;
    int64_t rows = 0;             // rows produced
    int32_t stride = 15, top_k = 0;
    int status = WAVESPEC_OK;
    ~Job() { if (done) cudaEventDestroy(done); }
};

struct Session {
    std::mutex mu;
    bool open = false;
    int device = 0;
    std::vector<cudaStream_t> streams;
    cudaStream_t side = nullptr;       // Kalman4D runs here beside the FFT kernels of the same call
    std::atomic<uint32_t> rr{0};
    std::map<int, std::unique_ptr<DeviceBuf>> tw;                       // N -> exp(-2 pi i m/N)
    std::map<std::pair<int, int>, std::unique_ptr<DeviceBuf>> win;      // (N,type) -> w[i]
    std::map<std::pair<int, double>, std::unique_ptr<DeviceBuf>> apow;  // (N,alpha) -> alpha^j
    std::map<int64_t, std::shared_ptr<Job>> jobs;
    int64_t next_job = 1;
    // per-stream band hand-off buffer (ws_sliding.cu -> ws_rows.cu); work on one stream is ordered,
    // so one buffer per stream can be reused launch after launch without synchronising
    std::map<cudaStream_t, std::unique_ptr<DeviceBuf>> band_scratch;
    std::map<cudaStream_t, std::unique_ptr<DeviceBuf>> phase_scratch;   // spectra of a window range (phase path)
};
Session g_s;

cudaStream_t pick_stream() {
    if (g_s.streams.empty()) return nullptr;
    return g_s.streams[g_s.rr.fetch_add(1) % g_s.streams.size()];
}

int ensure_open() {
    if (!g_s.open) return fail(WAVESPEC_BACKEND_UNAVAILABLE, "gpu_init has not been called (or failed)");
    cudaError_t e = cudaSetDevice(g_s.device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return WAVESPEC_OK;
}

int get_twiddles(int N, const double2** out) {
    std::lock_guard<std::mutex> lk(g_s.mu);
    auto it = g_s.tw.find(N);
    if (it == g_s.tw.end()) {
        std::vector<double> h(2 * (size_t)N);
        for (int m = 0; m < N; m++) {
            // exact table twiddles (long double evaluation, rounded once)
            long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)m / (long double)N;
            h[2 * m] = (double)cosl(a);
            h[2 * m + 1] = (double)sinl(a);
        }
        // exact values on the axes and diagonals
        h[0] = 1.0; h[1] = 0.0;
        if (N >= 2) { h[2 * (N / 2)] = -1.0; h[2 * (N / 2) + 1] = 0.0; }
        if (N >= 4) { h[2 * (N / 4)] = 0.0; h[2 * (N / 4) + 1] = -1.0; h[2 * (3 * N / 4)] = 0.0; h[2 * (3 * N / 4) + 1] = 1.0; }
        auto buf = std::make_unique<DeviceBuf>();
        WS_CUDA(buf->alloc(h.size() * 8), "cudaMalloc(twiddles)");
        WS_CUDA(cudaMemcpy(buf->p, h.data(), h.size() * 8, cudaMemcpyHostToDevice), "cudaMemcpy(twiddles)");
        it = g_s.tw.emplace(N, std::move(buf)).first;
    }
    *out = it->second->as<double2>();
    return WAVESPEC_OK;
}

int get_window(int N, int type, const double** out) {
    *out = nullptr;
    if (type == WAVESPEC_WINDOW_NONE) return WAVESPEC_OK;
    std::lock_guard<std::mutex> lk(g_s.mu);
    auto key = std::make_pair(N, type);
    auto it = g_s.win.find(key);
    if (it == g_s.win.end()) {
        std::vector<double> h;
        build_window(N, type, h);
        auto buf = std::make_unique<DeviceBuf>();
        WS_CUDA(buf->alloc(h.size() * 8), "cudaMalloc(window)");
        WS_CUDA(cudaMemcpy(buf->p, h.data(), h.size() * 8, cudaMemcpyHostToDevice), "cudaMemcpy(window)");
        it = g_s.win.emplace(key, std::move(buf)).first;
    }
    *out = it->second->as<double>();
    return WAVESPEC_OK;
}

int get_apow(int N, double alpha, const double** out) {
    std::lock_guard<std::mutex> lk(g_s.mu);
    auto key = std::make_pair(N, alpha);
    auto it = g_s.apow.find(key);
    if (it == g_s.apow.end()) {
        std::vector<double> h(N);
        for (int j = 0; j < N; j++) h[j] = std::pow(alpha, (double)j);
        auto buf = std::make_unique<DeviceBuf>();
        WS_CUDA(buf->alloc(h.size() * 8), "cudaMalloc(apow)");
        WS_CUDA(cudaMemcpy(buf->p, h.data(), h.size() * 8, cudaMemcpyHostToDevice), "cudaMemcpy(apow)");
        it = g_s.apow.emplace(key, std::move(buf)).first;
    }
    *out = it->second->as<double>();
    return WAVESPEC_OK;
}

int validate_cfg(const wavespec_pipeline_cfg* c, int32_t series_len) {
    if (!c) return fail(WAVESPEC_BAD_ARGS, "cfg is null");
    if (!is_pow2(c->window_len) || c->window_len < 2 || c->window_len > 8192)
        return fail(WAVESPEC_BAD_ARGS, "window_len must be a power of two in [2, 8192]");
    if (c->hop < 1) return fail(WAVESPEC_BAD_ARGS, "hop must be >= 1");
    if (c->top_k < 1 || c->top_k > ws::kMaxTopK) return fail(WAVESPEC_BAD_ARGS, "top_k must be in [1, 32]");
    if (c->row_stride < 1) return fail(WAVESPEC_BAD_ARGS, "row_stride must be >= 1");
    if (!(c->min_period > 0.0) || !(c->max_period > 0.0))
        return fail(WAVESPEC_BAD_ARGS, "min_period and max_period must be > 0");
    if (series_len < c->window_len) return fail(WAVESPEC_BAD_ARGS, "series shorter than one window");
    if (c->detrend < 0 || c->detrend > 2) return fail(WAVESPEC_BAD_ARGS, "unknown detrend mode");
    if (c->window_type < 0 || c->window_type > 5) return fail(WAVESPEC_BAD_ARGS, "unknown window type");
    if (c->select < 0 || c->select > 1) return fail(WAVESPEC_BAD_ARGS, "unknown select rule");
    if (c->feed < 0 || c->feed > 1) return fail(WAVESPEC_BAD_ARGS, "unknown feed");
    if (c->detrend == WAVESPEC_DETREND_IIR && !(c->trend_period > 0.0))
        return fail(WAVESPEC_BAD_ARGS, "trend_period must be > 0 for the IIR detrend");
    if (c->window_type != WAVESPEC_WINDOW_NONE && c->window_len < 2)
        return fail(WAVESPEC_BAD_ARGS, "window functions need window_len >= 2");
    return WAVESPEC_OK;
}

// The whole per-bar pipeline on device pointers.  Enqueues on `st`; synchronises only where a
// temporary has to be released (PLA feed chunks, weight-Kalman inputs).
int run_tracker_path(Params p, const wavespec_pipeline_cfg* c, int32_t* d_trk_index, double* d_trk_period,
                     bool plain, cudaStream_t st);

int run_pipeline(const double* d_series, int32_t n_series, int32_t series_len,
                 const wavespec_pipeline_cfg* c, double* d_spectra, double* d_rows, int32_t* d_bins,
                 double* d_waves, double* d_kalman, double* d_phase, double* d_wkalman,
                 cudaStream_t st, int32_t* d_trk_index = nullptr, double* d_trk_period = nullptr) {
    int rc = validate_cfg(c, series_len);
    if (rc) return rc;
    if (n_series < 1 || n_series > 65535) return fail(WAVESPEC_BAD_ARGS, "n_series must be in [1, 65535]");
    if (!d_series) return fail(WAVESPEC_BAD_ARGS, "series is null");
    const int N = c->window_len;
    const int64_t nwin = 1 + (int64_t)(series_len - N) / c->hop;

    Params p;
    std::memset(&p, 0, sizeof p);
    p.series = d_series; p.series_stride = series_len; p.n_series = n_series; p.series_len = series_len;
    p.N = N; p.log2N = ilog2(N); p.hop = c->hop; p.K = c->top_k; p.row_stride = c->row_stride;
    p.nwin = nwin; p.win_offset = 0; p.chunk_nwin = nwin;
    p.spec_nwin = nwin; p.spec_w0 = 0;
    // band: Legacy/...-gpuopt-nodetrend.mq5:540-542
    int lo = (int)std::ceil((double)N / c->max_period);
    int hi = (int)std::floor((double)N / c->min_period);
    if (hi >= N / 2) hi = N / 2 - 1;
    if (lo < 0) lo = 0;
    p.band_lo = lo; p.band_hi = hi;
    p.detrend = c->detrend; p.select = c->select; p.sample_rate_seconds = c->sample_rate_seconds;
    if ((rc = get_twiddles(N, &p.tw))) return rc;
    if ((rc = get_window(N, c->window_type, &p.wtab))) return rc;
    p.has_window = p.wtab != nullptr;
    if (c->detrend == WAVESPEC_DETREND_IIR) {
        // Legacy/...-kalman-fast.mq5:3367-3369
        double omega = 2.0 * kPi / c->trend_period;
        double alpha = (1.0 - std::sin(omega)) / std::cos(omega);
        p.iir_alpha = alpha; p.iir_c = (1.0 - alpha) / 2.0;
        if ((rc = get_apow(N, alpha, &p.apow))) return rc;
    }
    const bool want_wk = d_wkalman != nullptr;
    DeviceBuf tmp_contrib, tmp_bins, tmp_feed, tmp_z;
    p.spectra = d_spectra; p.rows = d_rows; p.bins = d_bins; p.waves = d_waves; p.phase = d_phase;
    if (want_wk) {
        WS_CUDA(tmp_contrib.alloc((size_t)n_series * nwin * c->top_k * 8), "cudaMalloc(contrib)");
        p.contrib = tmp_contrib.as<double>();
        if (!p.bins) {
            WS_CUDA(tmp_bins.alloc((size_t)n_series * nwin * c->top_k * 4), "cudaMalloc(bins)");
            p.bins = tmp_bins.as<int32_t>();
        }
    }
    const bool want_trk = d_trk_index && d_trk_period;
    if ((d_trk_index != nullptr) != (d_trk_period != nullptr))
        return fail(WAVESPEC_BAD_ARGS, "tracker planes come as a pair (index and period)");
    if (want_trk && c->feed == WAVESPEC_FEED_PLA)
        return fail(WAVESPEC_BAD_ARGS, "the tracker plane is not wired to the PLA feed yet");
    if (want_trk && p.band_hi < p.band_lo) return fail(WAVESPEC_BAD_ARGS, "tracker needs a non-empty band");
    const bool any_spectral = (p.spectra || p.rows || p.bins || p.waves || p.phase || p.contrib) && !want_trk;

    if (c->feed == WAVESPEC_FEED_PLA) {
        // PLA lines are window-private (the recursion restarts per window): build them chunk by
        // chunk into a bounded temporary and feed the per-window FFT kernel from it.
        const size_t budget = (size_t)1 << 30;   // bytes of feed per chunk
        int64_t chunk = (int64_t)(budget / ((size_t)n_series * N * 8));
        if (chunk < 1) chunk = 1;
        if (chunk > nwin) chunk = nwin;
        WS_CUDA(tmp_feed.alloc((size_t)n_series * chunk * N * 8), "cudaMalloc(pla feed)");
        if (d_kalman) WS_CUDA(tmp_z.alloc((size_t)n_series * nwin * 8), "cudaMalloc(kalman z)");
        for (int64_t wa = 0; wa < nwin; wa += chunk) {
            const int64_t cn = (wa + chunk <= nwin) ? chunk : nwin - wa;
            cudaError_t e = ws::launch_pla(d_series + wa * c->hop, series_len, n_series, cn, N, c->hop,
                                           c->pla_max_segments, c->pla_max_error, tmp_feed.as<double>(),
                                           nullptr, nullptr, 0, st);
            g_launches++;
            if (e != cudaSuccess) return cuda_fail(e, "pla kernel (recursion deeper than the on-chip stack?)");
            if (d_kalman) {
                // newest sample of each PLA line is the Kalman measurement (:3354-3360)
                // gather column N-1 of every line: one strided 2D copy per series
                for (int s = 0; s < n_series; s++)
                    WS_CUDA(cudaMemcpy2DAsync(tmp_z.as<double>() + (size_t)s * nwin + wa, 8,
                                              tmp_feed.as<double>() + ((size_t)s * cn) * N + (N - 1),
                                              (size_t)N * 8, 8, (size_t)cn, cudaMemcpyDeviceToDevice, st),
                            "cudaMemcpy2DAsync(kalman z)");
            }
            if (any_spectral) {
                Params q = p;
                q.feed = tmp_feed.as<double>(); q.win_offset = wa; q.chunk_nwin = cn;
                WS_CUDA(ws::launch_window_fft(q, st, &g_last_kernel), "window_fft kernel");
                g_launches++;
            }
            WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize(pla chunk)");
        }
        if (d_kalman) {
            ws::KalmanParams kp;
            std::memcpy(&kp, &c->kalman, sizeof kp);
            WS_CUDA(ws::launch_kalman4d(tmp_z.as<double>(), nwin, 1, n_series, nwin, kp, d_kalman, st), "kalman4d kernel");
            g_launches++;
        }
    } else {
        // Kalman4D (A9) only reads the series: it is one thread per series and strictly sequential over
        // bars (0.7 s for 1M bars), so it is forked onto the session's side stream and runs beside the
        // FFT kernels of this call; `st` joins it at the end.
        cudaEvent_t kalman_join = nullptr;
        if (d_kalman) {
            ws::KalmanParams kp;
            std::memcpy(&kp, &c->kalman, sizeof kp);
            cudaStream_t ks = (g_s.side && (any_spectral || want_trk)) ? g_s.side : st;
            if (ks != st) {
                cudaEvent_t fork = nullptr;
                WS_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming), "cudaEventCreate");
                WS_CUDA(cudaEventRecord(fork, st), "cudaEventRecord(fork)");          // inputs are ready on st
                WS_CUDA(cudaStreamWaitEvent(ks, fork, 0), "cudaStreamWaitEvent(fork)");
                cudaEventDestroy(fork);
            }
            WS_CUDA(ws::launch_kalman4d(d_series + (N - 1), series_len, c->hop, n_series, nwin, kp, d_kalman, ks),
                    "kalman4d kernel");
            g_launches++;
            if (ks != st) {
                WS_CUDA(cudaEventCreateWithFlags(&kalman_join, cudaEventDisableTiming), "cudaEventCreate");
                WS_CUDA(cudaEventRecord(kalman_join, ks), "cudaEventRecord(join)");
            }
        }
        // FFT dispatch for one Params (whole series or a window range): the shared-butterfly sliding
        // kernels for plain hop-1 windows, the per-window kernels otherwise
        auto dispatch_fft = [&](const Params& p) -> int {
            const bool plain = c->hop == 1 && c->detrend == WAVESPEC_DETREND_NONE &&
                               c->window_type == WAVESPEC_WINDOW_NONE && !p.phase;
            if (plain && ws::sliding_shared_supported(p)) {
                // Two ways to produce rows on this path: the fused in-kernel epilogue (default) or a
                // hand-off of the in-band bins to a separate full-occupancy rows kernel
                // (WAVESPEC_SPLIT=1).  Measured on B200 at N=1024 they are within 3 % of each
                // other (profiles/README.md); the fused form needs no scratch and one launch.
                static const bool split = getenv("WAVESPEC_SPLIT") != nullptr;
                if (split && p.win_offset == 0 && p.chunk_nwin == nwin && ws::rows_from_band_supported(p)) {
                    // sliding kernel = pure streaming writer + compact band hand-off; the rows kernel
                    // selects at full occupancy (ws_rows.cu).  The hand-off buffer is bounded: series
                    // (and, for very long series, window ranges) are processed in chunks on one stream.
                    const int band = p.band_hi - p.band_lo + 1;
                    const size_t budget = (size_t)4 << 30;
                    const size_t per_win = (size_t)band * 16;
                    int64_t wchunk = nwin, sgroup = n_series;
                    if (per_win * (size_t)nwin > budget) { sgroup = 1; wchunk = (int64_t)(budget / per_win); }
                    else { sgroup = (int64_t)(budget / (per_win * (size_t)nwin)); if (sgroup > n_series) sgroup = n_series; }
                    if (sgroup < 1) sgroup = 1;
                    if (wchunk < 1) wchunk = 1;
                    void* scratch = nullptr;
                    const size_t sbytes = per_win * (size_t)wchunk * (size_t)sgroup;
                    {
                        std::lock_guard<std::mutex> lk(g_s.mu);
                        auto& slot = g_s.band_scratch[st];
                        if (!slot) slot = std::make_unique<DeviceBuf>();
                        if (slot->bytes < sbytes) {
                            // growing: the old buffer may still be in use by queued kernels
                            WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize(band buffer)");
                            WS_CUDA(slot->alloc(sbytes), "cudaMalloc(band buffer)");
                        }
                        scratch = slot->p;
                    }
                    int rc2 = WAVESPEC_OK;
                    for (int64_t s0 = 0; s0 < n_series && rc2 == WAVESPEC_OK; s0 += sgroup) {
                        const int64_t ns = (s0 + sgroup <= n_series) ? sgroup : n_series - s0;
                        for (int64_t wa = 0; wa < nwin && rc2 == WAVESPEC_OK; wa += wchunk) {
                            Params q = p;
                            q.series = p.series + s0 * p.series_stride; q.n_series = (int32_t)ns;
                            if (p.spectra) q.spectra = p.spectra + s0 * nwin * N;
                            if (p.rows) q.rows = p.rows + s0 * nwin * p.K * p.row_stride;
                            if (p.bins) q.bins = p.bins + s0 * nwin * p.K;
                            if (p.waves) q.waves = p.waves + s0 * nwin * p.K;
                            if (p.contrib) q.contrib = p.contrib + s0 * nwin * p.K;
                            q.win_offset = wa; q.chunk_nwin = (wa + wchunk <= nwin) ? wchunk : nwin - wa;
                            q.band_buf = static_cast<double2*>(scratch);
                            cudaError_t e = ws::launch_sliding_shared(q, st);
                            g_launches++;
                            if (e == cudaSuccess) { e = ws::launch_rows_from_band(q, st); g_launches++; }
                            if (e != cudaSuccess) rc2 = cuda_fail(e, "sliding_shared / rows_from_band kernel");
                        }
                    }
                    g_launches--;      // the common increment below counts one of them
                    if (rc2) return rc2;
                    g_last_kernel = "sliding_shared";
                } else {
                    g_last_kernel = "sliding_shared";
                    WS_CUDA(ws::launch_sliding_shared(p, st, &g_last_kernel), "sliding_shared kernel");
                }
            } else {
                WS_CUDA(ws::launch_window_fft(p, st, &g_last_kernel), "window_fft kernel");
            }
            g_launches++;
            return WAVESPEC_OK;
        };
        if (any_spectral) {
            if (p.phase && ws::phase_from_spectra_supported(N)) {
                // A6 behind the FFT: the phase chain only needs the window's spectrum, so the fastest
                // FFT kernel runs without it and ws_phase.cu follows on the same stream — on the
                // caller's spectra plane when there is one, else on a scratch plane per window range
                Params q = p;
                q.phase = nullptr;
                if (p.spectra) {
                    if ((rc = dispatch_fft(q))) return rc;
                    WS_CUDA(ws::launch_phase_from_spectra(p.spectra, nwin, 0, n_series, 0, nwin, nwin, N, p.phase, st),
                            "phase_chain kernel");
                    g_launches++;
                } else {
                    int64_t chunk = (int64_t)(((size_t)2 << 30) / ((size_t)n_series * N * 8));
                    if (const char* e = getenv("WAVESPEC_PHASE_CHUNK")) { long v = atol(e); if (v > 0) chunk = v; }   // test hook
                    if (chunk < 1) chunk = 1;
                    if (chunk > nwin) chunk = nwin;
                    // per-stream scratch, kept between calls (work on a stream is ordered, so the next
                    // call may reuse it without waiting)
                    double* scratch_p = nullptr;
                    {
                        const size_t sbytes = (size_t)n_series * chunk * N * 8;
                        std::lock_guard<std::mutex> lk(g_s.mu);
                        auto& slot = g_s.phase_scratch[st];
                        if (!slot) slot = std::make_unique<DeviceBuf>();
                        if (slot->bytes < sbytes) {
                            WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize(phase scratch)");
                            WS_CUDA(slot->alloc(sbytes), "cudaMalloc(phase scratch spectra)");
                        }
                        scratch_p = slot->as<double>();
                    }
                    for (int64_t wa = 0; wa < nwin; wa += chunk) {
                        const int64_t cn = (wa + chunk <= nwin) ? chunk : nwin - wa;
                        q.win_offset = wa; q.chunk_nwin = cn;
                        q.spectra = scratch_p; q.spec_nwin = chunk; q.spec_w0 = wa;
                        if ((rc = dispatch_fft(q))) return rc;
                        WS_CUDA(ws::launch_phase_from_spectra(scratch_p, chunk, wa, n_series, wa, cn, nwin, N,
                                                              p.phase, st), "phase_chain kernel");
                        g_launches++;
                    }
                }
            } else {
                if ((rc = dispatch_fft(p))) return rc;
            }
        }
        if (want_trk) {
            const bool plain = c->hop == 1 && c->detrend == WAVESPEC_DETREND_NONE &&
                               c->window_type == WAVESPEC_WINDOW_NONE && !p.phase;
            if ((rc = run_tracker_path(p, c, d_trk_index, d_trk_period, plain, st))) return rc;
        }
        if (kalman_join) {
            WS_CUDA(cudaStreamWaitEvent(st, kalman_join, 0), "cudaStreamWaitEvent(kalman)");
            cudaEventDestroy(kalman_join);
        }
    }
    if (want_wk) {
        // measurement = close[bar] (Legacy/WaveSpecZZ_1.0.4-kalman.mq5:284)
        WS_CUDA(ws::launch_wkalman(p.contrib, p.bins, d_series + (N - 1), series_len, c->hop, n_series, nwin,
                                   c->top_k, c->wk_process_noise, c->wk_meas_noise, c->wk_init_variance,
                                   d_wkalman, st), "wkalman kernel");
        g_launches++;
        WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize(wkalman)");   // temporaries die here
    }
    return WAVESPEC_OK;
}

// A13 host logic.  Which tracker a band bin matches, when trackers are appended, expire and shift
// (Legacy/...-kalman-fast.mq5:1418-1529) depends on PERIODS only — never on the data — and every bar
// presents the same bins in the same order.  So the tracker structure is one deterministic
// sequence shared by all series, and it settles: from some bar B0 on it repeats itself, after
// which the 12 slots (sticky, refilled only from unused trackers) cannot change any more.  This
// returns B0 (first bar whose end state equals the previous bar's), or -1 if no fixed point shows
// up within `limit` bars (then the device walks every bar).
int64_t tracker_fixed_point(int N, int lo, int hi, double tol, int max_inactive, int64_t limit) {
    struct T { double period; int idx; int inactive; bool active; };
    std::vector<T> tr, prev;
    auto same = [](double p1, double p2, double tolp) {
        if (p1 <= 0 || p2 <= 0) return false;
        double diff = std::fabs(p1 - p2), avg = (p1 + p2) / 2.0;
        return (diff / avg) * 100.0 <= tolp;
    };
    for (int64_t b = 0; b < limit; b++) {
        for (int j = lo; j <= hi; j++) {
            double period = j > 0 ? (double)N / j : 0;
            if (period <= 0) continue;
            int best = -1; double smallest = 999999;
            for (size_t i = 0; i < tr.size(); i++) {
                if (tr[i].inactive > 0) continue;
                double diff = std::fabs(tr[i].period - period);
                if (same(period, tr[i].period, tol) && diff < smallest) { smallest = diff; best = (int)i; }
            }
            if (best >= 0) { tr[best].period = period; tr[best].idx = j; tr[best].active = true; tr[best].inactive = 0; }
            else if ((int)tr.size() < ws::kTrackerCap) tr.push_back(T{period, j, 0, true});
        }
        for (int i = (int)tr.size() - 1; i >= 0; i--)
            if (!tr[i].active && ++tr[i].inactive >= max_inactive) tr.erase(tr.begin() + i);
        for (auto& t : tr) t.active = false;
        bool eq = prev.size() == tr.size();
        for (size_t i = 0; eq && i < tr.size(); i++)
            eq = prev[i].period == tr[i].period && prev[i].idx == tr[i].idx && prev[i].inactive == tr[i].inactive;
        if (eq && b > 0) return b;
        prev = tr;
    }
    return -1;
}

// A13: FFT kernel -> compact band hand-off -> tracker kernel, chunked over windows so that the
// hand-off buffer stays bounded; the tracker state of every series persists across chunks.  The
// sequential kernel only walks the bars up to the structural fixed point; the rest is a broadcast.
int run_tracker_path(Params p, const wavespec_pipeline_cfg* c, int32_t* d_trk_index, double* d_trk_period,
                     bool plain, cudaStream_t st) {
    const int band = p.band_hi - p.band_lo + 1;
    const int64_t nwin = p.nwin;
    const bool sel = p.rows || p.bins || p.waves || p.contrib;
    const bool other = sel || p.spectra || p.phase;
    static const bool walk_all = getenv("WAVESPEC_TRACKER_WALK_ALL") != nullptr;      // testing hook
    int64_t fixed = walk_all ? -1 : tracker_fixed_point(p.N, p.band_lo, p.band_hi, c->tracker_tolerance,
                                                        c->tracker_max_inactive, 4096);
    // bars the sequential kernel has to walk: one past the fixed point (its slots are final)
    const int64_t walk = (fixed < 0 || fixed + 1 >= nwin) ? nwin : fixed + 1;
    const int64_t need = other ? nwin : walk;                 // windows the FFT kernels must cover
    const size_t per_win = (size_t)band * 16 * (size_t)p.n_series;
    int64_t wchunk = (int64_t)(((size_t)2 << 30) / per_win);
    if (wchunk < 1) wchunk = 1;
    if (wchunk > need) wchunk = need;
    DeviceBuf scratch, states;
    WS_CUDA(scratch.alloc(per_win * (size_t)wchunk), "cudaMalloc(band buffer)");
    WS_CUDA(states.alloc(sizeof(ws::TrackerState) * (size_t)p.n_series), "cudaMalloc(tracker state)");
    // the sliding kernel hands the band over only in its split form (insertion rule, K <= 8)
    bool use_sliding = plain && ws::sliding_shared_supported(p) && (!sel || ws::rows_from_band_supported(p));
    for (int64_t wa = 0; wa < need; wa += wchunk) {
        Params q = p;
        q.win_offset = wa; q.chunk_nwin = (wa + wchunk <= need) ? wchunk : need - wa;
        q.band_buf = scratch.as<double2>();
        if (use_sliding) {
            WS_CUDA(ws::launch_sliding_shared(q, st), "sliding_shared kernel");
            g_launches++;
            if (sel) { WS_CUDA(ws::launch_rows_from_band(q, st), "rows_from_band kernel"); g_launches++; }
            g_last_kernel = "sliding_shared";
        } else {
            WS_CUDA(ws::launch_window_fft(q, st, &g_last_kernel), "window_fft kernel");
            g_launches++;
        }
        if (wa < walk) {
            const int64_t np = (wa + q.chunk_nwin <= walk) ? q.chunk_nwin : walk - wa;
            WS_CUDA(ws::launch_tracker(q.band_buf, p.band_lo, band, p.n_series, q.chunk_nwin, np, wa, nwin, p.N,
                                       c->tracker_tolerance, c->tracker_max_inactive,
                                       states.as<ws::TrackerState>(), d_trk_index, d_trk_period, st), "tracker kernel");
            g_launches++;
        }
    }
    if (walk < nwin) {
        WS_CUDA(ws::launch_tracker_fill(p.n_series, nwin, walk - 1, d_trk_index, d_trk_period, st), "tracker fill kernel");
        g_launches++;
    }
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize(tracker)");      // scratch dies here
    return WAVESPEC_OK;
}

void cfg_for_cycles(wavespec_pipeline_cfg* c, int32_t window_len, int32_t hop, int32_t top_k,
                    double min_period, double max_period, double sample_rate_seconds, int32_t stride) {
    wavespec_default_cfg(c, window_len);
    c->hop = hop; c->top_k = top_k; c->row_stride = stride;
    c->min_period = min_period; c->max_period = max_period; c->sample_rate_seconds = sample_rate_seconds;
    c->outputs = WAVESPEC_OUT_ROWS;
}

int check_method(int32_t method) {
    // 0 FFT ridge, -1 auto, 1 MUSIC/ESPRIT (served by the FFT ridge extractor; rows say method=0)
    if (method < -1 || method > 1) return fail(WAVESPEC_BAD_ARGS, "method must be -1, 0 or 1");
    return WAVESPEC_OK;
}

std::shared_ptr<Job> find_job(int64_t id) {
    std::lock_guard<std::mutex> lk(g_s.mu);
    auto it = g_s.jobs.find(id);
    return it == g_s.jobs.end() ? nullptr : it->second;
}

int submit_common(const double* series, int32_t series_len, const wavespec_pipeline_cfg& c, int kind,
                  int64_t* job_id) {
    if (!job_id) return fail(WAVESPEC_BAD_ARGS, "job_id is null");
    *job_id = 0;
    int rc = ensure_open();
    if (rc) return rc;
    if (!series) return fail(WAVESPEC_BAD_ARGS, "series is null");
    if ((rc = validate_cfg(&c, series_len))) return rc;
    auto job = std::make_shared<Job>();
    job->kind = kind; job->stride = c.row_stride; job->top_k = c.top_k;
    const int64_t nwin = 1 + (int64_t)(series_len - c.window_len) / c.hop;
    job->rows = nwin * c.top_k;
    WS_CUDA(job->d_series.alloc((size_t)series_len * 8), "cudaMalloc(series)");
    WS_CUDA(job->d_rows.alloc((size_t)job->rows * c.row_stride * 8), "cudaMalloc(rows)");
    WS_CUDA(cudaEventCreateWithFlags(&job->done, cudaEventDisableTiming), "cudaEventCreate");
    cudaStream_t st = pick_stream();
    // the caller may reuse `series` as soon as we return (1.1.0 :1313-1339): the copy below
    // has left the caller's buffer by the time cudaMemcpyAsync returns for pageable memory,
    // and we wait for it explicitly so pinned callers are covered too.
    WS_CUDA(cudaMemcpyAsync(job->d_series.p, series, (size_t)series_len * 8, cudaMemcpyHostToDevice, st),
            "cudaMemcpyAsync(series)");
    cudaEvent_t copied;
    WS_CUDA(cudaEventCreateWithFlags(&copied, cudaEventDisableTiming), "cudaEventCreate");
    cudaEventRecord(copied, st);
    rc = run_pipeline(job->d_series.as<double>(), 1, series_len, &c, nullptr, job->d_rows.as<double>(),
                      nullptr, nullptr, nullptr, nullptr, nullptr, st);
    if (rc == WAVESPEC_OK) {
        cudaError_t e = cudaEventRecord(job->done, st);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventRecord");
    }
    cudaEventSynchronize(copied);
    cudaEventDestroy(copied);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_s.mu);
    int64_t id = g_s.next_job++;
    g_s.jobs[id] = job;
    *job_id = id;
    return WAVESPEC_OK;
}

int try_get_common(int64_t job_id, double* out, int64_t out_cap_doubles, int32_t out_stride, int kind,
                   int32_t* out_len, int32_t* ready) {
    if (out_len) *out_len = 0;
    if (ready) *ready = 0;
    int rc = ensure_open();
    if (rc) return rc;
    if (!out || !out_len || !ready) return fail(WAVESPEC_BAD_ARGS, "null output pointer");
    auto job = find_job(job_id);
    if (!job) return fail(WAVESPEC_BAD_ARGS, "unknown job id");
    std::lock_guard<std::mutex> lk(job->mu);
    if (job->kind != kind) return fail(WAVESPEC_BAD_ARGS, "job id belongs to the other job kind");
    cudaError_t q = cudaEventQuery(job->done);
    if (q == cudaErrorNotReady) return WAVESPEC_NOT_READY;
    if (q != cudaSuccess) return cuda_fail(q, "job failed on the device");
    int64_t rows = job->rows;
    if (kind == 0) {
        // single window: caller's stride may differ from the job's (fixed 15) -> repack per row
        int64_t cap_rows = out_cap_doubles;          // here: capacity in rows
        if (rows > cap_rows) rows = cap_rows;
        const int m = out_stride < job->stride ? out_stride : job->stride;
        std::vector<double> h((size_t)job->rows * job->stride);
        WS_CUDA(cudaMemcpy(h.data(), job->d_rows.p, h.size() * 8, cudaMemcpyDeviceToHost), "cudaMemcpy(rows)");
        for (int64_t r = 0; r < rows; r++) {
            for (int i = 0; i < m; i++) out[r * out_stride + i] = h[r * job->stride + i];
            for (int i = m; i < out_stride; i++) out[r * out_stride + i] = 0.0;
        }
    } else {
        int64_t cap_rows = out_cap_doubles / job->stride;
        if (rows > cap_rows) rows = cap_rows - cap_rows % job->top_k;   // whole windows only
        if (rows > 0)
            WS_CUDA(cudaMemcpy(out, job->d_rows.p, (size_t)rows * job->stride * 8, cudaMemcpyDeviceToHost),
                    "cudaMemcpy(rows)");
    }
    if (rows > 0x7fffffff) return fail(WAVESPEC_BAD_ARGS, "row count does not fit int32 out_len");
    *out_len = (int32_t)rows;
    *ready = 1;
    return WAVESPEC_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int32_t wavespec_version(void) { return 10000; }
int64_t wavespec_launch_count(void) { return g_launches.load(); }
const char* wavespec_last_kernel(void) { return g_last_kernel; }

int32_t gpu_init(int32_t device_index, int32_t stream_count) {
    std::lock_guard<std::mutex> lk(g_s.mu);
    if (g_s.open) {
        if (device_index != g_s.device)
            return fail(WAVESPEC_BAD_ARGS, "session already open on another device; call gpu_shutdown first");
        return WAVESPEC_OK;    // idempotent (Fetcher and indicator may share the process)
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(WAVESPEC_BACKEND_UNAVAILABLE,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device_index < 0 || device_index >= count) return fail(WAVESPEC_BAD_ARGS, "device_index out of range");
    WS_CUDA(cudaSetDevice(device_index), "cudaSetDevice");
    cudaDeviceProp prop;
    WS_CUDA(cudaGetDeviceProperties(&prop, device_index), "cudaGetDeviceProperties");
    if (prop.major < 10)
        return fail(WAVESPEC_BACKEND_UNAVAILABLE, "this library is built for sm_100a (B200) only");
    int n = stream_count < 1 ? 1 : (stream_count > 32 ? 32 : stream_count);   // more CUDA streams buy nothing
    g_s.streams.resize(n);
    for (int i = 0; i < n; i++) WS_CUDA(cudaStreamCreateWithFlags(&g_s.streams[i], cudaStreamNonBlocking), "cudaStreamCreate");
    WS_CUDA(cudaStreamCreateWithFlags(&g_s.side, cudaStreamNonBlocking), "cudaStreamCreate(side)");
    g_s.device = device_index;
    g_s.open = true;
    return WAVESPEC_OK;
}

void gpu_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_s.mu);
    if (!g_s.open) return;
    cudaSetDevice(g_s.device);
    cudaDeviceSynchronize();
    g_s.jobs.clear();
    g_s.band_scratch.clear();
    g_s.phase_scratch.clear();
    g_pool.trim();
    g_s.tw.clear(); g_s.win.clear(); g_s.apow.clear();
    for (auto s : g_s.streams) cudaStreamDestroy(s);
    if (g_s.side) { cudaStreamDestroy(g_s.side); g_s.side = nullptr; }
    g_s.streams.clear();
    g_s.open = false;
}

int32_t gpu_get_last_error_w(uint16_t* buf, int32_t buf_len) {
    if (!buf || buf_len <= 0) return 0;
    const std::string& m = t_last_error;
    int n = (int)m.size();
    if (n > buf_len - 1) n = buf_len - 1;
    for (int i = 0; i < n; i++) buf[i] = (uint16_t)(unsigned char)m[i];
    buf[n] = 0;
    return n + 1;     // code units written including the terminator (1.1.0 :743-744)
}

void wavespec_default_cfg(wavespec_pipeline_cfg* c, int32_t window_len) {
    std::memset(c, 0, sizeof *c);
    c->window_len = window_len; c->hop = 1; c->top_k = 8; c->row_stride = 15;
    c->min_period = 18; c->max_period = 200;     // Legacy/...-gpuopt-nodetrend.mq5:22-23
    c->sample_rate_seconds = 60.0;
    c->feed = WAVESPEC_FEED_CLOSE; c->detrend = WAVESPEC_DETREND_NONE; c->trend_period = 1024;  // ...-kalman-fast.mq5:804
    c->window_type = WAVESPEC_WINDOW_NONE; c->select = WAVESPEC_SELECT_INSERTION;
    c->pla_max_segments = 32; c->pla_max_error = 0.0005;                                          // :822-823
    c->outputs = WAVESPEC_OUT_SPECTRA | WAVESPEC_OUT_ROWS | WAVESPEC_OUT_BINS;
    c->wk_process_noise = 0.25; c->wk_meas_noise = 9.0; c->wk_init_variance = 25.0;               // 1.0.4-kalman.mq5:33-35
    wavespec_kalman4d_params& k = c->kalman;                                                      // ...-kalman-fast.mq5:885-901
    k.follow_strength = 1.0; k.q_pos = 0.01; k.q_vel = 0.003; k.q_acc = 0.0008; k.q_jerk = 0.0002;
    k.adapt_gain = 0.8; k.meas_noise = 1.0; k.init_var_pos = 16.0; k.init_var_vel = 9.0;
    k.init_var_acc = 4.0; k.init_var_jerk = 1.0; k.init_vel = 0.0; k.init_acc = 0.0; k.init_jerk = 0.0;
    k.clip_std = 6.0; k.ema_blend_period = 0.0;
    c->tracker_tolerance = 5.0; c->tracker_max_inactive = 3; c->reserved0 = 0;                    // :985-986
}

int64_t wavespec_num_windows(int32_t series_len, int32_t window_len, int32_t hop) {
    if (window_len < 1 || hop < 1 || series_len < window_len) return 0;
    return 1 + (int64_t)(series_len - window_len) / hop;
}

int32_t wavespec_pipeline_device(const double* d_series, int32_t n_series, int32_t series_len,
                                 const wavespec_pipeline_cfg* cfg, double* d_spectra, double* d_rows,
                                 int32_t* d_bins, double* d_waves, double* d_kalman, double* d_phase,
                                 double* d_wkalman, int32_t* d_trk_index, double* d_trk_period, void* stream) {
    int rc = ensure_open();
    if (rc) return rc;
    return run_pipeline(d_series, n_series, series_len, cfg, d_spectra, d_rows, d_bins, d_waves, d_kalman,
                        d_phase, d_wkalman, static_cast<cudaStream_t>(stream), d_trk_index, d_trk_period);
}

int32_t wavespec_pipeline_host(const double* series, int32_t n_series, int32_t series_len,
                               const wavespec_pipeline_cfg* cfg, double* spectra, double* rows,
                               int32_t* bins, double* waves, double* kalman, double* phase,
                               double* wkalman, int32_t* trk_index, double* trk_period) {
    int rc = ensure_open();
    if (rc) return rc;
    if ((rc = validate_cfg(cfg, series_len))) return rc;
    if (!series) return fail(WAVESPEC_BAD_ARGS, "series is null");
    const int N = cfg->window_len, K = cfg->top_k;
    const int64_t nwin = 1 + (int64_t)(series_len - N) / cfg->hop;
    const size_t tot = (size_t)n_series * nwin;
    cudaStream_t st = pick_stream();
    DeviceBuf ds, dsp, drw, dbn, dwv, dkl, dph, dwk, dti, dtp;
    WS_CUDA(ds.alloc((size_t)n_series * series_len * 8), "cudaMalloc(series)");
    if (spectra) WS_CUDA(dsp.alloc(tot * N * 8), "cudaMalloc(spectra)");
    if (rows)    WS_CUDA(drw.alloc(tot * K * cfg->row_stride * 8), "cudaMalloc(rows)");
    if (bins)    WS_CUDA(dbn.alloc(tot * K * 4), "cudaMalloc(bins)");
    if (waves)   WS_CUDA(dwv.alloc(tot * K * 8), "cudaMalloc(waves)");
    if (kalman)  WS_CUDA(dkl.alloc(tot * 8), "cudaMalloc(kalman)");
    if (phase)   WS_CUDA(dph.alloc(tot * 3 * (N / 2) * 8), "cudaMalloc(phase)");
    if (wkalman) WS_CUDA(dwk.alloc(tot * 8), "cudaMalloc(wkalman)");
    if (trk_index) WS_CUDA(dti.alloc(tot * 12 * 4), "cudaMalloc(tracker index)");
    if (trk_period) WS_CUDA(dtp.alloc(tot * 12 * 8), "cudaMalloc(tracker period)");
    WS_CUDA(cudaMemcpyAsync(ds.p, series, ds.bytes, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(series)");
    rc = run_pipeline(ds.as<double>(), n_series, series_len, cfg, dsp.as<double>(), drw.as<double>(),
                      dbn.as<int32_t>(), dwv.as<double>(), dkl.as<double>(), dph.as<double>(),
                      dwk.as<double>(), st, dti.as<int32_t>(), dtp.as<double>());
    if (rc) { cudaStreamSynchronize(st); return rc; }
    if (spectra) WS_CUDA(cudaMemcpyAsync(spectra, dsp.p, dsp.bytes, cudaMemcpyDeviceToHost, st), "D2H spectra");
    if (rows)    WS_CUDA(cudaMemcpyAsync(rows, drw.p, drw.bytes, cudaMemcpyDeviceToHost, st), "D2H rows");
    if (bins)    WS_CUDA(cudaMemcpyAsync(bins, dbn.p, dbn.bytes, cudaMemcpyDeviceToHost, st), "D2H bins");
    if (waves)   WS_CUDA(cudaMemcpyAsync(waves, dwv.p, dwv.bytes, cudaMemcpyDeviceToHost, st), "D2H waves");
    if (kalman)  WS_CUDA(cudaMemcpyAsync(kalman, dkl.p, dkl.bytes, cudaMemcpyDeviceToHost, st), "D2H kalman");
    if (phase)   WS_CUDA(cudaMemcpyAsync(phase, dph.p, dph.bytes, cudaMemcpyDeviceToHost, st), "D2H phase");
    if (wkalman) WS_CUDA(cudaMemcpyAsync(wkalman, dwk.p, dwk.bytes, cudaMemcpyDeviceToHost, st), "D2H wkalman");
    if (trk_index) WS_CUDA(cudaMemcpyAsync(trk_index, dti.p, dti.bytes, cudaMemcpyDeviceToHost, st), "D2H tracker index");
    if (trk_period) WS_CUDA(cudaMemcpyAsync(trk_period, dtp.p, dtp.bytes, cudaMemcpyDeviceToHost, st), "D2H tracker period");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return WAVESPEC_OK;
}

// ---- FFT entry points ---------------------------------------------------------------------------
static int fft_forward_common(const double* in, int32_t window_len, int32_t hop, int32_t series_len,
                              double* out) {
    int rc = ensure_open();
    if (rc) return rc;
    if (!in || !out) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    wavespec_pipeline_cfg c;
    wavespec_default_cfg(&c, window_len);
    c.hop = hop; c.outputs = WAVESPEC_OUT_SPECTRA;
    return wavespec_pipeline_host(in, 1, series_len, &c, out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int32_t gpu_fft_real_forward(const double* in, int32_t len, double* out) {
    if (!is_pow2(len) || len < 2) return fail(WAVESPEC_BAD_ARGS, "len must be a power of two >= 2");
    return fft_forward_common(in, len, 1, len, out);
}

int32_t gpu_fft_real_forward_batch(const double* in, int32_t window_len, int32_t n_windows, double* out) {
    if (!is_pow2(window_len) || window_len < 2) return fail(WAVESPEC_BAD_ARGS, "window_len must be a power of two >= 2");
    if (n_windows < 1) return fail(WAVESPEC_BAD_ARGS, "n_windows must be >= 1");
    if ((int64_t)window_len * n_windows > 0x7fffffff) return fail(WAVESPEC_BAD_ARGS, "batch too large for int32 lengths");
    return fft_forward_common(in, window_len, window_len, window_len * n_windows, out);
}

int32_t wavespec_fft_real_forward_sliding(const double* series, int32_t series_len, int32_t window_len,
                                          int32_t hop, double* out) {
    if (!is_pow2(window_len) || window_len < 2) return fail(WAVESPEC_BAD_ARGS, "window_len must be a power of two >= 2");
    if (hop < 1) return fail(WAVESPEC_BAD_ARGS, "hop must be >= 1");
    return fft_forward_common(series, window_len, hop, series_len, out);
}

int32_t gpu_fft_real_inverse(const double* in_spec, int32_t len, double* out) {
    int rc = ensure_open();
    if (rc) return rc;
    if (!in_spec || !out) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    if (!is_pow2(len) || len < 4 || len > 8192) return fail(WAVESPEC_BAD_ARGS, "len must be a power of two in [4, 8192]");
    const double2* tw;
    if ((rc = get_twiddles(len, &tw))) return rc;
    DeviceBuf din, dout;
    WS_CUDA(din.alloc((size_t)len * 8), "cudaMalloc");
    WS_CUDA(dout.alloc((size_t)len * 8), "cudaMalloc");
    cudaStream_t st = pick_stream();
    WS_CUDA(cudaMemcpyAsync(din.p, in_spec, (size_t)len * 8, cudaMemcpyHostToDevice, st), "H2D spectrum");
    WS_CUDA(ws::launch_inverse_real(din.as<double>(), len, 1, tw, dout.as<double>(), st), "inverse_real kernel");
    g_launches++;
    WS_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)len * 8, cudaMemcpyDeviceToHost, st), "D2H samples");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return WAVESPEC_OK;
}

// ---- cycle extraction ---------------------------------------------------------------------------
int32_t gpu_extract_cycles(const double* series, int32_t len, int32_t top_k, double min_period,
                           double max_period, double sample_rate_seconds, int32_t method, int32_t ar_order,
                           double* out, int32_t out_stride, int32_t out_capacity, int32_t* out_len) {
    (void)ar_order;
    if (out_len) *out_len = 0;
    int rc = ensure_open();
    if (rc) return rc;
    if ((rc = check_method(method))) return rc;
    if (!series || !out || !out_len) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    if (out_stride < 1 || out_capacity < 0) return fail(WAVESPEC_BAD_ARGS, "bad out_stride / out_capacity");
    wavespec_pipeline_cfg c;
    cfg_for_cycles(&c, len, 1, top_k, min_period, max_period, sample_rate_seconds, out_stride);
    if ((rc = validate_cfg(&c, len))) return rc;
    std::vector<double> rows((size_t)top_k * out_stride);
    rc = wavespec_pipeline_host(series, 1, len, &c, nullptr, rows.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    int n = top_k < out_capacity ? top_k : out_capacity;
    std::memcpy(out, rows.data(), (size_t)n * out_stride * 8);
    *out_len = n;
    return WAVESPEC_OK;
}

int32_t gpu_submit_extract_cycles(const double* series, int32_t len, int32_t top_k, double min_period,
                                  double max_period, double sample_rate_seconds, int32_t method,
                                  int32_t ar_order, int64_t* job_id) {
    (void)ar_order;
    int rc = check_method(method);
    if (rc) { if (job_id) *job_id = 0; return rc; }
    wavespec_pipeline_cfg c;
    cfg_for_cycles(&c, len, 1, top_k, min_period, max_period, sample_rate_seconds, 15);
    return submit_common(series, len, c, 0, job_id);
}

int32_t gpu_try_get_cycles(int64_t job_id, double* out, int32_t out_stride, int32_t out_capacity,
                           int32_t* out_len, int32_t* ready) {
    if (out_stride < 1 || out_capacity < 0) return fail(WAVESPEC_BAD_ARGS, "bad out_stride / out_capacity");
    return try_get_common(job_id, out, out_capacity, out_stride, 0, out_len, ready);
}

int32_t gpu_submit_extract_cycles_batch(const double* series, int32_t series_len, int32_t window_len,
                                        int32_t hop, int32_t top_k, double min_period, double max_period,
                                        double sample_rate_seconds, int32_t method, int32_t ar_order,
                                        int32_t stride, int64_t* job_id) {
    (void)ar_order;
    int rc = check_method(method);
    if (rc) { if (job_id) *job_id = 0; return rc; }
    wavespec_pipeline_cfg c;
    cfg_for_cycles(&c, window_len, hop, top_k, min_period, max_period, sample_rate_seconds, stride);
    return submit_common(series, series_len, c, 1, job_id);
}

int32_t gpu_try_get_cycles_batch(int64_t job_id, double* out, int32_t out_cap, int32_t* out_len,
                                 int32_t* ready) {
    if (out_cap < 0) return fail(WAVESPEC_BAD_ARGS, "bad out_cap");
    return try_get_common(job_id, out, out_cap, 0, 1, out_len, ready);
}

int32_t gpu_free_job(int64_t job_id) {
    std::shared_ptr<Job> job;
    {
        std::lock_guard<std::mutex> lk(g_s.mu);
        auto it = g_s.jobs.find(job_id);
        if (it == g_s.jobs.end()) return fail(WAVESPEC_BAD_ARGS, "unknown job id");
        job = it->second;
        g_s.jobs.erase(it);
    }
    // an in-flight job keeps its buffers until the device is done with them
    std::lock_guard<std::mutex> lk(job->mu);
    if (g_s.open) { cudaSetDevice(g_s.device); if (job->done) cudaEventSynchronize(job->done); }
    return WAVESPEC_OK;
}

int32_t wavespec_pla_windows_host(const double* series, int32_t series_len, int32_t window_len,
                                  int32_t hop, int32_t max_segments, double max_error, double* lines,
                                  int32_t* seg_bounds, int32_t* seg_counts) {
    int rc = ensure_open();
    if (rc) return rc;
    if (!series) return fail(WAVESPEC_BAD_ARGS, "series is null");
    if (window_len < 2 || hop < 1 || series_len < window_len) return fail(WAVESPEC_BAD_ARGS, "bad window/hop/series_len");
    const int64_t nwin = 1 + (int64_t)(series_len - window_len) / hop;
    const int cap = 2 * (max_segments < 1 ? 1 : max_segments) + 2;    // (start,end) pairs kept per window
    DeviceBuf ds, dl, db, dc;
    WS_CUDA(ds.alloc((size_t)series_len * 8), "cudaMalloc(series)");
    if (lines) WS_CUDA(dl.alloc((size_t)nwin * window_len * 8), "cudaMalloc(lines)");
    if (seg_bounds) WS_CUDA(db.alloc((size_t)nwin * cap * 2 * 4), "cudaMalloc(bounds)");
    if (seg_counts) WS_CUDA(dc.alloc((size_t)nwin * 4), "cudaMalloc(counts)");
    cudaStream_t st = pick_stream();
    WS_CUDA(cudaMemcpyAsync(ds.p, series, ds.bytes, cudaMemcpyHostToDevice, st), "H2D series");
    cudaError_t e = ws::launch_pla(ds.as<double>(), series_len, 1, nwin, window_len, hop, max_segments, max_error,
                                   dl.as<double>(), db.as<int32_t>(), dc.as<int32_t>(), cap, st);
    g_launches++;
    if (e != cudaSuccess) return cuda_fail(e, "pla kernel (recursion deeper than the on-chip stack?)");
    if (lines) WS_CUDA(cudaMemcpyAsync(lines, dl.p, dl.bytes, cudaMemcpyDeviceToHost, st), "D2H lines");
    if (seg_bounds) WS_CUDA(cudaMemcpyAsync(seg_bounds, db.p, db.bytes, cudaMemcpyDeviceToHost, st), "D2H bounds");
    if (seg_counts) WS_CUDA(cudaMemcpyAsync(seg_counts, dc.p, dc.bytes, cudaMemcpyDeviceToHost, st), "D2H counts");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return WAVESPEC_OK;
}

int32_t wavespec_zigzag_feed_host(const double* zz_main, const double* zz_high, const double* zz_low,
                                  int32_t series_len, int32_t window_len, int32_t hop, int32_t pivot_rule,
                                  int32_t mode, double fallback, int32_t min_pivots, double* lines,
                                  int32_t* valid) {
    int rc = ensure_open();
    if (rc) return rc;
    if (!zz_main || !zz_high || !zz_low || !lines) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    if (window_len < 1 || hop < 1 || series_len < window_len) return fail(WAVESPEC_BAD_ARGS, "bad window/hop/series_len");
    if (pivot_rule < 0 || pivot_rule > 1 || mode < 0 || mode > 2) return fail(WAVESPEC_BAD_ARGS, "bad pivot_rule / mode");
    const int64_t nwin = 1 + (int64_t)(series_len - window_len) / hop;
    const size_t sb = (size_t)series_len * 8;
    DeviceBuf dm, dh, dl, dfb, dpv, dprev, dnext, dlines, dvalid;
    WS_CUDA(dm.alloc(sb), "cudaMalloc"); WS_CUDA(dh.alloc(sb), "cudaMalloc"); WS_CUDA(dl.alloc(sb), "cudaMalloc");
    WS_CUDA(dfb.alloc(8), "cudaMalloc"); WS_CUDA(dpv.alloc(sb), "cudaMalloc");
    WS_CUDA(dprev.alloc((size_t)series_len * 4), "cudaMalloc"); WS_CUDA(dnext.alloc((size_t)series_len * 4), "cudaMalloc");
    WS_CUDA(dlines.alloc((size_t)nwin * window_len * 8), "cudaMalloc(lines)");
    if (valid) WS_CUDA(dvalid.alloc((size_t)nwin * 4), "cudaMalloc(valid)");
    cudaStream_t st = pick_stream();
    WS_CUDA(cudaMemcpyAsync(dm.p, zz_main, sb, cudaMemcpyHostToDevice, st), "H2D");
    WS_CUDA(cudaMemcpyAsync(dh.p, zz_high, sb, cudaMemcpyHostToDevice, st), "H2D");
    WS_CUDA(cudaMemcpyAsync(dl.p, zz_low, sb, cudaMemcpyHostToDevice, st), "H2D");
    WS_CUDA(cudaMemcpyAsync(dfb.p, &fallback, 8, cudaMemcpyHostToDevice, st), "H2D");
    WS_CUDA(ws::launch_zigzag(dm.as<double>(), dh.as<double>(), dl.as<double>(), dfb.as<double>(), 1, series_len,
                              window_len, hop, pivot_rule, mode, min_pivots, dpv.as<double>(), dprev.as<int32_t>(),
                              dnext.as<int32_t>(), dlines.as<double>(), dvalid.as<int32_t>(), st), "zigzag kernels");
    g_launches += 2;
    WS_CUDA(cudaMemcpyAsync(lines, dlines.p, dlines.bytes, cudaMemcpyDeviceToHost, st), "D2H lines");
    if (valid) WS_CUDA(cudaMemcpyAsync(valid, dvalid.p, dvalid.bytes, cudaMemcpyDeviceToHost, st), "D2H valid");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return WAVESPEC_OK;
}

static int applied_price_args(const double* o, const double* h, const double* l, const double* c, int64_t n,
                              int32_t mode, const double* out) {
    if (n < 1 || !out) return fail(WAVESPEC_BAD_ARGS, "n_bars must be >= 1 and out non-null");
    if (mode < WAVESPEC_PRICE_CLOSE || mode > WAVESPEC_PRICE_WEIGHTED) return fail(WAVESPEC_BAD_ARGS, "bad applied-price mode");
    const bool need_o = mode == WAVESPEC_PRICE_OPEN;
    const bool need_c = mode == WAVESPEC_PRICE_CLOSE || mode >= WAVESPEC_PRICE_TYPICAL;
    const bool need_h = mode == WAVESPEC_PRICE_HIGH || mode >= WAVESPEC_PRICE_MEDIAN;
    const bool need_l = mode == WAVESPEC_PRICE_LOW || mode >= WAVESPEC_PRICE_MEDIAN;
    if ((need_o && !o) || (need_c && !c) || (need_h && !h) || (need_l && !l))
        return fail(WAVESPEC_BAD_ARGS, "a price series this mode reads is null");
    return WAVESPEC_OK;
}

int32_t wavespec_applied_price_device(const double* d_open, const double* d_high, const double* d_low,
                                      const double* d_close, int64_t n_bars, int32_t mode, double* d_out,
                                      void* stream) {
    int rc = ensure_open();
    if (rc) return rc;
    if ((rc = applied_price_args(d_open, d_high, d_low, d_close, n_bars, mode, d_out))) return rc;
    WS_CUDA(ws::launch_applied_price(d_open, d_high, d_low, d_close, n_bars, mode, d_out,
                                     static_cast<cudaStream_t>(stream)), "applied_price kernel");
    g_launches++;
    g_last_kernel = "applied_price";
    return WAVESPEC_OK;
}

int32_t wavespec_applied_price_host(const double* open, const double* high, const double* low, const double* close,
                                    int64_t n_bars, int32_t mode, double* out) {
    int rc = ensure_open();
    if (rc) return rc;
    if ((rc = applied_price_args(open, high, low, close, n_bars, mode, out))) return rc;
    const size_t sb = (size_t)n_bars * 8;
    DeviceBuf dbuf[4], dout;
    const double* hsrc[4] = {open, high, low, close};
    cudaStream_t st = pick_stream();
    WS_CUDA(dout.alloc(sb), "cudaMalloc(applied price)");
    for (int i = 0; i < 4; i++) {
        if (!hsrc[i]) continue;
        WS_CUDA(dbuf[i].alloc(sb), "cudaMalloc(price series)");
        WS_CUDA(cudaMemcpyAsync(dbuf[i].p, hsrc[i], sb, cudaMemcpyHostToDevice, st), "H2D price series");
    }
    WS_CUDA(ws::launch_applied_price(dbuf[0].as<double>(), dbuf[1].as<double>(), dbuf[2].as<double>(),
                                     dbuf[3].as<double>(), n_bars, mode, dout.as<double>(), st), "applied_price kernel");
    g_launches++;
    g_last_kernel = "applied_price";
    WS_CUDA(cudaMemcpyAsync(out, dout.p, sb, cudaMemcpyDeviceToHost, st), "D2H applied price");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return WAVESPEC_OK;
}

int32_t wavespec_cycle_cache_host(const double* rows, int32_t n_windows, int32_t top_k, int32_t stride,
                                  int32_t window_len, int32_t hop, int32_t bars, double period_seconds,
                                  const wavespec_cache_params* params, double* out) {
    int rc = ensure_open();
    if (rc) return rc;
    if (!rows || !out || !params) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    if (n_windows < 1 || top_k < 1 || stride < 14 || window_len < 1 || hop < 1 || bars < 1)
        return fail(WAVESPEC_BAD_ARGS, "bad shape (stride must be >= 14)");
    DeviceBuf dr, dout;
    const size_t rb = (size_t)n_windows * top_k * stride * 8;
    WS_CUDA(dr.alloc(rb), "cudaMalloc(rows)");
    WS_CUDA(dout.alloc((size_t)bars * 20 * 8), "cudaMalloc(cache)");
    cudaStream_t st = pick_stream();
    WS_CUDA(cudaMemcpyAsync(dr.p, rows, rb, cudaMemcpyHostToDevice, st), "H2D rows");
    WS_CUDA(ws::launch_cycle_cache(dr.as<double>(), n_windows, top_k, stride, window_len, hop, bars, period_seconds,
                                   *params, dout.as<double>(), st), "cycle_cache kernel");
    g_launches++;
    WS_CUDA(cudaMemcpyAsync(out, dout.p, dout.bytes, cudaMemcpyDeviceToHost, st), "D2H cache");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return WAVESPEC_OK;
}

}  // extern "C"
