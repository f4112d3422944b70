// ws_pipeline.cu — dispatch of the per-bar pipeline onto the kernels.
//
// Everything here only enqueues work on the caller's stream: temporaries come from the device's
// stream-ordered pool (released in stream order, no synchronisation), per-stream scratch planes are
// kept between calls.  The one exception is the PLA feed, whose recursion-overflow flag has to be
// read back before the call can report success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <tuple>
#include <vector>

#include "ws_runtime.h"

namespace wsrt {

using ws::Params;

static const double kPi = 3.14159265358979323846;   // MQL5 M_PI

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

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

// per-stream scratch plane that only grows; work on one stream is ordered, so the next call may
// reuse it without waiting
static int stream_scratch(Device& dev, std::map<cudaStream_t, std::unique_ptr<DeviceBuf>>& slots, cudaStream_t st,
                          size_t bytes, void** out, const char* what) {
    std::lock_guard<std::mutex> lk(dev.mu);
    auto& slot = slots[st];
    if (!slot) slot = std::make_unique<DeviceBuf>();
    if (slot->bytes < bytes) {
        // growing: the old buffer may still be in use by queued kernels
        WS_CUDA(cudaStreamSynchronize(st), what);
        WS_CUDA(slot->alloc(bytes), what);
    }
    *out = slot->p;
    return WAVESPEC_OK;
}

// A13 host logic.  Which tracker a band bin matches, when trackers are appended, expire and shift
// (Legacy/...-kalman-fast.mq5:1418-1529) depends on PERIODS only — never on the data — and every bar
// presents the same bins in the same order.  So the tracker structure is one deterministic
// sequence shared by all series, and it settles: from some bar B0 on it repeats itself, after
// which the 12 slots (sticky, refilled only from unused trackers) cannot change any more.  This
// returns B0 (first bar whose end state equals the previous bar's), or -1 if no fixed point shows
// up within `limit` bars (then the device walks every bar).
static int64_t tracker_fixed_point(int N, int lo, int hi, double tol, int max_inactive, int64_t limit) {
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
static int run_tracker_path(Params p, const wavespec_pipeline_cfg* c, int32_t* d_trk_index, double* d_trk_period,
                            bool plain, cudaStream_t st) {
    const int band = p.band_hi - p.band_lo + 1;
    const int64_t nwin = p.nwin;
    const bool sel = p.rows || p.bins || p.waves || p.contrib;
    const bool other = sel || p.spectra || p.phase;
    static const bool walk_all = getenv("WAVESPEC_TRACKER_WALK_ALL") != nullptr;      // testing hook
    // the fixed point is a pure function of (N, band, tolerance, max_inactive): computed once per shape
    int64_t fixed = -1;
    if (!walk_all) {
        static std::mutex mu;
        static std::map<std::tuple<int, int, int, double, int>, int64_t> cache;
        const auto key = std::make_tuple(p.N, p.band_lo, p.band_hi, c->tracker_tolerance, c->tracker_max_inactive);
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it == cache.end())
            it = cache.emplace(key, tracker_fixed_point(p.N, p.band_lo, p.band_hi, c->tracker_tolerance,
                                                        c->tracker_max_inactive, 4096)).first;
        fixed = it->second;
    }
    // bars the sequential kernel has to walk: one past the fixed point (its slots are final)
    const int64_t walk = (fixed < 0 || fixed + 1 >= nwin) ? nwin : fixed + 1;
    const int64_t need = other ? nwin : walk;                 // windows the FFT kernels must cover
    const size_t per_win = (size_t)band * 16 * (size_t)p.n_series;
    int64_t wchunk = (int64_t)(((size_t)2 << 30) / per_win);
    if (wchunk < 1) wchunk = 1;
    if (wchunk > need) wchunk = need;
    AsyncBuf scratch, states;                                 // released in stream order when this returns
    WS_CUDA(scratch.alloc(per_win * (size_t)wchunk, st), "cudaMallocAsync(band buffer)");
    WS_CUDA(states.alloc(sizeof(ws::TrackerState) * (size_t)p.n_series, st), "cudaMallocAsync(tracker state)");
    // the sliding kernel hands the band over only in its split form (insertion rule, K <= 8)
    bool use_sliding = plain && ws::sliding_shared_supported(p) && (!sel || ws::rows_from_band_supported(p));
    const char* which = "sliding_shared";
    for (int64_t wa = 0; wa < need; wa += wchunk) {
        Params q = p;
        q.win_offset = wa; q.chunk_nwin = (wa + wchunk <= need) ? wchunk : need - wa;
        q.band_buf = scratch.as<double2>();
        if (use_sliding) {
            WS_CUDA(ws::launch_sliding_shared(q, st), "sliding_shared kernel");
            g_launches++;
            if (sel) { WS_CUDA(ws::launch_rows_from_band(q, st), "rows_from_band kernel"); g_launches++; }
        } else {
            WS_CUDA(ws::launch_window_fft(q, st, &which), "window_fft kernel");
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
    g_last_kernel = which;
    if (walk < nwin) {
        WS_CUDA(ws::launch_tracker_fill(p.n_series, nwin, walk - 1, d_trk_index, d_trk_period, st), "tracker fill kernel");
        g_launches++;
    }
    return WAVESPEC_OK;
}

int run_pipeline(Device& dev, const double* d_series, int32_t n_series, int32_t series_len,
                 const wavespec_pipeline_cfg* c, const Planes& out, cudaStream_t st,
                 int64_t w_begin, int64_t w_count) {
    int rc = validate_cfg(c, series_len);
    if (rc) return rc;
    if (n_series < 1 || n_series > 65535) return fail(WAVESPEC_BAD_ARGS, "n_series must be in [1, 65535]");
    if (!d_series) return fail(WAVESPEC_BAD_ARGS, "series is null");
    const int N = c->window_len;
    const int64_t nwin = 1 + (int64_t)(series_len - N) / c->hop;
    const bool ranged = w_count >= 0;
    if (!ranged) { w_begin = 0; w_count = nwin; }
    if (w_begin < 0 || w_count < 1 || w_begin + w_count > nwin) return fail(WAVESPEC_BAD_ARGS, "window range outside the series");
    const bool whole = w_begin == 0 && w_count == nwin;

    Params p;
    std::memset(&p, 0, sizeof p);
    p.series = d_series; p.series_stride = series_len; p.n_series = n_series; p.series_len = series_len;
    p.N = N; p.log2N = ilog2(N); p.hop = c->hop; p.K = c->top_k; p.row_stride = c->row_stride;
    p.nwin = nwin; p.win_offset = w_begin; p.chunk_nwin = w_count;
    p.spec_nwin = nwin; p.spec_w0 = 0;
    // band: Legacy/...-gpuopt-nodetrend.mq5:540-542
    int lo = (int)std::ceil((double)N / c->max_period);
    int hi = (int)std::floor((double)N / c->min_period);
    if (hi >= N / 2) hi = N / 2 - 1;
    if (lo < 0) lo = 0;
    p.band_lo = lo; p.band_hi = hi;
    p.detrend = c->detrend; p.select = c->select; p.sample_rate_seconds = c->sample_rate_seconds;
    if ((rc = dev.get_twiddles(N, &p.tw))) return rc;
    if ((rc = dev.get_window(N, c->window_type, &p.wtab))) return rc;
    p.has_window = p.wtab != nullptr;
    if (out.rows && (rc = dev.get_rowtab(N, &p.rowtab))) return rc;
    if (c->detrend == WAVESPEC_DETREND_IIR) {
        // Legacy/...-kalman-fast.mq5:3367-3369
        double omega = 2.0 * kPi / c->trend_period;
        double alpha = (1.0 - std::sin(omega)) / std::cos(omega);
        p.iir_alpha = alpha; p.iir_c = (1.0 - alpha) / 2.0;
        if ((rc = dev.get_apow(N, alpha, &p.apow))) return rc;
    }
    const bool want_wk = out.wkalman != nullptr;
    AsyncBuf tmp_contrib, tmp_bins, tmp_feed, tmp_z, tmp_flag;     // stream-ordered: die when this returns
    p.spectra = out.spectra; p.rows = out.rows; p.bins = out.bins; p.waves = out.waves; p.phase = out.phase;
    p.contrib = out.contrib;
    if (want_wk) {
        if (!p.contrib) {
            WS_CUDA(tmp_contrib.alloc((size_t)n_series * nwin * c->top_k * 8, st), "cudaMallocAsync(contrib)");
            p.contrib = tmp_contrib.as<double>();
        }
        if (!p.bins) {
            WS_CUDA(tmp_bins.alloc((size_t)n_series * nwin * c->top_k * 4, st), "cudaMallocAsync(bins)");
            p.bins = tmp_bins.as<int32_t>();
        }
    }
    int32_t* d_trk_index = out.trk_index;
    double* d_trk_period = out.trk_period;
    const bool want_trk = d_trk_index && d_trk_period;
    if ((d_trk_index != nullptr) != (d_trk_period != nullptr))
        return fail(WAVESPEC_BAD_ARGS, "tracker planes come as a pair (index and period)");
    if (want_trk && c->feed == WAVESPEC_FEED_PLA)
        return fail(WAVESPEC_BAD_ARGS, "the tracker plane is not wired to the PLA feed yet");
    if (want_trk && p.band_hi < p.band_lo) return fail(WAVESPEC_BAD_ARGS, "tracker needs a non-empty band");
    if (!whole && (want_trk || want_wk || out.kalman))
        return fail(WAVESPEC_BAD_ARGS, "window ranges serve the stateless planes only (no Kalman / tracker recursion)");
    const bool any_spectral = (p.spectra || p.rows || p.bins || p.waves || p.phase || p.contrib) && !want_trk;
    const char* which = "none";

    if (c->feed == WAVESPEC_FEED_PLA) {
        // PLA lines are window-private (the recursion restarts per window): build them chunk by
        // chunk into a bounded temporary and feed the per-window FFT kernel from it.
        // bytes of feed per chunk: a PLA thread owns one window for a long walk, so a launch needs
        // several hundred thousand windows to fill the device (180 GB of HBM: 8 GB is cheap)
        const size_t budget = (size_t)8 << 30;
        int64_t chunk = (int64_t)(budget / ((size_t)n_series * N * 8));
        if (chunk < 1) chunk = 1;
        if (chunk > w_count) chunk = w_count;
        WS_CUDA(tmp_feed.alloc((size_t)n_series * chunk * N * 8, st), "cudaMallocAsync(pla feed)");
        WS_CUDA(tmp_flag.alloc(sizeof(int32_t), st), "cudaMallocAsync(pla flag)");
        WS_CUDA(cudaMemsetAsync(tmp_flag.p, 0, sizeof(int32_t), st), "cudaMemsetAsync(pla flag)");
        if (out.kalman) WS_CUDA(tmp_z.alloc((size_t)n_series * nwin * 8, st), "cudaMallocAsync(kalman z)");
        for (int64_t wa = w_begin; wa < w_begin + w_count; wa += chunk) {
            const int64_t cn = (wa + chunk <= w_begin + w_count) ? chunk : w_begin + w_count - wa;
            WS_CUDA(ws::launch_pla(d_series + wa * c->hop, series_len, n_series, cn, N, c->hop,
                                   c->pla_max_segments, c->pla_max_error, tmp_feed.as<double>(),
                                   nullptr, nullptr, 0, tmp_flag.as<int32_t>(), st), "pla kernel");
            g_launches++;
            if (out.kalman) {
                // newest sample of each PLA line is the Kalman measurement (:3354-3360)
                WS_CUDA(ws::launch_gather_last(tmp_feed.as<double>(), n_series, cn, N, tmp_z.as<double>(), nwin, wa, st),
                        "gather kernel");
                g_launches++;
            }
            if (any_spectral) {
                Params q = p;
                q.feed = tmp_feed.as<double>(); q.win_offset = wa; q.chunk_nwin = cn;
                WS_CUDA(ws::launch_window_fft(q, st, &which), "window_fft kernel");
                g_last_kernel = which;
                g_launches++;
            }
        }
        if (out.kalman) {
            ws::KalmanParams kp;
            std::memcpy(&kp, &c->kalman, sizeof kp);
            WS_CUDA(ws::launch_kalman4d(tmp_z.as<double>(), nwin, 1, n_series, nwin, kp, out.kalman, st), "kalman4d kernel");
            g_launches++;
        }
        // the recursion-overflow flag decides whether the feed is valid: read it back
        size_t got = 0;
        int32_t* h_flag = static_cast<int32_t*>(dev.pinned.get(sizeof(int32_t), &got));
        if (!h_flag) return fail(WAVESPEC_NO_MEM, "pinned staging for the PLA flag");
        cudaError_t e = cudaMemcpyAsync(h_flag, tmp_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        const int32_t ov = *h_flag;
        dev.pinned.put(h_flag, got);
        if (e != cudaSuccess) return cuda_fail(e, "pla feed");
        if (ov) return fail(WAVESPEC_INTERNAL_ERROR, "pla kernel: recursion deeper than the on-chip stack");
    } else {
        // Kalman4D (A9) only reads the series: it is one thread per series and strictly sequential over
        // bars (0.7 s for 1M bars), so it is forked onto the device's side stream and runs beside the
        // FFT kernels of this call; `st` joins it at the end.
        cudaEvent_t kalman_join = nullptr;
        if (out.kalman) {
            ws::KalmanParams kp;
            std::memcpy(&kp, &c->kalman, sizeof kp);
            cudaStream_t ks = (dev.side && (any_spectral || want_trk)) ? dev.side : st;
            if (ks != st) {
                cudaEvent_t fork = nullptr;
                WS_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming), "cudaEventCreate");
                WS_CUDA(cudaEventRecord(fork, st), "cudaEventRecord(fork)");          // inputs are ready on st
                WS_CUDA(cudaStreamWaitEvent(ks, fork, 0), "cudaStreamWaitEvent(fork)");
                cudaEventDestroy(fork);
            }
            WS_CUDA(ws::launch_kalman4d(d_series + (N - 1), series_len, c->hop, n_series, nwin, kp, out.kalman, ks),
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
            // a handful of windows (the per-bar calls of the live loop: one window per call) shares
            // nothing: the per-window kernel transforms exactly those, the sliding kernel a whole tile
            // (gpu_fft_real_forward(1024): 57 -> 48 us; WAVESPEC_TINY overrides the threshold)
            static const long tiny = [] { const char* e = getenv("WAVESPEC_TINY"); return e ? atol(e) : 2L; }();
            const bool few = (int64_t)p.n_series * p.chunk_nwin <= tiny && !p.band_buf;
            if (plain && !few && ws::sliding_shared_supported(p)) {
                // Two ways to produce rows on this path: the fused in-kernel epilogue (default) or a
                // hand-off of the in-band bins to a separate full-occupancy rows kernel
                // (WAVESPEC_SPLIT=1).  Measured on B200 at N=1024 they are within 3 % of each
                // other (profiles/README.md); the fused form needs no scratch and one launch.
                static const bool split = getenv("WAVESPEC_SPLIT") != nullptr;
                if (split && whole && p.win_offset == 0 && p.chunk_nwin == nwin && ws::rows_from_band_supported(p)) {
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
                    int rc2 = stream_scratch(dev, dev.band_scratch, st, per_win * (size_t)wchunk * (size_t)sgroup, &scratch,
                                             "band buffer");
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
                    which = "sliding_shared";
                } else {
                    which = "sliding_shared";
                    const char* tf = getenv("WAVESPEC_TIMING_FILE");      // debug: phase stamps of the first 1024 CTAs
                    if (tf && p.spectra && p.rows) {
                        Params q = p;
                        DeviceBuf stamps;
                        WS_CUDA(stamps.alloc(1024 * 4 * 8), "cudaMalloc(stamps)");
                        WS_CUDA(cudaMemsetAsync(stamps.p, 0, stamps.bytes, st), "cudaMemsetAsync(stamps)");
                        q.dbg = stamps.as<long long>();
                        WS_CUDA(ws::launch_sliding_shared(q, st, &which), "sliding_shared kernel");
                        std::vector<long long> h(1024 * 4);
                        WS_CUDA(cudaMemcpyAsync(h.data(), stamps.p, stamps.bytes, cudaMemcpyDeviceToHost, st), "D2H stamps");
                        WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize(stamps)");
                        if (FILE* f = fopen(tf, "w")) {
                            for (int i = 0; i < 1024; i++) fprintf(f, "%lld %lld %lld %lld\n", h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
                            fclose(f);
                        }
                    } else {
                        WS_CUDA(ws::launch_sliding_shared(p, st, &which), "sliding_shared kernel");
                    }
                }
            } else {
                WS_CUDA(ws::launch_window_fft(p, st, &which), "window_fft kernel");
            }
            g_last_kernel = which;
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
                    WS_CUDA(ws::launch_phase_from_spectra(p.spectra, nwin, 0, n_series, w_begin, w_count, nwin, N, p.phase, st),
                            "phase_chain kernel");
                    g_launches++;
                } else {
                    int64_t chunk = (int64_t)(((size_t)2 << 30) / ((size_t)n_series * N * 8));
                    if (const char* e = getenv("WAVESPEC_PHASE_CHUNK")) { long v = atol(e); if (v > 0) chunk = v; }   // test hook
                    if (chunk < 1) chunk = 1;
                    if (chunk > w_count) chunk = w_count;
                    void* sp = nullptr;
                    if ((rc = stream_scratch(dev, dev.phase_scratch, st, (size_t)n_series * chunk * N * 8, &sp,
                                             "phase scratch spectra"))) return rc;
                    double* scratch_p = static_cast<double*>(sp);
                    for (int64_t wa = w_begin; wa < w_begin + w_count; wa += chunk) {
                        const int64_t cn = (wa + chunk <= w_begin + w_count) ? chunk : w_begin + w_count - wa;
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
                                   out.wkalman, st), "wkalman kernel");
        g_launches++;
    }
    return WAVESPEC_OK;
}

}  // namespace wsrt
