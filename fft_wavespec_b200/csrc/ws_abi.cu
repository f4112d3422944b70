// ws_abi.cu — the extern "C" entry points of libwavespec.so.  Each one and the reference interface
// it replaces is documented in include/wavespec_abi.h (imports.mqh:5-21 and the two Legacy
// declarations).  The runtime behind them lives in ws_runtime.cu (devices), ws_jobs.cu (job table)
// and ws_pipeline.cu (kernel dispatch).
//
// No CPU fallback: when no CUDA device can be opened every compute call returns
// WAVESPEC_BACKEND_UNAVAILABLE and says why through gpu_get_last_error_w.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ws_runtime.h"

using namespace wsrt;

namespace {

bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

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

// Host-buffer pipeline: upload, run, download, all on one stream of `dev`; the device planes come
// from the stream-ordered pool.
int pipeline_host_impl(const double* series, int32_t n_series, int32_t series_len,
                       const wavespec_pipeline_cfg* cfg, const wavespec_planes& h) {
    Device* dev = primary_device();
    if (!dev) return WAVESPEC_BACKEND_UNAVAILABLE;
    int rc = validate_cfg(cfg, series_len);
    if (rc) return rc;
    if (!series) return fail(WAVESPEC_BAD_ARGS, "series is null");
    if (n_series < 1) return fail(WAVESPEC_BAD_ARGS, "n_series must be >= 1");
    DeviceGuard guard(dev->index);
    const int N = cfg->window_len, K = cfg->top_k;
    const int64_t nwin = 1 + (int64_t)(series_len - N) / cfg->hop;
    const size_t tot = (size_t)n_series * nwin;
    cudaStream_t st = dev->pick_stream();
    // Small calls (the per-bar calls of the 1.1.0 live loop: one window in, a few rows or one spectrum
    // out) skip the device planes altogether: the kernels read the series from, and write the results
    // to, one pinned host block that the device addresses directly (cudaHostAlloc memory is mapped under
    // UVA), so a call is two memcpy, the launches and one stream synchronisation — no allocation, no
    // copy commands.  Measured from Python on the bench box: gpu_fft_real_forward(1024) 47 -> 22 us,
    // gpu_extract_cycles 52 -> 27 us.  WAVESPEC_SMALL_CALL_BYTES=0 switches it off.
    {
        const size_t in_bytes = (size_t)n_series * series_len * 8;
        struct Sec { void* host; size_t bytes, off; };
        Sec sec[10] = {
            {h.spectra, tot * N * 8, 0}, {h.rows, tot * K * cfg->row_stride * 8, 0}, {h.bins, tot * K * 4, 0},
            {h.waves, tot * K * 8, 0}, {h.contrib, tot * K * 8, 0}, {h.kalman, tot * 8, 0},
            {h.phase, tot * 3 * (size_t)(N / 2) * 8, 0}, {h.wkalman, tot * 8, 0},
            {h.trk_index, tot * 12 * 4, 0}, {h.trk_period, tot * 12 * 8, 0}};
        size_t total = (in_bytes + 255) & ~(size_t)255;
        for (Sec& q : sec) if (q.host) { q.off = total; total += (q.bytes + 255) & ~(size_t)255; }
        static const size_t small_call = [] { const char* e = getenv("WAVESPEC_SMALL_CALL_BYTES"); return e ? (size_t)atoll(e) : (size_t)256 * 1024; }();
        if (total <= small_call) {
            size_t got = 0;
            char* pin = static_cast<char*>(dev->pinned.get(total, &got));
            if (pin) {
                std::memcpy(pin, series, in_bytes);
                auto at = [&](int i) -> void* { return sec[i].host ? pin + sec[i].off : nullptr; };
                Planes d;
                d.spectra = static_cast<double*>(at(0)); d.rows = static_cast<double*>(at(1)); d.bins = static_cast<int32_t*>(at(2));
                d.waves = static_cast<double*>(at(3)); d.contrib = static_cast<double*>(at(4)); d.kalman = static_cast<double*>(at(5));
                d.phase = static_cast<double*>(at(6)); d.wkalman = static_cast<double*>(at(7));
                d.trk_index = static_cast<int32_t*>(at(8)); d.trk_period = static_cast<double*>(at(9));
                rc = run_pipeline(*dev, reinterpret_cast<const double*>(pin), n_series, series_len, cfg, d, st);
                const cudaError_t e = cudaStreamSynchronize(st);
                if (rc == WAVESPEC_OK && e == cudaSuccess)
                    for (const Sec& q : sec) if (q.host) std::memcpy(q.host, pin + q.off, q.bytes);
                dev->pinned.put(pin, got);
                if (rc) { cudaGetLastError(); return rc; }
                WS_CUDA(e, "cudaStreamSynchronize");
                return WAVESPEC_OK;
            }
        }
    }
    AsyncBuf ds, dsp, drw, dbn, dwv, dct, dkl, dph, dwk, dti, dtp;
    WS_CUDA(ds.alloc((size_t)n_series * series_len * 8, st), "cudaMallocAsync(series)");
    if (h.spectra)    WS_CUDA(dsp.alloc(tot * N * 8, st), "cudaMallocAsync(spectra)");
    if (h.rows)       WS_CUDA(drw.alloc(tot * K * cfg->row_stride * 8, st), "cudaMallocAsync(rows)");
    if (h.bins)       WS_CUDA(dbn.alloc(tot * K * 4, st), "cudaMallocAsync(bins)");
    if (h.waves)      WS_CUDA(dwv.alloc(tot * K * 8, st), "cudaMallocAsync(waves)");
    if (h.contrib)    WS_CUDA(dct.alloc(tot * K * 8, st), "cudaMallocAsync(contrib)");
    if (h.kalman)     WS_CUDA(dkl.alloc(tot * 8, st), "cudaMallocAsync(kalman)");
    if (h.phase)      WS_CUDA(dph.alloc(tot * 3 * (N / 2) * 8, st), "cudaMallocAsync(phase)");
    if (h.wkalman)    WS_CUDA(dwk.alloc(tot * 8, st), "cudaMallocAsync(wkalman)");
    if (h.trk_index)  WS_CUDA(dti.alloc(tot * 12 * 4, st), "cudaMallocAsync(tracker index)");
    if (h.trk_period) WS_CUDA(dtp.alloc(tot * 12 * 8, st), "cudaMallocAsync(tracker period)");
    WS_CUDA(cudaMemcpyAsync(ds.p, series, ds.bytes, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(series)");
    Planes d;
    d.spectra = dsp.as<double>(); d.rows = drw.as<double>(); d.bins = dbn.as<int32_t>(); d.waves = dwv.as<double>();
    d.contrib = dct.as<double>(); d.kalman = dkl.as<double>(); d.phase = dph.as<double>(); d.wkalman = dwk.as<double>();
    d.trk_index = dti.as<int32_t>(); d.trk_period = dtp.as<double>();
    rc = run_pipeline(*dev, ds.as<double>(), n_series, series_len, cfg, d, st);
    if (rc) { cudaStreamSynchronize(st); cudaGetLastError(); return rc; }
    if (h.spectra)    WS_CUDA(cudaMemcpyAsync(h.spectra, dsp.p, dsp.bytes, cudaMemcpyDeviceToHost, st), "D2H spectra");
    if (h.rows)       WS_CUDA(cudaMemcpyAsync(h.rows, drw.p, drw.bytes, cudaMemcpyDeviceToHost, st), "D2H rows");
    if (h.bins)       WS_CUDA(cudaMemcpyAsync(h.bins, dbn.p, dbn.bytes, cudaMemcpyDeviceToHost, st), "D2H bins");
    if (h.waves)      WS_CUDA(cudaMemcpyAsync(h.waves, dwv.p, dwv.bytes, cudaMemcpyDeviceToHost, st), "D2H waves");
    if (h.contrib)    WS_CUDA(cudaMemcpyAsync(h.contrib, dct.p, dct.bytes, cudaMemcpyDeviceToHost, st), "D2H contrib");
    if (h.kalman)     WS_CUDA(cudaMemcpyAsync(h.kalman, dkl.p, dkl.bytes, cudaMemcpyDeviceToHost, st), "D2H kalman");
    if (h.phase)      WS_CUDA(cudaMemcpyAsync(h.phase, dph.p, dph.bytes, cudaMemcpyDeviceToHost, st), "D2H phase");
    if (h.wkalman)    WS_CUDA(cudaMemcpyAsync(h.wkalman, dwk.p, dwk.bytes, cudaMemcpyDeviceToHost, st), "D2H wkalman");
    if (h.trk_index)  WS_CUDA(cudaMemcpyAsync(h.trk_index, dti.p, dti.bytes, cudaMemcpyDeviceToHost, st), "D2H tracker index");
    if (h.trk_period) WS_CUDA(cudaMemcpyAsync(h.trk_period, dtp.p, dtp.bytes, cudaMemcpyDeviceToHost, st), "D2H tracker period");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return WAVESPEC_OK;
}

int pipeline_device_impl(const double* d_series, int32_t n_series, int32_t series_len,
                         const wavespec_pipeline_cfg* cfg, const wavespec_planes& o, void* stream) {
    Device* dev = device_of_pointer(d_series);
    if (!dev) return g_rt.devs.empty() ? WAVESPEC_BACKEND_UNAVAILABLE : WAVESPEC_BAD_ARGS;
    DeviceGuard guard(dev->index);
    Planes d;
    d.spectra = o.spectra; d.rows = o.rows; d.bins = o.bins; d.waves = o.waves; d.contrib = o.contrib;
    d.kalman = o.kalman; d.phase = o.phase; d.wkalman = o.wkalman; d.trk_index = o.trk_index; d.trk_period = o.trk_period;
    return run_pipeline(*dev, d_series, n_series, series_len, cfg, d, static_cast<cudaStream_t>(stream));
}

int fft_forward_common(const double* in, int32_t window_len, int32_t hop, int32_t series_len, double* out) {
    if (!in || !out) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    wavespec_pipeline_cfg c;
    wavespec_default_cfg(&c, window_len);
    c.hop = hop; c.outputs = WAVESPEC_OUT_SPECTRA;
    wavespec_planes h;
    std::memset(&h, 0, sizeof h);
    h.spectra = out;
    return pipeline_host_impl(in, 1, series_len, &c, h);
}

int applied_price_args(const double* o, const double* h, const double* l, const double* c, int64_t n,
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

int32_t narrow_len(int rc, int64_t n, int32_t* out_len) {
    if (rc == WAVESPEC_OK && n > 0x7fffffff) return fail(WAVESPEC_BAD_ARGS, "result count does not fit int32 out_len");
    if (out_len) *out_len = (int32_t)n;
    return rc;
}

}  // namespace

// =================================================================================================
extern "C" {

int32_t wavespec_version(void) { return 20000; }
int64_t wavespec_launch_count(void) { return g_launches.load(); }
const char* wavespec_last_kernel(void) { return g_last_kernel.load(); }

int32_t gpu_init(int32_t device_index, int32_t stream_count) { return open_device(device_index, stream_count); }

void gpu_shutdown(void) { close_all_devices(); }

int32_t wavespec_device_count(void) {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    return (int32_t)g_rt.devs.size();
}

int32_t wavespec_job_device(int64_t job_id) {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    auto it = g_rt.jobs.find(job_id);
    return it == g_rt.jobs.end() ? -1 : it->second->dev->index;
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
    wavespec_planes o;
    std::memset(&o, 0, sizeof o);
    o.spectra = d_spectra; o.rows = d_rows; o.bins = d_bins; o.waves = d_waves; o.kalman = d_kalman;
    o.phase = d_phase; o.wkalman = d_wkalman; o.trk_index = d_trk_index; o.trk_period = d_trk_period;
    return pipeline_device_impl(d_series, n_series, series_len, cfg, o, stream);
}

int32_t wavespec_pipeline_device_planes(const double* d_series, int32_t n_series, int32_t series_len,
                                        const wavespec_pipeline_cfg* cfg, const wavespec_planes* planes,
                                        void* stream) {
    if (!planes) return fail(WAVESPEC_BAD_ARGS, "planes is null");
    return pipeline_device_impl(d_series, n_series, series_len, cfg, *planes, stream);
}

int32_t wavespec_pipeline_host(const double* series, int32_t n_series, int32_t series_len,
                               const wavespec_pipeline_cfg* cfg, double* spectra, double* rows,
                               int32_t* bins, double* waves, double* kalman, double* phase,
                               double* wkalman, int32_t* trk_index, double* trk_period) {
    wavespec_planes h;
    std::memset(&h, 0, sizeof h);
    h.spectra = spectra; h.rows = rows; h.bins = bins; h.waves = waves; h.kalman = kalman; h.phase = phase;
    h.wkalman = wkalman; h.trk_index = trk_index; h.trk_period = trk_period;
    return pipeline_host_impl(series, n_series, series_len, cfg, h);
}

int32_t wavespec_pipeline_host_planes(const double* series, int32_t n_series, int32_t series_len,
                                      const wavespec_pipeline_cfg* cfg, const wavespec_planes* planes) {
    if (!planes) return fail(WAVESPEC_BAD_ARGS, "planes is null");
    return pipeline_host_impl(series, n_series, series_len, cfg, *planes);
}

// ---- FFT entry points ---------------------------------------------------------------------------
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

static int inverse_device_impl(const double* d_spec, const int32_t* d_bins, int32_t top_k, int32_t window_len,
                               int64_t n_windows, double* d_out, void* stream) {
    Device* dev = device_of_pointer(d_spec);
    if (!dev) return g_rt.devs.empty() ? WAVESPEC_BACKEND_UNAVAILABLE : WAVESPEC_BAD_ARGS;
    if (!d_spec || !d_out) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    if (!is_pow2(window_len) || window_len < 4 || window_len > 8192)
        return fail(WAVESPEC_BAD_ARGS, "window_len must be a power of two in [4, 8192]");
    if (n_windows < 1) return fail(WAVESPEC_BAD_ARGS, "n_windows must be >= 1");
    if (d_bins && (top_k < 1 || top_k > ws::kMaxTopK)) return fail(WAVESPEC_BAD_ARGS, "top_k must be in [1, 32]");
    DeviceGuard guard(dev->index);
    const double2* tw;
    int rc = dev->get_twiddles(window_len, &tw);
    if (rc) return rc;
    WS_CUDA(ws::launch_inverse_real(d_spec, window_len, n_windows, tw, d_out, static_cast<cudaStream_t>(stream),
                                    d_bins, top_k), "inverse_real kernel");
    g_launches++;
    g_last_kernel = "inverse_real_warp";
    return WAVESPEC_OK;
}

static int inverse_host_impl(const double* in_spec, const int32_t* bins, int32_t top_k, int32_t window_len,
                             int32_t n_windows, double* out) {
    Device* dev = primary_device();
    if (!dev) return WAVESPEC_BACKEND_UNAVAILABLE;
    if (!in_spec || !out) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    if (n_windows < 1) return fail(WAVESPEC_BAD_ARGS, "n_windows must be >= 1");
    if (!is_pow2(window_len) || window_len < 4 || window_len > 8192)
        return fail(WAVESPEC_BAD_ARGS, "len must be a power of two in [4, 8192]");
    DeviceGuard guard(dev->index);
    cudaStream_t st = dev->pick_stream();
    const size_t bytes = (size_t)window_len * n_windows * 8;
    AsyncBuf din, dout, dbins;
    WS_CUDA(din.alloc(bytes, st), "cudaMallocAsync");
    WS_CUDA(dout.alloc(bytes, st), "cudaMallocAsync");
    WS_CUDA(cudaMemcpyAsync(din.p, in_spec, bytes, cudaMemcpyHostToDevice, st), "H2D spectrum");
    if (bins) {
        WS_CUDA(dbins.alloc((size_t)n_windows * top_k * 4, st), "cudaMallocAsync(bins)");
        WS_CUDA(cudaMemcpyAsync(dbins.p, bins, dbins.bytes, cudaMemcpyHostToDevice, st), "H2D bins");
    }
    int rc = inverse_device_impl(din.as<double>(), dbins.as<int32_t>(), top_k, window_len, n_windows, dout.as<double>(), st);
    if (rc) { cudaStreamSynchronize(st); cudaGetLastError(); return rc; }
    WS_CUDA(cudaMemcpyAsync(out, dout.p, bytes, cudaMemcpyDeviceToHost, st), "D2H samples");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return WAVESPEC_OK;
}

int32_t wavespec_fft_real_inverse_batch_device(const double* d_spec, int32_t window_len, int64_t n_windows,
                                               double* d_out, void* stream) {
    return inverse_device_impl(d_spec, nullptr, 0, window_len, n_windows, d_out, stream);
}

int32_t wavespec_fft_real_inverse_batch_host(const double* in_spec, int32_t window_len, int32_t n_windows, double* out) {
    return inverse_host_impl(in_spec, nullptr, 0, window_len, n_windows, out);
}

int32_t wavespec_reconstruct_topk_device(const double* d_spectra, const int32_t* d_bins, int32_t window_len,
                                         int32_t top_k, int64_t n_windows, double* d_out, void* stream) {
    if (!d_bins) return fail(WAVESPEC_BAD_ARGS, "bins is null");
    return inverse_device_impl(d_spectra, d_bins, top_k, window_len, n_windows, d_out, stream);
}

int32_t wavespec_reconstruct_topk_host(const double* spectra, const int32_t* bins, int32_t window_len,
                                       int32_t top_k, int32_t n_windows, double* out) {
    if (!bins) return fail(WAVESPEC_BAD_ARGS, "bins is null");
    if (top_k < 1 || top_k > ws::kMaxTopK) return fail(WAVESPEC_BAD_ARGS, "top_k must be in [1, 32]");
    return inverse_host_impl(spectra, bins, top_k, window_len, n_windows, out);
}

int32_t gpu_fft_real_inverse(const double* in_spec, int32_t len, double* out) {
    return wavespec_fft_real_inverse_batch_host(in_spec, len, 1, out);
}

// ---- cycle extraction ---------------------------------------------------------------------------
int32_t gpu_extract_cycles(const double* series, int32_t len, int32_t top_k, double min_period,
                           double max_period, double sample_rate_seconds, int32_t method, int32_t ar_order,
                           double* out, int32_t out_stride, int32_t out_capacity, int32_t* out_len) {
    (void)ar_order;
    if (out_len) *out_len = 0;
    if (!primary_device()) return WAVESPEC_BACKEND_UNAVAILABLE;
    int rc = check_method(method);
    if (rc) return rc;
    if (!series || !out || !out_len) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    if (out_stride < 1 || out_capacity < 0) return fail(WAVESPEC_BAD_ARGS, "bad out_stride / out_capacity");
    wavespec_pipeline_cfg c;
    cfg_for_cycles(&c, len, 1, top_k, min_period, max_period, sample_rate_seconds, out_stride);
    if ((rc = validate_cfg(&c, len))) return rc;
    std::vector<double> rows((size_t)top_k * out_stride);
    wavespec_planes h;
    std::memset(&h, 0, sizeof h);
    h.rows = rows.data();
    rc = pipeline_host_impl(series, 1, len, &c, h);
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
    return submit_job(series, len, c, kJobWindow, nullptr, job_id);
}

int32_t gpu_try_get_cycles(int64_t job_id, double* out, int32_t out_stride, int32_t out_capacity,
                           int32_t* out_len, int32_t* ready) {
    if (out_len) *out_len = 0;
    if (ready) *ready = 0;
    if (out_stride < 1 || out_capacity < 0) return fail(WAVESPEC_BAD_ARGS, "bad out_stride / out_capacity");
    if (!out_len) return fail(WAVESPEC_BAD_ARGS, "null output pointer");
    int64_t n = 0;
    const int rc = try_get_job(job_id, out, out_capacity, out_stride, kJobWindow, &n, ready);
    return narrow_len(rc, n, out_len);
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
    return submit_job(series, series_len, c, kJobBatchRows, nullptr, job_id);
}

int32_t gpu_try_get_cycles_batch(int64_t job_id, double* out, int32_t out_cap, int32_t* out_len,
                                 int32_t* ready) {
    if (out_len) *out_len = 0;
    if (ready) *ready = 0;
    if (out_cap < 0) return fail(WAVESPEC_BAD_ARGS, "bad out_cap");
    if (!out_len) return fail(WAVESPEC_BAD_ARGS, "null output pointer");
    int64_t n = 0;
    const int rc = try_get_job(job_id, out, out_cap, 0, kJobBatchRows, &n, ready);
    return narrow_len(rc, n, out_len);
}

int32_t wavespec_try_get_cycles_batch64(int64_t job_id, double* out, int64_t out_cap, int64_t* out_len,
                                        int32_t* ready) {
    if (out_cap < 0) return fail(WAVESPEC_BAD_ARGS, "bad out_cap");
    return try_get_job(job_id, out, out_cap, 0, kJobBatchRows, out_len, ready);
}

int32_t gpu_free_job(int64_t job_id) { return free_job(job_id); }

int32_t wavespec_submit_cycle_cache_batch(const double* series, int32_t series_len, int32_t window_len,
                                          int32_t hop, int32_t top_k, double min_period, double max_period,
                                          double sample_rate_seconds, int32_t method, int32_t ar_order,
                                          const wavespec_cache_params* params, int64_t* job_id) {
    (void)ar_order;
    int rc = check_method(method);
    if (rc) { if (job_id) *job_id = 0; return rc; }
    if (!params) { if (job_id) *job_id = 0; return fail(WAVESPEC_BAD_ARGS, "params is null"); }
    wavespec_pipeline_cfg c;
    cfg_for_cycles(&c, window_len, hop, top_k, min_period, max_period, sample_rate_seconds, 15);
    return submit_job(series, series_len, c, kJobCacheRecord, params, job_id);
}

int32_t wavespec_try_get_cycle_cache(int64_t job_id, double* out, int64_t out_cap, int32_t* out_bars,
                                     int32_t* ready) {
    if (out_bars) *out_bars = 0;
    if (ready) *ready = 0;
    if (out_cap < 0) return fail(WAVESPEC_BAD_ARGS, "bad out_cap");
    if (!out_bars) return fail(WAVESPEC_BAD_ARGS, "null output pointer");
    int64_t n = 0;
    const int rc = try_get_job(job_id, out, out_cap, 0, kJobCacheRecord, &n, ready);
    return narrow_len(rc, n, out_bars);
}

// ---- feeds ---------------------------------------------------------------------------------------
int32_t wavespec_pla_windows_host(const double* series, int32_t series_len, int32_t window_len,
                                  int32_t hop, int32_t max_segments, double max_error, double* lines,
                                  int32_t* seg_bounds, int32_t* seg_counts) {
    Device* dev = primary_device();
    if (!dev) return WAVESPEC_BACKEND_UNAVAILABLE;
    if (!series) return fail(WAVESPEC_BAD_ARGS, "series is null");
    if (window_len < 2 || hop < 1 || series_len < window_len) return fail(WAVESPEC_BAD_ARGS, "bad window/hop/series_len");
    DeviceGuard guard(dev->index);
    const int64_t nwin = 1 + (int64_t)(series_len - window_len) / hop;
    const int cap = 2 * (max_segments < 1 ? 1 : max_segments) + 2;    // (start,end) pairs kept per window
    cudaStream_t st = dev->pick_stream();
    AsyncBuf ds, dl, db, dc, dflag;
    WS_CUDA(ds.alloc((size_t)series_len * 8, st), "cudaMallocAsync(series)");
    if (lines) WS_CUDA(dl.alloc((size_t)nwin * window_len * 8, st), "cudaMallocAsync(lines)");
    if (seg_bounds) WS_CUDA(db.alloc((size_t)nwin * cap * 2 * 4, st), "cudaMallocAsync(bounds)");
    if (seg_counts) WS_CUDA(dc.alloc((size_t)nwin * 4, st), "cudaMallocAsync(counts)");
    WS_CUDA(dflag.alloc(sizeof(int32_t), st), "cudaMallocAsync(flag)");
    WS_CUDA(cudaMemsetAsync(dflag.p, 0, sizeof(int32_t), st), "cudaMemsetAsync(flag)");
    WS_CUDA(cudaMemcpyAsync(ds.p, series, ds.bytes, cudaMemcpyHostToDevice, st), "H2D series");
    WS_CUDA(ws::launch_pla(ds.as<double>(), series_len, 1, nwin, window_len, hop, max_segments, max_error,
                           dl.as<double>(), db.as<int32_t>(), dc.as<int32_t>(), cap, dflag.as<int32_t>(), st), "pla kernel");
    g_launches++;
    int32_t ov = 0;
    if (lines) WS_CUDA(cudaMemcpyAsync(lines, dl.p, dl.bytes, cudaMemcpyDeviceToHost, st), "D2H lines");
    if (seg_bounds) WS_CUDA(cudaMemcpyAsync(seg_bounds, db.p, db.bytes, cudaMemcpyDeviceToHost, st), "D2H bounds");
    if (seg_counts) WS_CUDA(cudaMemcpyAsync(seg_counts, dc.p, dc.bytes, cudaMemcpyDeviceToHost, st), "D2H counts");
    WS_CUDA(cudaMemcpyAsync(&ov, dflag.p, sizeof ov, cudaMemcpyDeviceToHost, st), "D2H flag");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    if (ov) return fail(WAVESPEC_INTERNAL_ERROR, "pla kernel: recursion deeper than the on-chip stack");
    return WAVESPEC_OK;
}

int32_t wavespec_zigzag_feed_host(const double* zz_main, const double* zz_high, const double* zz_low,
                                  int32_t series_len, int32_t window_len, int32_t hop, int32_t pivot_rule,
                                  int32_t mode, double fallback, int32_t min_pivots, double* lines,
                                  int32_t* valid) {
    Device* dev = primary_device();
    if (!dev) return WAVESPEC_BACKEND_UNAVAILABLE;
    if (!zz_main || !zz_high || !zz_low || !lines) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    if (window_len < 1 || hop < 1 || series_len < window_len) return fail(WAVESPEC_BAD_ARGS, "bad window/hop/series_len");
    if (pivot_rule < 0 || pivot_rule > 1 || mode < 0 || mode > 2) return fail(WAVESPEC_BAD_ARGS, "bad pivot_rule / mode");
    DeviceGuard guard(dev->index);
    const int64_t nwin = 1 + (int64_t)(series_len - window_len) / hop;
    const size_t sb = (size_t)series_len * 8;
    cudaStream_t st = dev->pick_stream();
    AsyncBuf dm, dh, dl, dfb, dpv, dprev, dnext, dlines, dvalid;
    WS_CUDA(dm.alloc(sb, st), "cudaMallocAsync"); WS_CUDA(dh.alloc(sb, st), "cudaMallocAsync");
    WS_CUDA(dl.alloc(sb, st), "cudaMallocAsync"); WS_CUDA(dfb.alloc(8, st), "cudaMallocAsync");
    WS_CUDA(dpv.alloc(sb, st), "cudaMallocAsync");
    WS_CUDA(dprev.alloc((size_t)series_len * 4, st), "cudaMallocAsync");
    WS_CUDA(dnext.alloc((size_t)series_len * 4, st), "cudaMallocAsync");
    WS_CUDA(dlines.alloc((size_t)nwin * window_len * 8, st), "cudaMallocAsync(lines)");
    if (valid) WS_CUDA(dvalid.alloc((size_t)nwin * 4, st), "cudaMallocAsync(valid)");
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

int32_t wavespec_applied_price_device(const double* d_open, const double* d_high, const double* d_low,
                                      const double* d_close, int64_t n_bars, int32_t mode, double* d_out,
                                      void* stream) {
    Device* dev = device_of_pointer(d_out);
    if (!dev) return g_rt.devs.empty() ? WAVESPEC_BACKEND_UNAVAILABLE : WAVESPEC_BAD_ARGS;
    int rc = applied_price_args(d_open, d_high, d_low, d_close, n_bars, mode, d_out);
    if (rc) return rc;
    DeviceGuard guard(dev->index);
    WS_CUDA(ws::launch_applied_price(d_open, d_high, d_low, d_close, n_bars, mode, d_out,
                                     static_cast<cudaStream_t>(stream)), "applied_price kernel");
    g_launches++;
    g_last_kernel = "applied_price";
    return WAVESPEC_OK;
}

int32_t wavespec_applied_price_host(const double* open, const double* high, const double* low, const double* close,
                                    int64_t n_bars, int32_t mode, double* out) {
    Device* dev = primary_device();
    if (!dev) return WAVESPEC_BACKEND_UNAVAILABLE;
    int rc = applied_price_args(open, high, low, close, n_bars, mode, out);
    if (rc) return rc;
    DeviceGuard guard(dev->index);
    const size_t sb = (size_t)n_bars * 8;
    AsyncBuf dbuf[4], dout;
    const double* hsrc[4] = {open, high, low, close};
    cudaStream_t st = dev->pick_stream();
    WS_CUDA(dout.alloc(sb, st), "cudaMallocAsync(applied price)");
    for (int i = 0; i < 4; i++) {
        if (!hsrc[i]) continue;
        WS_CUDA(dbuf[i].alloc(sb, st), "cudaMallocAsync(price series)");
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
    Device* dev = primary_device();
    if (!dev) return WAVESPEC_BACKEND_UNAVAILABLE;
    if (!rows || !out || !params) return fail(WAVESPEC_BAD_ARGS, "null buffer");
    if (n_windows < 1 || top_k < 1 || stride < 14 || window_len < 1 || hop < 1 || bars < 1)
        return fail(WAVESPEC_BAD_ARGS, "bad shape (stride must be >= 14)");
    DeviceGuard guard(dev->index);
    cudaStream_t st = dev->pick_stream();
    AsyncBuf dr, dout;
    const size_t rb = (size_t)n_windows * top_k * stride * 8;
    WS_CUDA(dr.alloc(rb, st), "cudaMallocAsync(rows)");
    WS_CUDA(dout.alloc((size_t)bars * 20 * 8, st), "cudaMallocAsync(cache)");
    WS_CUDA(cudaMemcpyAsync(dr.p, rows, rb, cudaMemcpyHostToDevice, st), "H2D rows");
    WS_CUDA(ws::launch_cycle_cache(dr.as<double>(), n_windows, top_k, stride, window_len, hop, bars, 0, bars,
                                   period_seconds, *params, dout.as<double>(), st), "cycle_cache kernel");
    g_launches++;
    WS_CUDA(cudaMemcpyAsync(out, dout.p, dout.bytes, cudaMemcpyDeviceToHost, st), "D2H cache");
    WS_CUDA(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return WAVESPEC_OK;
}

}  // extern "C"
