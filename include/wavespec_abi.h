/*
 * wavespec_abi.h — C ABI of libwavespec.so, the B200 (sm_100a) replacement for the
 * bridge DLL the WaveSpecZZ indicators import.
 *
 * Section 1 is the drop-in boundary: exactly the symbols MQL5 binds through
 *   #import "mt-bridge.dll"  (reference: Include/imports.mqh:5-21)
 * plus the two Legacy declarations that belong to the same hot path
 *   gpu_fft_real_inverse        (Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:27)
 *   gpu_fft_real_forward_batch  (Legacy/WaveSpecZZ_1.0.3-pla-batch.mq5:29).
 * MQL5 type mapping (MT5 is x64 only): int -> int32_t, long -> int64_t,
 * `const double &a[]` -> const double* (element 0 of the physical storage, oldest bar
 * first), `double &a[]` -> caller-allocated double*, `int &x` -> int32_t*,
 * `ushort &buf[]` -> uint16_t* (UTF-16 code units).
 *
 * Section 2 are new-build extensions (prefix wavespec_): the fused per-bar pipeline
 * with the prologue/epilogue options the Legacy indicators ran on the CPU, and
 * device-pointer entry points used by bench.py and the parity tests.  No torch types
 * cross this boundary; `stream` arguments are a cudaStream_t passed as void*.
 *
 * There is no CPU fallback anywhere behind this header: every compute entry point
 * returns WAVESPEC_BACKEND_UNAVAILABLE when no CUDA device can be opened.
 */
#ifndef WAVESPEC_ABI_H
#define WAVESPEC_ABI_H

#include <stdint.h>

#if defined(_WIN32)
#  define WAVESPEC_API __declspec(dllexport)
#else
#  define WAVESPEC_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: WaveCyclesBatchFetcher.mq5:15-21, Legacy/WaveSpecZZ_gpu_wip.mq5:263-269 */
enum {
    WAVESPEC_OK                  =  0,
    WAVESPEC_BAD_ARGS            = -1,
    WAVESPEC_BACKEND_UNAVAILABLE = -2,
    WAVESPEC_TIMEOUT             = -3,
    WAVESPEC_INTERNAL_ERROR      = -4,
    WAVESPEC_NOT_READY           = -5,
    WAVESPEC_NO_MEM              = -6
};

/* Result row layout (doubles), WaveSpecZZ_1.1.0-gpuopt.mq5:329 and decode at :1476-1490.
 * Older callers pass stride 8 (Legacy/WaveSpecZZ_1.0.4-old.mq5:787-798), 12
 * (Legacy/WaveSpecZZ_gpu_wip.mq5:123-137) or 4: the library writes min(stride,15) fields. */
enum {
    WAVESPEC_ROW_AMPLITUDE = 0, WAVESPEC_ROW_FREQ = 1, WAVESPEC_ROW_PERIOD = 2,
    WAVESPEC_ROW_PHASE = 3, WAVESPEC_ROW_ETA_BARS = 4, WAVESPEC_ROW_ETA_SECONDS = 5,
    WAVESPEC_ROW_ENERGY_RATIO = 6, WAVESPEC_ROW_COHERENCE = 7, WAVESPEC_ROW_SNR_DB = 8,
    WAVESPEC_ROW_RESIDUAL_POWER = 9, WAVESPEC_ROW_EIGEN_RATIO = 10, WAVESPEC_ROW_SCORE = 11,
    WAVESPEC_ROW_KALMAN_PRED = 12, WAVESPEC_ROW_ETA_CONFIDENCE = 13, WAVESPEC_ROW_METHOD = 14,
    WAVESPEC_ROW_FIELDS = 15
};

/* ------------------------------------------------------------------------------------
 * Section 1 — the imports.mqh boundary
 * ---------------------------------------------------------------------------------- */

/* imports.mqh:6.  Idempotent lazy session on CUDA device `device_index`; `stream_count`
 * (2 in Legacy, 16..512 in 1.1.0 :729-735, 64 in the Fetcher :105-106) sizes the stream pool.
 * Every opened device keeps its own session (streams, coefficient tables, memory pools, one host
 * worker thread).  device_index = -1 opens every device of the box; calling gpu_init again with
 * another index adds that device.  With several devices open, submitted jobs are bound to the
 * devices round robin (series shard over GPUs, no collective) and the synchronous
 * calls run on the first opened device. */
WAVESPEC_API int32_t gpu_init(int32_t device_index, int32_t stream_count);

/* imports.mqh:7.  Waits for in-flight jobs, frees every job and the session. */
WAVESPEC_API void gpu_shutdown(void);

/* imports.mqh:8.  Real -> half-complex forward FFT, synchronous.  `len` is a power of two
 * in [2, 8192]; out[2k] = Re X[k], out[2k+1] = Im X[k], k = 0..len/2-1 (Nyquist dropped),
 * unnormalised, X[k] = sum x[n] e^{-2 pi i k n / len}.  Contract pinned by
 * Legacy/WaveSpecZZ_1.0.4-new.mq5:3171-3194 vs :3208 (CPU FourierTransformManual). */
WAVESPEC_API int32_t gpu_fft_real_forward(const double* in, int32_t len, double* out);

/* imports.mqh:9-11.  One window -> up to top_k rows, strongest first.  out_capacity and
 * *out_len count ROWS; row r starts at out[r*out_stride].  method: 0 FFT ridge, -1 auto
 * (resolved to 0).  method 1 (MUSIC/ESPRIT) has no statement in the reference; it is served
 * by the FFT ridge extractor and rows are tagged method=0. */
WAVESPEC_API int32_t gpu_extract_cycles(const double* series, int32_t len, int32_t top_k,
                                        double min_period, double max_period,
                                        double sample_rate_seconds, int32_t method,
                                        int32_t ar_order, double* out, int32_t out_stride,
                                        int32_t out_capacity, int32_t* out_len);

/* imports.mqh:12-13.  Asynchronous form; `series` is copied before returning
 * (the caller reuses it on the next bar, WaveSpecZZ_1.1.0-gpuopt.mq5:1313-1339). */
WAVESPEC_API int32_t gpu_submit_extract_cycles(const double* series, int32_t len, int32_t top_k,
                                               double min_period, double max_period,
                                               double sample_rate_seconds, int32_t method,
                                               int32_t ar_order, int64_t* job_id);

/* imports.mqh:14.  Non-blocking poll: WAVESPEC_NOT_READY with *ready=0 while running
 * (the only "still running" answer the 1.1.0 wait loop :1342-1374 accepts), WAVESPEC_OK with
 * *ready=1 and rows copied when done.  Never waits for the device. */
WAVESPEC_API int32_t gpu_try_get_cycles(int64_t job_id, double* out, int32_t out_stride,
                                        int32_t out_capacity, int32_t* out_len, int32_t* ready);

/* imports.mqh:15-17.  Sliding extraction over a whole series:
 * nwin = 1 + (series_len - window_len)/hop; window w covers [w*hop, w*hop+window_len);
 * exactly top_k rows of `stride` doubles per window, zero rows when the band has fewer bins. */
WAVESPEC_API int32_t gpu_submit_extract_cycles_batch(const double* series, int32_t series_len,
                                                     int32_t window_len, int32_t hop, int32_t top_k,
                                                     double min_period, double max_period,
                                                     double sample_rate_seconds, int32_t method,
                                                     int32_t ar_order, int32_t stride,
                                                     int64_t* job_id);

/* imports.mqh:18.  out_cap counts DOUBLES, *out_len counts ROWS
 * (WaveSpecZZ_1.1.0-gpuopt.mq5:1016-1017, :1067-1089).  While the job runs the answer is WAVESPEC_OK
 * with *ready=0: WaveCyclesBatchFetcher.mq5:127-132 sleeps only on that answer (NOT_READY would make
 * it spin through its 4000 tries), and the 1.1.0 warm-up loop (:1029-1040) accepts it too.
 * The poll never waits for the device.  The first poll arms `out`: the job's result is copied into
 * it chunk by chunk while later chunks are still being computed (directly by DMA when `out` is
 * page-locked), so keep polling with the same buffer; *ready=1 means every row has landed.  A
 * batch is limited to 2^31-1 rows / doubles by the int32 arguments (wavespec_try_get_cycles_batch64
 * lifts that). */
WAVESPEC_API int32_t gpu_try_get_cycles_batch(int64_t job_id, double* out, int32_t out_cap,
                                              int32_t* out_len, int32_t* ready);

/* same with 64-bit capacity and row count */
WAVESPEC_API int32_t wavespec_try_get_cycles_batch64(int64_t job_id, double* out, int64_t out_cap,
                                                     int64_t* out_len, int32_t* ready);

/* imports.mqh:19.  Any job kind, finished or in flight (an in-flight job keeps running; its device
 * memory returns to the pool in stream order, result copies into `out` are waited for). */
WAVESPEC_API int32_t gpu_free_job(int64_t job_id);

/* imports.mqh:20.  Copies the calling thread's last error text as UTF-16; returns the number
 * of code units written INCLUDING the terminator (WaveSpecZZ_1.1.0-gpuopt.mq5:743-744), 0 if
 * buf_len <= 0. */
WAVESPEC_API int32_t gpu_get_last_error_w(uint16_t* buf, int32_t buf_len);

/* Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:27; used Legacy/WaveSpecZZ_1.0.4-core.mq5:426.
 * Inverse of gpu_fft_real_forward: in_spec holds len/2 interleaved bins (Nyquist taken as 0),
 * out[n] = (1/len) * sum over the Hermitian-extended spectrum.  The 1/len normalisation is a
 * design decision of this build (the reference does not evidence one): it makes
 * inverse(forward(x)) == x up to the dropped Nyquist bin. */
WAVESPEC_API int32_t gpu_fft_real_inverse(const double* in_spec, int32_t len, double* out);

/* Legacy/WaveSpecZZ_1.0.3-pla-batch.mq5:29.  n_windows contiguous, non-overlapping windows
 * in -> n_windows*window_len interleaved doubles out. */
WAVESPEC_API int32_t gpu_fft_real_forward_batch(const double* in, int32_t window_len,
                                                int32_t n_windows, double* out);

/* ------------------------------------------------------------------------------------
 * Section 2 — new-build extensions
 * ---------------------------------------------------------------------------------- */

/* Feed construction (reference rows A1/A11 of SURVEY.md section 8a) */
enum { WAVESPEC_FEED_CLOSE = 0, WAVESPEC_FEED_PLA = 1 };
/* Detrend (A2a/A2b/A2c) */
enum { WAVESPEC_DETREND_NONE = 0, WAVESPEC_DETREND_IIR = 1, WAVESPEC_DETREND_MEAN = 2 };
/* Window (A3); values 0..4 follow WINDOW_TYPE order of Legacy/...-kalman-fast.mq5:1158-1176,
 * 5 is the gpu_wip form `0.5 - 0.5*cos((2*pi*i)/(n-1))` (Legacy/WaveSpecZZ_gpu_wip.mq5:954). */
enum { WAVESPEC_WINDOW_NONE = 0, WAVESPEC_WINDOW_HANN = 1, WAVESPEC_WINDOW_HAMMING = 2,
       WAVESPEC_WINDOW_BLACKMAN = 3, WAVESPEC_WINDOW_BARTLETT = 4, WAVESPEC_WINDOW_HANN_WIP = 5 };
/* Top-K rule: insertion with strict '>' (Legacy/...-gpuopt-nodetrend.mq5:537-554) or the
 * swap-based selection sort of Legacy/WaveSpecZZ_1.0.4-kalman.mq5:143-180 (also forces k>=1). */
enum { WAVESPEC_SELECT_INSERTION = 0, WAVESPEC_SELECT_SORT = 1 };
/* Output planes of the per-bar pipeline */
enum {
    WAVESPEC_OUT_SPECTRA = 1,   /* nwin * window_len doubles, interleaved half spectrum  */
    WAVESPEC_OUT_ROWS    = 2,   /* nwin * top_k * row_stride doubles                    */
    WAVESPEC_OUT_BINS    = 4,   /* nwin * top_k int32 selected bins (-1 when absent)    */
    WAVESPEC_OUT_WAVES   = 8,   /* nwin * top_k doubles: A8a last-sample reconstruction */
    WAVESPEC_OUT_KALMAN  = 16,  /* nwin doubles: StepKalman4D on the newest window sample */
    WAVESPEC_OUT_PHASE   = 32,  /* nwin * 3 * window_len/2 doubles: phase, unwrapped, group delay */
    WAVESPEC_OUT_WKALMAN = 64,  /* nwin doubles: weight-Kalman blend (1.0.4-kalman.mq5:194-231) */
    WAVESPEC_OUT_TRACKER = 128, /* nwin * 12 int32 bins + nwin * 12 double periods: the stable slots of
                                   the period tracker pool (Legacy/...-kalman-fast.mq5:1415-1667)  */
    WAVESPEC_OUT_CONTRIB = 256  /* nwin * top_k doubles: A8b single-bin inverse DFT at the newest sample
                                   (Legacy/WaveSpecZZ_1.0.4-kalman.mq5:182-192), the weight-Kalman input */
};

/* Output planes of the *_planes pipeline entry points (NULL = not wanted); layouts as above. */
typedef struct wavespec_planes {
    double*  spectra;
    double*  rows;
    int32_t* bins;
    double*  waves;
    double*  contrib;
    double*  kalman;
    double*  phase;
    double*  wkalman;
    int32_t* trk_index;
    double*  trk_period;
} wavespec_planes;

/* Kalman4D parameters, defaults of Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:885-901 */
typedef struct wavespec_kalman4d_params {
    double follow_strength, q_pos, q_vel, q_acc, q_jerk, adapt_gain, meas_noise;
    double init_var_pos, init_var_vel, init_var_acc, init_var_jerk;
    double init_vel, init_acc, init_jerk, clip_std, ema_blend_period;
} wavespec_kalman4d_params;

typedef struct wavespec_pipeline_cfg {
    int32_t window_len;        /* power of two, 2..8192                                   */
    int32_t hop;               /* >= 1                                                     */
    int32_t top_k;             /* 1..32                                                    */
    int32_t row_stride;        /* >= 1; min(row_stride,15) fields written per row          */
    double  min_period, max_period, sample_rate_seconds;
    int32_t feed;              /* WAVESPEC_FEED_*                                          */
    int32_t detrend;           /* WAVESPEC_DETREND_*                                       */
    double  trend_period;      /* InpTrendPeriod (default 1024)                            */
    int32_t window_type;       /* WAVESPEC_WINDOW_*                                        */
    int32_t select;            /* WAVESPEC_SELECT_*                                        */
    int32_t pla_max_segments;  /* default 32                                               */
    int32_t outputs;           /* bitmask of WAVESPEC_OUT_*                                */
    double  pla_max_error;     /* default 0.0005                                           */
    double  wk_process_noise, wk_meas_noise, wk_init_variance; /* weight-Kalman Q, R, P0    */
    wavespec_kalman4d_params kalman;
    double  tracker_tolerance;     /* InpTrackerTolerance, % (default 5.0)                    */
    int32_t tracker_max_inactive;  /* InpMaxInactiveBars (default 3)                          */
    int32_t reserved0;
} wavespec_pipeline_cfg;

/* Fills cfg with the reference defaults for a given window length. */
WAVESPEC_API void wavespec_default_cfg(wavespec_pipeline_cfg* cfg, int32_t window_len);

/* Number of windows / bytes of each output plane for a (series_len, cfg) pair. */
WAVESPEC_API int64_t wavespec_num_windows(int32_t series_len, int32_t window_len, int32_t hop);

/* Host-buffer pipeline over a batch of equally long series (row-major [n_series][series_len]).
 * Any output pointer may be NULL when its WAVESPEC_OUT_ bit is clear.  Synchronous; copies in,
 * runs on the session streams, copies out. */
WAVESPEC_API int32_t wavespec_pipeline_host(const double* series, int32_t n_series,
                                            int32_t series_len, const wavespec_pipeline_cfg* cfg,
                                            double* spectra, double* rows, int32_t* bins,
                                            double* waves, double* kalman, double* phase,
                                            double* wkalman, int32_t* trk_index, double* trk_period);

/* Device-pointer pipeline (all pointers are device memory on the session's device); enqueues on
 * `stream` and returns without synchronising.  Used by bench.py (`value`) and the GPU tests. */
WAVESPEC_API int32_t wavespec_pipeline_device(const double* d_series, int32_t n_series,
                                              int32_t series_len, const wavespec_pipeline_cfg* cfg,
                                              double* d_spectra, double* d_rows, int32_t* d_bins,
                                              double* d_waves, double* d_kalman, double* d_phase,
                                              double* d_wkalman, int32_t* d_trk_index,
                                              double* d_trk_period, void* stream);

/* Same two calls with the planes in a struct (adds the A8b contribution plane). */
WAVESPEC_API int32_t wavespec_pipeline_host_planes(const double* series, int32_t n_series, int32_t series_len,
                                                   const wavespec_pipeline_cfg* cfg,
                                                   const wavespec_planes* planes);
WAVESPEC_API int32_t wavespec_pipeline_device_planes(const double* d_series, int32_t n_series,
                                                     int32_t series_len, const wavespec_pipeline_cfg* cfg,
                                                     const wavespec_planes* planes, void* stream);

/* Batched form of gpu_fft_real_inverse (Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:27, used
 * Legacy/WaveSpecZZ_1.0.4-core.mq5:426): n_windows half spectra of window_len interleaved doubles
 * each (the layout gpu_fft_real_forward_batch and the spectra plane produce) -> n_windows *
 * window_len real samples, 1/window_len normalised.  The device form enqueues on `stream`. */
WAVESPEC_API int32_t wavespec_fft_real_inverse_batch_host(const double* in_spec, int32_t window_len,
                                                          int32_t n_windows, double* out);
WAVESPEC_API int32_t wavespec_fft_real_inverse_batch_device(const double* d_spec, int32_t window_len,
                                                            int64_t n_windows, double* d_out, void* stream);

/* Inverse-FFT wave reconstruction of the selected cycles (north star; BASELINE config 4): the
 * inverse transform of each window's spectrum masked to its top_k selected bins (bins plane of the
 * pipeline, -1 = absent; the conjugate half is implied) -> n_windows * window_len samples, i.e. the
 * sum of the selected cycles over the whole window.  Its last sample equals the sum of the A8b
 * contributions (Legacy/WaveSpecZZ_1.0.4-kalman.mq5:182-192) of the window. */
WAVESPEC_API int32_t wavespec_reconstruct_topk_host(const double* spectra, const int32_t* bins,
                                                    int32_t window_len, int32_t top_k, int32_t n_windows,
                                                    double* out);
WAVESPEC_API int32_t wavespec_reconstruct_topk_device(const double* d_spectra, const int32_t* d_bins,
                                                      int32_t window_len, int32_t top_k, int64_t n_windows,
                                                      double* d_out, void* stream);

/* Sliding variant of gpu_fft_real_forward_batch (hop-spaced overlapping windows of one host
 * series) — named distinctly, as SURVEY.md section 8b requires. */
WAVESPEC_API int32_t wavespec_fft_real_forward_sliding(const double* series, int32_t series_len,
                                                       int32_t window_len, int32_t hop,
                                                       double* out);

/* PLA feed of one window (A11): line[] and the integer segment bounds (pivots). */
WAVESPEC_API int32_t wavespec_pla_windows_host(const double* series, int32_t series_len,
                                               int32_t window_len, int32_t hop,
                                               int32_t max_segments, double max_error,
                                               double* lines, int32_t* seg_bounds,
                                               int32_t* seg_counts);

/* Applied-price series (A1): the per-bar series the windows are cut from, for the price sources of
 * Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3308-3316 (FFT_APPLIED_PRICE_SOURCE, :807-818, values
 * of MQL5's ENUM_APPLIED_PRICE): close, open, high, low, median (high+low)/2, typical
 * (high+low+close)/3, weighted (high+low+2*close)/4 — evaluated per bar in the reference's operand
 * order, bit-identical to the MQL5 loop.  Unused inputs of a mode may be NULL. */
enum {
    WAVESPEC_PRICE_CLOSE = 1, WAVESPEC_PRICE_OPEN = 2, WAVESPEC_PRICE_HIGH = 3, WAVESPEC_PRICE_LOW = 4,
    WAVESPEC_PRICE_MEDIAN = 5, WAVESPEC_PRICE_TYPICAL = 6, WAVESPEC_PRICE_WEIGHTED = 7
};
WAVESPEC_API int32_t wavespec_applied_price_host(const double* open, const double* high, const double* low,
                                                 const double* close, int64_t n_bars, int32_t mode, double* out);
/* same on device pointers, asynchronous on `stream` (a cudaStream_t, NULL = default stream) */
WAVESPEC_API int32_t wavespec_applied_price_device(const double* d_open, const double* d_high, const double* d_low,
                                                   const double* d_close, int64_t n_bars, int32_t mode,
                                                   double* d_out, void* stream);

/* ZigZag pivot -> feed expansion of every window of one host series (A12).  zz_main / zz_high /
 * zz_low are the per-bar indicator buffers in chronological order (for 1.1.0, `main` is the
 * channel LoadWindow builds: buffer 0 where non-zero, else buffer 1 — WaveSpecZZ_1.1.0-gpuopt.mq5:
 * 371-384).  pivot_rule 0 = 1.1.0 (:393-451: pivot where main != 0), 1 = Legacy
 * (Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:237-262: main, else high, else low, finite).
 * mode 0 = STEP / ALTERNATING, 1 = INTERP / CONTINUOUS, 2 = MID ((high+low)/2).
 * fallback = value of a window without pivots ((high[0]+low[0])/2 of the chart in 1.1.0).
 * valid[w] (may be NULL) = 1 when the window has at least min_pivots pivots (Legacy skips bars
 * with fewer than 2, :305-306). */
WAVESPEC_API int32_t wavespec_zigzag_feed_host(const double* zz_main, const double* zz_high,
                                               const double* zz_low, int32_t series_len,
                                               int32_t window_len, int32_t hop, int32_t pivot_rule,
                                               int32_t mode, double fallback, int32_t min_pivots,
                                               double* lines, int32_t* valid);

/* Batch result -> per-bar cycle cache record, on the device.  Replaces the
 * indicator's O(rows x N) sine back-propagation loop (WaveSpecZZ_1.1.0-gpuopt.mq5:1067-1099) and
 * yields exactly what SaveCycleCache writes (:294-324): 20 doubles per bar,
 * [Wave1 Wave2 Period1 Period2 Eta1 Eta2 Phase1 Phase2 Energy1 Energy2 Coher1 Coher2 Snr1 Snr2
 *  Score1 Score2 Eigen1 Eigen2 EtaConf1 EtaConf2], EMPTY_VALUE (DBL_MAX) where the loop writes nothing.
 * `rows` is the buffer gpu_try_get_cycles_batch fills (n_windows * top_k rows of `stride` doubles,
 * stride >= 14); `bars` is the series length (`got`).  The inputs of the weighting are the
 * indicator's inputs (:71-77, :64). */
typedef struct wavespec_cache_params {
    int32_t music_only;          /* InpMusicOnly: rows with method != 1 are skipped              */
    int32_t use_music_weights;   /* InpUseMusicWeights                                           */
    double  min_coherence;       /* InpMinCoherence (0.05)                                       */
    double  min_score;           /* InpMinScore (0.01)                                           */
    double  min_snr_db;          /* InpMinSnrDb (-40)                                            */
} wavespec_cache_params;
WAVESPEC_API int32_t wavespec_cycle_cache_host(const double* rows, int32_t n_windows, int32_t top_k,
                                               int32_t stride, int32_t window_len, int32_t hop,
                                               int32_t bars, double period_seconds,
                                               const wavespec_cache_params* params, double* out);

/* The cycle cache record as a JOB PRODUCT: same arguments as
 * gpu_submit_extract_cycles_batch plus the indicator's weighting inputs; the sliding extraction and
 * the decode of WaveSpecZZ_1.1.0-gpuopt.mq5:1067-1099 both run on the device and only the record
 * SaveCycleCache writes (:294-324) crosses PCIe: 20 doubles per BAR (160 B) instead of top_k * 15
 * doubles per window.  Poll with wavespec_try_get_cycle_cache (out_cap in doubles, *out_bars in
 * bars; same non-blocking, chunked delivery as gpu_try_get_cycles_batch), release with gpu_free_job. */
WAVESPEC_API int32_t wavespec_submit_cycle_cache_batch(const double* series, int32_t series_len,
                                                       int32_t window_len, int32_t hop, int32_t top_k,
                                                       double min_period, double max_period,
                                                       double sample_rate_seconds, int32_t method,
                                                       int32_t ar_order, const wavespec_cache_params* params,
                                                       int64_t* job_id);
WAVESPEC_API int32_t wavespec_try_get_cycle_cache(int64_t job_id, double* out, int64_t out_cap,
                                                  int32_t* out_bars, int32_t* ready);

/* Devices opened by gpu_init, and the device a job was bound to (-1: unknown job). */
WAVESPEC_API int32_t wavespec_device_count(void);
WAVESPEC_API int32_t wavespec_job_device(int64_t job_id);

/* Number of kernels launched by this library since gpu_init (bench.py `gpu_launches`). */
WAVESPEC_API int64_t wavespec_launch_count(void);

/* Name of the kernel family the last pipeline call dispatched ("sliding_shared", "window_fft"). */
WAVESPEC_API const char* wavespec_last_kernel(void);

/* Library/ABI version: major*10000 + minor*100 + patch */
WAVESPEC_API int32_t wavespec_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WAVESPEC_ABI_H */
