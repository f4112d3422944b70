/*
 * wavespec_oracle.h — CPU restatement of the reference's own arithmetic for the sliding
 * spectral hot path.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Nothing under fft_wavespec_b200/
 * may link, import or execute it.
 *
 * PARITY UNPINNED: the reference (MQL5 only) ships no tests, golden vectors or fixtures and
 * cannot be compiled or run here (no MetaEditor/MT5), so this transcription is pinned only by
 * analytic known answers (tests/test_oracle.py) and by the committed vectors it generated
 * itself (tests/golden/).  Transcendentals are glibc's, not the MT5 CRT's.
 *
 * Every function cites the reference lines it follows.  R/ = /root/reference/,
 * L/ = /root/reference/Legacy/.  Build: -O2 -ffp-contract=off -fno-fast-math (see Makefile).
 */
#ifndef WAVESPEC_ORACLE_H
#define WAVESPEC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* A4  L/WaveSpecZZ_1.0.2.mq5:938-975 FourierTransformManual: full complex radix-2 DIT on real
 * input, twiddles by recurrence.  re/im have n entries (all n bins, mirror half included). */
void oracle_fft_forward(const double* data, int n, double* re, double* im);

/* A4' what the bridge DLL hands back for the same input (L/...-kalman-fast.mq5:3422-3433):
 * out[2k]=re[k], out[2k+1]=im[k], k<n/2. */
void oracle_fft_interleaved(const double* data, int n, double* out);
/* A8d inverse of that contract: n/2 interleaved bins (Nyquist taken as 0) -> n real samples, 1/n
 * normalised; conj -> FourierTransformManual's butterfly loop -> conj (L/WaveSpecZZ_1.0.2.mq5:943-974;
 * signature L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:27, use L/WaveSpecZZ_1.0.4-core.mq5:426). */
void oracle_fft_inverse(const double* spec, int n, double* out);

/* A3  L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:1126-1177; type 0 none,1 Hann,2 Hamming,3 Blackman,
 * 4 Bartlett; 5 = Hann as written in L/WaveSpecZZ_gpu_wip.mq5:954. In place. */
void oracle_apply_window(double* data, int n, int type);

/* A2a L/...-kalman-fast.mq5:3367-3379 one-pole trend IIR, restarted per window. */
void oracle_detrend_iir(const double* price, int n, double trend_period, double* trend,
                        double* detrended);

/* A2b L/WaveSpecZZ_gpu_wip.mq5:935-957 mean removal fused with Hann. */
void oracle_mean_hann(const double* price, int n, double* out);

/* A5  L/...-kalman-fast.mq5:3441-3445 power of bins 0..n/2-1. */
void oracle_power(const double* re, const double* im, int n, double* spectrum);

/* A6  L/...-kalman-fast.mq5:1183-1263 phase / unwrap / group delay over `count` bins. */
void oracle_phase_chain(const double* re, const double* im, int count, double* phase,
                        double* unwrapped, double* group_delay);

/* A7a L/...-gpuopt-nodetrend.mq5:537-554 insertion top-K (reference K=8) with strict '>'.
 * top_bin[s] = -1 / top_pow[s] = -1.0 for unused slots. */
void oracle_topk_insertion(const double* spectrum, int n, double min_period, double max_period,
                           int top_k, int* top_bin, double* top_pow);

/* A7b L/WaveSpecZZ_1.0.4-kalman.mq5:143-180 candidate collection + swap selection sort.
 * idx/pow must hold n/2 entries; returns candidate count (sorted descending). */
int oracle_collect_sorted(const double* re, const double* im, int n, double min_period,
                          double max_period, int* idx, double* pow);

/* A8a L/...-gpuopt-nodetrend.mq5:559-568 last-sample reconstruction (cos form) + period. */
void oracle_recon_last(const double* re, const double* im, int n, int bin, double power,
                       double* wave, double* period);

/* A8b L/WaveSpecZZ_1.0.4-kalman.mq5:182-192 single-bin inverse DFT at sample n-1. */
double oracle_contribution(const double* re, const double* im, int n, int k);

/* A9  L/...-kalman-fast.mq5:2015-2125, defaults :885-901. */
typedef struct oracle_kalman4d_params {
    double follow_strength, q_pos, q_vel, q_acc, q_jerk, adapt_gain, meas_noise;
    double init_var_pos, init_var_vel, init_var_acc, init_var_jerk;
    double init_vel, init_acc, init_jerk, clip_std, ema_blend_period;
} oracle_kalman4d_params;
typedef struct oracle_kalman4d_state {
    double pos, vel, acc, jerk;
    double P[4][4];
    int ready;
    double ema_prev;
    int ema_ready;
} oracle_kalman4d_state;
void oracle_kalman4d_defaults(oracle_kalman4d_params* p);
void oracle_kalman4d_reset(oracle_kalman4d_state* s, const oracle_kalman4d_params* p, double first);
double oracle_kalman4d_step(oracle_kalman4d_state* s, const oracle_kalman4d_params* p, double z);
/* bar loop order of :3354-3360: reset on the first processed bar, then step that same bar. */
void oracle_kalman4d_series(const double* z, int count, const oracle_kalman4d_params* p, double* out);

/* A10 L/WaveSpecZZ_1.0.4-kalman.mq5:194-231 (= L/WaveSpecZZ_1.0.4-old.mq5:2606-2649). */
typedef struct oracle_wkalman_state { double weights[32]; double cov[32]; } oracle_wkalman_state;
void oracle_wkalman_reset(oracle_wkalman_state* s, double init_variance);
double oracle_wkalman_update(oracle_wkalman_state* s, const double* cycle_vals, int cycle_count,
                             double measurement, double q, double r);

/* A11 L/...-kalman-fast.mq5:387-502 PLA.  window -> line; seg_* arrays need n
 * entries; returns the segment count. */
int oracle_pla_build(const double* window, int n, int max_segments, double max_error, double* line,
                     int* seg_starts, int* seg_ends, double* seg_slopes, double* seg_intercepts);

/* A12 R/WaveSpecZZ_1.1.0-gpuopt.mq5:393-451 ZigZag pivots -> feed; mode 0 STEP,1 INTERP,2 MID.
 * main/high/low are the chronological channels LoadWindow (:361-391) produces. */
void oracle_zigzag_feed_110(const double* main_ch, const double* high_ch, const double* low_ch,
                            int len, int mode, double high0, double low0, double* feed);
/* A12 L/...-kalman-fast.mq5:237-357 (current-timeframe branch); mode 0 CONTINUOUS, 1 ALTERNATING.
 * returns 0 when fewer than two pivots (the reference skips the bar). */
/* Applied price (A1): Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3308-3316; mode = MQL5
 * ENUM_APPLIED_PRICE value (1 close, 2 open, 3 high, 4 low, 5 median, 6 typical, 7 weighted).
 * Returns 0, or -1 for an unknown mode. */
int oracle_applied_price(const double* open, const double* high, const double* low, const double* close,
                         long long n, int mode, double* out);

int oracle_zigzag_series_legacy(const double* zz_main, const double* zz_high, const double* zz_low,
                                int n, int mode, double* price_data);

/* A13 L/...-kalman-fast.mq5:1415-1667 persistent period tracker pool + 12 stable slots, driven per
 * bar as in :3450-3504: candidates = every band bin ascending (period = N / bin), match against the
 * closest ACTIVE tracker within `tolerance_pct` else append, deactivate / erase unseen trackers
 * (erase shifts the array; slot indices are NOT remapped — reproduced), refill free slots by
 * descending power (bubble sort with '<', stable). */
#define ORACLE_TRACKER_CAP 512
typedef struct oracle_tracker_state {
    int count;
    double period[ORACLE_TRACKER_CAP];
    double power[ORACLE_TRACKER_CAP];
    int fft_index[ORACLE_TRACKER_CAP];
    int is_active[ORACLE_TRACKER_CAP];
    int bars_inactive[ORACLE_TRACKER_CAP];
    int slot[12];
} oracle_tracker_state;
void oracle_tracker_reset(oracle_tracker_state* st);
/* spectrum: n/2 powers of this bar; writes the 12 slots (index 0 / period 0.0 when empty). */
void oracle_tracker_step(oracle_tracker_state* st, const double* spectrum, int n, double min_period,
                         double max_period, double tolerance_pct, int max_inactive_bars,
                         int32_t* slot_index, double* slot_period);

/* A8c + cache record: R/WaveSpecZZ_1.1.0-gpuopt.mq5:1044-1099 (buffers initialised to EMPTY_VALUE,
 * sine back-propagation of every row over up to N bars, later rows overwrite) followed by the
 * 20-doubles-per-bar record of SaveCycleCache (:294-324).  out has bars * 20 doubles. */
void oracle_cycle_cache(const double* rows, int out_len, int stride, int top_k, int window_len, int hop,
                        int bars, double period_seconds, int music_only, int use_music_weights,
                        double min_coherence, double min_score, double min_snr_db, double* out);

/* ---- per-series pipelines (bar loops), used by parity tests and the timed CPU baseline ---- */
typedef struct oracle_pipeline_cfg {
    int window_len, hop, top_k, row_stride;
    double min_period, max_period, sample_rate_seconds;
    int feed, detrend;
    double trend_period;
    int window_type, select, pla_max_segments, outputs;
    double pla_max_error;
    double wk_process_noise, wk_meas_noise, wk_init_variance;
    oracle_kalman4d_params kalman;
    double tracker_tolerance;      /* InpTrackerTolerance, default 5.0 (%)  */
    int tracker_max_inactive;      /* InpMaxInactiveBars, default 3          */
    int reserved0;
} oracle_pipeline_cfg;   /* field-for-field the same as wavespec_pipeline_cfg */

void oracle_default_cfg(oracle_pipeline_cfg* cfg, int window_len);

/* One series.  Output planes as in include/wavespec_abi.h (NULL to skip).  Rows follow the
 * new-build definition documented in DESIGN.md (amplitude = 2*sqrt(P)/N, phase such that
 * amplitude*sin(phase) equals A8b). */
void oracle_pipeline_series(const double* series, int series_len, const oracle_pipeline_cfg* cfg,
                            double* spectra, double* rows, int32_t* bins, double* waves,
                            double* kalman, double* phase, double* wkalman);
/* same, plus the tracker planes: trk_index [nwin][12] int32, trk_period [nwin][12] double */
void oracle_pipeline_series_trk(const double* series, int series_len, const oracle_pipeline_cfg* cfg,
                                double* spectra, double* rows, int32_t* bins, double* waves,
                                double* kalman, double* phase, double* wkalman,
                                int32_t* trk_index, double* trk_period);

/* n_series series (row-major), split over `threads` std::threads by series then by bar range;
 * only the stateless planes (spectra/rows/bins/waves) may be requested with bar-range splits.
 * max_windows > 0 limits each series to its first max_windows windows (bounded CPU sample).
 * Returns the number of windows processed. */
int64_t oracle_pipeline_batch_mt(const double* series, int n_series, int series_len,
                                 const oracle_pipeline_cfg* cfg, int threads, int64_t max_windows,
                                 double* spectra, double* rows, int32_t* bins, double* waves);

int oracle_version(void);

#ifdef __cplusplus
}
#endif
#endif
