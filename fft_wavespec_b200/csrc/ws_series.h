// ws_series.h — host-visible declarations of the kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wavespec_abi.h"
#include "ws_common.cuh"

namespace ws {

// Same field order as wavespec_kalman4d_params (include/wavespec_abi.h).
struct KalmanParams {
    double follow_strength, q_pos, q_vel, q_acc, q_jerk, adapt_gain, meas_noise;
    double init_var_pos, init_var_vel, init_var_acc, init_var_jerk;
    double init_vel, init_acc, init_jerk, clip_std, ema_blend_period;
};

// A13 tracker pool state of one series (ws_series.cu)
constexpr int kTrackerCap = 512;
struct TrackerState {
    int32_t count;
    int32_t slot[12];
    double period[kTrackerCap];
    double power[kTrackerCap];
    int32_t fft_index[kTrackerCap];
    int32_t is_active[kTrackerCap];
    int32_t bars_inactive[kTrackerCap];
};
cudaError_t launch_tracker(const double2* band, int32_t band_lo, int32_t nband, int32_t n_series,
                           int64_t chunk_nwin, int64_t n_process, int64_t win_offset, int64_t nwin, int32_t N,
                           double tol, int32_t max_inactive, TrackerState* states, int32_t* trk_index,
                           double* trk_period, cudaStream_t stream);
cudaError_t launch_tracker_fill(int32_t n_series, int64_t nwin, int64_t last, int32_t* trk_index,
                                double* trk_period, cudaStream_t stream);

// ws_window_fft.cu: dispatches to the warp-per-window kernel when it serves the request; the name of
// the kernel that ran ("window_fft" or "window_fft_warp") is returned through *which when non-null
cudaError_t launch_window_fft(Params p, cudaStream_t stream, const char** which = nullptr);

// ws_window_fft_warp.cu
bool window_fft_warp_supported(const Params& p);
cudaError_t launch_window_fft_warp(Params p, cudaStream_t stream);

// ws_phase.cu: A6 phase chain of every window of a spectra plane (rows s * spec_nwin + (w - spec_w0))
bool phase_from_spectra_supported(int N);
cudaError_t launch_phase_from_spectra(const double* spectra, int64_t spec_nwin, int64_t spec_w0, int n_series,
                                      int64_t win_offset, int64_t chunk_nwin, int64_t nwin, int N,
                                      double* phase, cudaStream_t stream);

// ws_sliding.cu
bool sliding_shared_supported(const Params& p);
// *which (when non-null) receives "sliding_overlap" if the producer / consumer form ran
cudaError_t launch_sliding_shared(Params p, cudaStream_t stream, const char** which = nullptr);

// ws_rows.cu
bool rows_from_band_supported(const Params& p);
cudaError_t launch_rows_from_band(const Params& p, cudaStream_t stream);

// ws_series.cu
cudaError_t launch_kalman4d(const double* z_base, int64_t series_stride, int64_t z_step,
                            int32_t n_series, int64_t nwin, const KalmanParams& kp, double* out,
                            cudaStream_t stream);
cudaError_t launch_wkalman(const double* contrib, const int32_t* bins, const double* meas_base,
                           int64_t series_stride, int64_t meas_step, int32_t n_series, int64_t nwin,
                           int32_t K, double q, double r, double p0, double* out, cudaStream_t stream);

// ws_pla.cu
cudaError_t launch_pla(const double* series, int64_t series_stride, int32_t n_series, int64_t nwin,
                       int32_t N, int32_t hop, int32_t max_segments, double max_error, double* lines,
                       int32_t* seg_bounds, int32_t* seg_counts, int32_t bounds_cap, int32_t* overflow,
                       cudaStream_t stream);
cudaError_t launch_gather_last(const double* feed, int32_t n_series, int64_t cn, int32_t N, double* z,
                               int64_t z_stride, int64_t wa, cudaStream_t stream);

// ws_zigzag.cu
cudaError_t launch_applied_price(const double* o, const double* h, const double* l, const double* c, int64_t n,
                                 int mode, double* out, cudaStream_t stream);
cudaError_t launch_zigzag(const double* zmain, const double* zhigh, const double* zlow, const double* fallback,
                          int32_t n_series, int32_t len, int32_t N, int32_t hop, int rule, int mode, int min_pivots,
                          double* pv, int32_t* prev, int32_t* next, double* lines, int32_t* valid,
                          cudaStream_t stream);

// ws_cache.cu
// bars [bar0, bar0 + nbars) of the per-bar cache record; `rows` holds every window of the series
cudaError_t launch_cycle_cache(const double* rows, int64_t n_windows, int32_t top_k, int32_t stride, int32_t N,
                               int32_t hop, int64_t bars, int64_t bar0, int64_t nbars, double period_seconds,
                               const ::wavespec_cache_params& cp, double* out, cudaStream_t stream);

// ws_inverse.cu
// d_bins (optional, [n_windows][K] int32, -1 = absent): keep only these bins of each window's spectrum
cudaError_t launch_inverse_real(const double* d_spec, int32_t N, int64_t n_windows, const double2* tw,
                                double* d_out, cudaStream_t stream, const int32_t* d_bins = nullptr,
                                int32_t K = 0);

}  // namespace ws
