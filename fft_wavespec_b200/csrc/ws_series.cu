// ws_series.cu — per-series sequential recursions, one series per thread across the batch.
// Compiled with -fmad=false: every add/mul/div/sqrt is a separately rounded IEEE operation in
// the reference's evaluation order, so given identical inputs the results are bit-identical to
// the CPU statement.
//
//  * Kalman4D  : ResetKalmanState / StepKalman4D,
//                Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:2015-2125 (defaults :885-901),
//                driven as in the bar loop :3354-3360 (reset on the first processed bar, then
//                step that same bar).  State = 4 + 16 doubles carried across bars.
//  * weight-Kalman blend : UpdateKalman, Legacy/WaveSpecZZ_1.0.4-kalman.mq5:194-231
//                (= UpdateKalmanWave, Legacy/WaveSpecZZ_1.0.4-old.mq5:2606-2649).
#include "ws_common.cuh"
#include "ws_series.h"

namespace ws {

__global__ void kalman4d_kernel(const double* __restrict__ z_base, int64_t series_stride,
                                int64_t z_step, int32_t n_series, int64_t nwin,
                                const KalmanParams kp, double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_series) return;
    const double* z = z_base + (int64_t)s * series_stride;
    double* o = out + (int64_t)s * nwin;

    const double q_scale = fmax(0.05, kp.follow_strength);
    const double Qp = fmax(1e-9, kp.q_pos * q_scale);
    const double Qv = fmax(1e-9, kp.q_vel * q_scale);
    const double Qa = fmax(1e-9, kp.q_acc * q_scale);
    const double Qj = fmax(1e-9, kp.q_jerk * q_scale);
    const double R = fmax(1e-9, kp.meas_noise);
    const double c16 = 1.0 / 6.0, c112 = 1.0 / 12.0, c136 = 1.0 / 36.0;

    double pos = 0, vel = 0, acc = 0, jerk = 0;
    double P00 = 0, P01 = 0, P02 = 0, P03 = 0, P10 = 0, P11 = 0, P12 = 0, P13 = 0;
    double P20 = 0, P21 = 0, P22 = 0, P23 = 0, P30 = 0, P31 = 0, P32 = 0, P33 = 0;
    double ema_prev = 0.0;
    bool ema_ready = false;

    // measurements are prefetched kPre steps ahead: the recursion is a dependent FP64 chain and a
    // global load issued at the top of each step would put its full latency on that chain
    constexpr int kPre = 8;
    double zq[kPre];
#pragma unroll
    for (int i = 0; i < kPre; i++) zq[i] = (i < nwin) ? z[(int64_t)i * z_step] : 0.0;
    for (int64_t w = 0; w < nwin; w++) {
        const double zz = zq[0];
#pragma unroll
        for (int i = 0; i < kPre - 1; i++) zq[i] = zq[i + 1];
        zq[kPre - 1] = (w + kPre < nwin) ? z[(w + kPre) * z_step] : 0.0;
        if (w == 0) {   // ResetKalmanState(first measurement)
            pos = zz; vel = kp.init_vel; acc = kp.init_acc; jerk = kp.init_jerk;
            P00 = fmax(1e-9, kp.init_var_pos); P11 = fmax(1e-9, kp.init_var_vel);
            P22 = fmax(1e-9, kp.init_var_acc); P33 = fmax(1e-9, kp.init_var_jerk);
            P01 = P02 = P03 = P10 = P12 = P13 = P20 = P21 = P23 = P30 = P31 = P32 = 0.0;
            ema_ready = false;
        }
        double x0p = pos + vel + 0.5 * acc + c16 * jerk;
        double x1p = vel + acc + 0.5 * jerk;
        double x2p = acc + jerk;
        double x3p = jerk;

        double P00p = P00 + P01 + 0.5 * P02 + c16 * P03
                    + P10 + P11 + 0.5 * P12 + c16 * P13
                    + 0.5 * P20 + 0.5 * P21 + 0.25 * P22 + c112 * P23
                    + c16 * P30 + c16 * P31 + c112 * P32 + c136 * P33
                    + Qp;
        double P01p = P01 + P02 + 0.5 * P03 + P11 + P12 + 0.5 * P13 + 0.5 * P21 + 0.5 * P22 + 0.25 * P23 + c16 * P31 + c16 * P32 + c112 * P33;
        double P02p = P02 + P03 + P12 + P13 + 0.5 * P22 + 0.5 * P23 + c16 * P32 + c16 * P33;
        double P03p = P03 + P13 + 0.5 * P23 + c16 * P33;
        double P11p = P11 + 2.0 * P12 + P13 + P21 + 2.0 * P22 + P23 + 0.5 * P31 + 0.5 * P32 + 0.25 * P33 + Qv;
        double P12p = P12 + P13 + P22 + P23 + 0.5 * P32 + 0.5 * P33;
        double P13p = P13 + P23 + 0.5 * P33;
        double P22p = P22 + 2.0 * P23 + P33 + Qa;
        double P23p = P23 + P33;
        double P33p = P33 + Qj;
        double P10p = P01p, P20p = P02p, P30p = P03p;
        double P21p = P12p, P31p = P13p, P32p = P23p;

        double y = zz - x0p;
        double S = P00p + R;
        if (kp.adapt_gain > 0.0) {
            double sigma = sqrt(S);
            double k = fmin(5.0, fabs(y) / sigma) * kp.adapt_gain;
            double boost = 1.0 + k;
            P00p += (boost - 1.0) * Qp;
            P11p += (boost - 1.0) * Qv;
            P22p += (boost - 1.0) * Qa;
            P33p += (boost - 1.0) * Qj;
            S = P00p + R;
        }
        if (kp.clip_std > 0.0) {
            double sigma = sqrt(S);
            double lim = kp.clip_std * sigma;
            if (y > lim) y = lim;
            if (y < -lim) y = -lim;
        }
        double K0 = P00p / S, K1 = P10p / S, K2 = P20p / S, K3 = P30p / S;

        pos = x0p + K0 * y;
        vel = x1p + K1 * y;
        acc = x2p + K2 * y;
        jerk = x3p + K3 * y;

        double P00n = (1.0 - K0) * P00p, P01n = (1.0 - K0) * P01p, P02n = (1.0 - K0) * P02p, P03n = (1.0 - K0) * P03p;
        double P10n = P10p - K1 * P00p, P11n = P11p - K1 * P01p, P12n = P12p - K1 * P02p, P13n = P13p - K1 * P03p;
        double P20n = P20p - K2 * P00p, P21n = P21p - K2 * P01p, P22n = P22p - K2 * P02p, P23n = P23p - K2 * P03p;
        double P30n = P30p - K3 * P00p, P31n = P31p - K3 * P01p, P32n = P32p - K3 * P02p, P33n = P33p - K3 * P03p;

        P00 = fmax(1e-12, P00n); P01 = P01n; P02 = P02n; P03 = P03n;
        P10 = P10n; P11 = fmax(1e-12, P11n); P12 = P12n; P13 = P13n;
        P20 = P20n; P21 = P21n; P22 = fmax(1e-12, P22n); P23 = P23n;
        P30 = P30n; P31 = P31n; P32 = P32n; P33 = fmax(1e-12, P33n);

        double outv = pos;
        if (kp.ema_blend_period > 0.0) {
            double alpha = 2.0 / (kp.ema_blend_period + 1.0);
            if (!ema_ready) { ema_prev = outv; ema_ready = true; }
            ema_prev = alpha * outv + (1.0 - alpha) * ema_prev;
            outv = ema_prev;
        }
        o[w] = outv;
    }
}

__global__ void wkalman_kernel(const double* __restrict__ contrib, const int32_t* __restrict__ bins,
                               const double* __restrict__ meas_base, int64_t series_stride,
                               int64_t meas_step, int32_t n_series, int64_t nwin, int32_t K,
                               double q, double r, double p0, double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_series) return;
    const double Q = fmax(1e-9, q), R = fmax(1e-9, r);
    double wgt[kMaxTopK], cov[kMaxTopK];
    for (int i = 0; i < kMaxTopK; i++) { wgt[i] = 0.0; cov[i] = fmax(1e-6, p0); }
    const double* meas = meas_base + (int64_t)s * series_stride;
    // inputs of the next bar are fetched while the current bar's dependent chain runs
    // (fast path K <= 8 keeps them in registers)
    double nv[8]; int nb[8]; double nm = 0.0;
    const bool small = K <= 8;
    if (small && nwin > 0) {
        const double* cv0 = contrib + (int64_t)s * nwin * K; const int32_t* bn0 = bins + (int64_t)s * nwin * K;
#pragma unroll
        for (int k = 0; k < 8; k++) { nv[k] = k < K ? cv0[k] : 0.0; nb[k] = k < K ? bn0[k] : -1; }
        nm = meas[0];
    }
    for (int64_t w = 0; w < nwin; w++) {
        const double* cv = contrib + ((int64_t)s * nwin + w) * K;
        const int32_t* bn = bins + ((int64_t)s * nwin + w) * K;
        double vals[kMaxTopK];
        int use = 0;
        double mz;
        if (small) {
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < K && nb[k] >= 0) vals[use++] = nv[k];
            mz = nm;
            if (w + 1 < nwin) {
#pragma unroll
                for (int k = 0; k < 8; k++) { nv[k] = k < K ? cv[K + k] : 0.0; nb[k] = k < K ? bn[K + k] : -1; }
                nm = meas[(w + 1) * meas_step];
            }
        } else {
            for (int k = 0; k < K; k++) if (bn[k] >= 0) vals[use++] = cv[k];
            mz = meas[w * meas_step];
        }
        double blended = 0.0;
        if (use > 0) {
            double residual = mz;
            double innovation = R;
            double cov_tmp[kMaxTopK], w_tmp[kMaxTopK];
            for (int i = 0; i < use; i++) {
                cov[i] += Q;
                cov_tmp[i] = cov[i];
                w_tmp[i] = wgt[i];
                residual -= vals[i] * w_tmp[i];
                innovation += vals[i] * vals[i] * cov_tmp[i];
            }
            if (innovation < 1e-9) innovation = R;
            for (int i = 0; i < use; i++) {
                const double H = vals[i];
                const double c = cov_tmp[i];
                const double Kg = (c * H) / innovation;
                const double nw = w_tmp[i] + Kg * residual;
                const double nc = (1.0 - Kg * H) * c;
                wgt[i] = nw;
                cov[i] = fmax(nc, 1e-9);
                blended += nw * H;
            }
        }
        out[(int64_t)s * nwin + w] = blended;
    }
}

// ---- A13: persistent period tracker pool + 12 stable slots -------------------------------------
// Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:1415-1667, driven per bar as in :3450-3504.
// One series per WARP, strictly sequential over bars AND over the band candidates of a bar (each
// UpdateTracker rewrites the period later candidates are matched against).  What is parallel is the
// inside of a step: the lanes scan the tracker pool for the closest active tracker (the reference's
// first-strictly-smallest rule = argmin over (difference, index)), and the refill of free slots
// (first strictly largest power = argmax over (power desc, index asc)).  The pool lives in shared
// memory while the kernel runs and in global memory (one TrackerState per series) between window
// chunks.  Round 1 walked it with one thread per series out of global memory: ~10 ms for the 38 bars
// the structure needs to settle, whatever the series count.
// The reference quirks are kept: inactive trackers are never re-matched (:1437), erasing shifts
// the array while the slot table keeps raw indices (:1514-1519, :1584-1589).
__global__ void __launch_bounds__(32)
tracker_kernel(const double2* __restrict__ band, int32_t band_lo, int32_t nband,
               int32_t n_series, int64_t chunk_nwin, int64_t n_process, int64_t win_offset,
               int64_t nwin, int32_t N, double tol, int32_t max_inactive,
               TrackerState* __restrict__ states, int32_t* __restrict__ trk_index,
               double* __restrict__ trk_period) {
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ double sh_period[kTrackerCap], sh_power[kTrackerCap];
    __shared__ int32_t sh_index[kTrackerCap], sh_active[kTrackerCap], sh_inactive[kTrackerCap];
    __shared__ int32_t sh_slot[12];
    const int s = blockIdx.x;
    const int lane = threadIdx.x;
    if (s >= n_series) return;
    TrackerState& st = states[s];
    int count = 0;
    if (win_offset == 0) {
        if (lane < 12) sh_slot[lane] = -1;
    } else {
        count = st.count;
        for (int i = lane; i < count; i += 32) {
            sh_period[i] = st.period[i]; sh_power[i] = st.power[i]; sh_index[i] = st.fft_index[i];
            sh_active[i] = st.is_active[i]; sh_inactive[i] = st.bars_inactive[i];
        }
        if (lane < 12) sh_slot[lane] = st.slot[lane];
    }
    __syncwarp();
    for (int64_t wl = 0; wl < n_process; wl++) {
        const double2* bw = band + ((int64_t)s * chunk_nwin + wl) * nband;
        for (int c = 0; c < nband; c++) {
            const int j = band_lo + c;
            const double period = (j > 0) ? (double)N / j : 0;
            if (period <= 0) continue;
            const double2 x = bw[c];
            const double power = (x.x * x.x) + (x.y * x.y);
            // FindClosestTracker (:1433-1452): lane-strided scan in ascending index, then the warp's argmin
            int best = -1;
            double smallest = 999999;
            for (int i = lane; i < count; i += 32) {
                if (sh_inactive[i] > 0) continue;
                const double tp = sh_period[i];
                const double diff = fabs(tp - period);
                bool same = false;
                if (period > 0 && tp > 0) {
                    const double d2 = fabs(period - tp);
                    const double avg = (period + tp) / 2.0;
                    const double pct = (d2 / avg) * 100.0;
                    same = pct <= tol;
                }
                if (same && diff < smallest) { smallest = diff; best = i; }
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const double os = __shfl_xor_sync(kFull, smallest, m);
                const int ob = __shfl_xor_sync(kFull, best, m);
                if (ob >= 0 && (best < 0 || os < smallest || (os == smallest && ob < best))) { smallest = os; best = ob; }
            }
            if (best >= 0) {
                if (lane == 0) {
                    sh_period[best] = period; sh_index[best] = j; sh_power[best] = power;
                    sh_active[best] = 1; sh_inactive[best] = 0;
                }
            } else if (count < kTrackerCap) {
                if (lane == 0) {
                    sh_period[count] = period; sh_index[count] = j; sh_power[count] = power;
                    sh_active[count] = 1; sh_inactive[count] = 0;
                }
                count++;
            }
            __syncwarp();
        }
        // DeactivateUnseenTrackers (:1497-1529): the descending erase-shift loop visits every tracker
        // once, so it is a stable compaction of the survivors
        if (lane == 0) {
            int w = 0;
            for (int i = 0; i < count; i++) {
                bool keep = true;
                if (!sh_active[i]) { sh_inactive[i]++; keep = sh_inactive[i] < max_inactive; }
                if (keep) {
                    if (w != i) {
                        sh_period[w] = sh_period[i]; sh_power[w] = sh_power[i]; sh_index[w] = sh_index[i];
                        sh_active[w] = sh_active[i]; sh_inactive[w] = sh_inactive[i];
                    }
                    w++;
                }
            }
            count = w;
        }
        count = __shfl_sync(kFull, count, 0);
        __syncwarp();
        for (int i = lane; i < count; i += 32) sh_active[i] = 0;
        // UpdateStableSlots (:1570-1667)
        if (lane < 12) { const int t = sh_slot[lane]; if (t < 0 || t >= count) sh_slot[lane] = -1; }
        __syncwarp();
        const int64_t o = ((int64_t)s * nwin + win_offset + wl) * 12;
        for (int q = 0; q < 12; q++) {
            int t = sh_slot[q];                       // uniform
            if (t < 0) {
                // the strongest unused tracker in the order of the reference's stable descending
                // bubble sort: larger power first, equal powers keep ascending index
                int chosen = -1; double bp = 0.0;
                for (int i = lane; i < count; i += 32) {
                    bool used = false;
#pragma unroll
                    for (int r = 0; r < 12; r++) used |= (sh_slot[r] == i);
                    if (used) continue;
                    const double pw = sh_power[i];
                    if (chosen < 0 || pw > bp) { chosen = i; bp = pw; }
                }
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) {
                    const double op = __shfl_xor_sync(kFull, bp, m);
                    const int oc = __shfl_xor_sync(kFull, chosen, m);
                    if (oc >= 0 && (chosen < 0 || op > bp || (op == bp && oc < chosen))) { bp = op; chosen = oc; }
                }
                t = chosen;
                __syncwarp();
                if (lane == 0) sh_slot[q] = chosen;
                __syncwarp();
            }
            if (lane == 0) {
                if (t >= 0) { trk_period[o + q] = sh_period[t]; trk_index[o + q] = sh_index[t]; }
                else { trk_period[o + q] = 0.0; trk_index[o + q] = 0; }
            }
        }
        __syncwarp();
    }
    for (int i = lane; i < count; i += 32) {
        st.period[i] = sh_period[i]; st.power[i] = sh_power[i]; st.fft_index[i] = sh_index[i];
        st.is_active[i] = sh_active[i]; st.bars_inactive[i] = sh_inactive[i];
    }
    if (lane < 12) st.slot[lane] = sh_slot[lane];
    if (lane == 0) st.count = count;
}

cudaError_t launch_tracker(const double2* band, int32_t band_lo, int32_t nband, int32_t n_series,
                           int64_t chunk_nwin, int64_t n_process, int64_t win_offset, int64_t nwin, int32_t N,
                           double tol, int32_t max_inactive, TrackerState* states, int32_t* trk_index,
                           double* trk_period, cudaStream_t stream) {
    tracker_kernel<<<n_series, 32, 0, stream>>>(
        band, band_lo, nband, n_series, chunk_nwin, n_process, win_offset, nwin, N, tol, max_inactive, states,
        trk_index, trk_period);
    return cudaGetLastError();
}

// Once the tracker STRUCTURE has reached its fixed point (see tracker_fixed_point in ws_abi.cu) the
// twelve slots of every later bar repeat those of bar `last`: one coalesced broadcast.
__global__ void tracker_fill_kernel(int32_t n_series, int64_t nwin, int64_t last, int32_t* __restrict__ trk_index,
                                    double* __restrict__ trk_period) {
    const int s = blockIdx.y;
    const int64_t base = (int64_t)s * nwin * 12;
    const int64_t total = (nwin - 1 - last) * 12;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % 12);
        trk_index[base + (last + 1) * 12 + i] = trk_index[base + last * 12 + q];
        trk_period[base + (last + 1) * 12 + i] = trk_period[base + last * 12 + q];
    }
}

cudaError_t launch_tracker_fill(int32_t n_series, int64_t nwin, int64_t last, int32_t* trk_index,
                                double* trk_period, cudaStream_t stream) {
    if (last >= nwin - 1) return cudaSuccess;
    int64_t blocks = ((nwin - 1 - last) * 12 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    dim3 grid((unsigned)blocks, (unsigned)n_series);
    tracker_fill_kernel<<<grid, 256, 0, stream>>>(n_series, nwin, last, trk_index, trk_period);
    return cudaGetLastError();
}

cudaError_t launch_kalman4d(const double* z_base, int64_t series_stride, int64_t z_step,
                            int32_t n_series, int64_t nwin, const KalmanParams& kp, double* out,
                            cudaStream_t stream) {
    const int threads = 32;
    const int blocks = (n_series + threads - 1) / threads;
    kalman4d_kernel<<<blocks, threads, 0, stream>>>(z_base, series_stride, z_step, n_series, nwin, kp, out);
    return cudaGetLastError();
}

cudaError_t launch_wkalman(const double* contrib, const int32_t* bins, const double* meas_base,
                           int64_t series_stride, int64_t meas_step, int32_t n_series, int64_t nwin,
                           int32_t K, double q, double r, double p0, double* out, cudaStream_t stream) {
    const int threads = 32;
    const int blocks = (n_series + threads - 1) / threads;
    wkalman_kernel<<<blocks, threads, 0, stream>>>(contrib, bins, meas_base, series_stride, meas_step,
                                                   n_series, nwin, K, q, r, p0, out);
    return cudaGetLastError();
}

}  // namespace ws
