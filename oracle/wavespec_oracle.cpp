// wavespec_oracle.cpp — CPU restatement of the reference's MQL5 arithmetic (see the header for
// scope and the "parity unpinned" note).  TEST INFRASTRUCTURE ONLY.
//
// Transcription rules: IEEE double, expression order as the MQL5 source
// evaluates it (left to right, ints promoted to double), no FMA contraction (the Makefile passes
// -ffp-contract=off), MathRound = round half away from zero, (int) truncates.
#include "wavespec_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {
const double kPi = 3.14159265358979323846;  // MQL5 M_PI
}

extern "C" {

int oracle_version(void) { return 10000; }

// ---------------------------------------------------------------------------------------------
// A4  L/WaveSpecZZ_1.0.2.mq5:938-975
void oracle_fft_forward(const double* data, int n, double* re, double* im) {
    if (n <= 1) {
        if (n == 1) { re[0] = data[0]; im[0] = 0.0; }
        return;
    }
    std::vector<double> tmp(data, data + n);
    // bit-reversal permutation (:943-949)
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; (j & bit) != 0; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t = tmp[i]; tmp[i] = tmp[j]; tmp[j] = t; }
    }
    for (int i = 0; i < n; i++) { re[i] = tmp[i]; im[i] = 0.0; }
    // butterflies with the multiplicative twiddle recurrence (:952-974)
    for (int len = 2; len <= n; len <<= 1) {
        double ang = -2 * kPi / len;
        double wlen_r = std::cos(ang), wlen_i = std::sin(ang);
        for (int i = 0; i < n; i += len) {
            double w_r = 1.0, w_i = 0.0;
            for (int j = 0; j < len / 2; j++) {
                int a = i + j, b = i + j + len / 2;
                double t_r = re[b] * w_r - im[b] * w_i;
                double t_i = re[b] * w_i + im[b] * w_r;
                re[b] = re[a] - t_r;
                im[b] = im[a] - t_i;
                re[a] += t_r;
                im[a] += t_i;
                double w_t = w_r;
                w_r = w_r * wlen_r - w_i * wlen_i;
                w_i = w_t * wlen_i + w_i * wlen_r;
            }
        }
    }
}

// A8d  inverse of the DLL contract of gpu_fft_real_forward (declared L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:27,
// used L/WaveSpecZZ_1.0.4-core.mq5:426).  The reference holds no CPU inverse; this is the textbook
// identity IFFT(X) = conj(FFT(conj X)) / n run through the SAME butterfly loop as
// FourierTransformManual (L/WaveSpecZZ_1.0.2.mq5:943-974: bit reversal, twiddle recurrence), applied to
// the Hermitian extension of the n/2 interleaved bins (Nyquist bin, which the forward contract
// drops, taken as 0).  The 1/n normalisation is this build's decision (include/wavespec_abi.h).
void oracle_fft_inverse(const double* spec, int n, double* out) {
    if (n < 2) { if (n == 1) out[0] = spec[0]; return; }
    std::vector<double> re(n, 0.0), im(n, 0.0);
    const int h = n / 2;
    for (int k = 0; k < h; k++) {
        const double xr = spec[2 * k], xi = (k == 0) ? 0.0 : spec[2 * k + 1];   // X[0] of a real signal is real
        re[k] = xr; im[k] = -xi;                                                  // conj X[k]
        if (k > 0) { re[n - k] = xr; im[n - k] = xi; }                            // conj X[n-k] = conj conj X[k]
    }
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; (j & bit) != 0; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    for (int len = 2; len <= n; len <<= 1) {
        double ang = -2 * kPi / len;
        double wlen_r = std::cos(ang), wlen_i = std::sin(ang);
        for (int i = 0; i < n; i += len) {
            double w_r = 1.0, w_i = 0.0;
            for (int j = 0; j < len / 2; j++) {
                int a = i + j, b = i + j + len / 2;
                double t_r = re[b] * w_r - im[b] * w_i;
                double t_i = re[b] * w_i + im[b] * w_r;
                re[b] = re[a] - t_r;
                im[b] = im[a] - t_i;
                re[a] += t_r;
                im[a] += t_i;
                double w_t = w_r;
                w_r = w_r * wlen_r - w_i * wlen_i;
                w_i = w_t * wlen_i + w_i * wlen_r;
            }
        }
    }
    for (int i = 0; i < n; i++) out[i] = re[i] / (double)n;
}

// A4' L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3422-3433 (inverse view: what the DLL must return)
void oracle_fft_interleaved(const double* data, int n, double* out) {
    std::vector<double> re(n), im(n);
    oracle_fft_forward(data, n, re.data(), im.data());
    for (int k = 0; k < n / 2; k++) { out[2 * k] = re[k]; out[2 * k + 1] = im[k]; }
}

// ---------------------------------------------------------------------------------------------
// A3  L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:1126-1177 (+ L/WaveSpecZZ_gpu_wip.mq5:954 as type 5)
void oracle_apply_window(double* data, int n, int type) {
    switch (type) {
        case 1:
            for (int i = 0; i < n; i++) {
                double w = 0.5 * (1.0 - std::cos(2.0 * kPi * i / (n - 1)));
                data[i] *= w;
            }
            break;
        case 2:
            for (int i = 0; i < n; i++) {
                double w = 0.54 - 0.46 * std::cos(2.0 * kPi * i / (n - 1));
                data[i] *= w;
            }
            break;
        case 3:
            for (int i = 0; i < n; i++) {
                double w = 0.42 - 0.5 * std::cos(2.0 * kPi * i / (n - 1)) +
                           0.08 * std::cos(4.0 * kPi * i / (n - 1));
                data[i] *= w;
            }
            break;
        case 4:
            for (int i = 0; i < n; i++) {
                double w = 1.0 - std::fabs((2.0 * i - n + 1) / (n - 1));
                data[i] *= w;
            }
            break;
        case 5: {
            const double denom = (double)(n - 1);
            for (int i = 0; i < n; i++) {
                const double w = 0.5 - 0.5 * std::cos((2.0 * kPi * i) / denom);
                data[i] = data[i] * w;
            }
            break;
        }
        default:
            break;
    }
}

// A2a L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3367-3379
void oracle_detrend_iir(const double* price, int n, double trend_period, double* trend,
                        double* detrended) {
    double omega = 2.0 * kPi / trend_period;
    double alpha = (1.0 - std::sin(omega)) / std::cos(omega);
    double c = (1.0 - alpha) / 2.0;
    trend[0] = c * (price[0] + price[0]);
    if (n > 1) trend[1] = c * (price[1] + price[0]) + alpha * trend[0];
    for (int j = 2; j < n; j++) trend[j] = c * (price[j] + price[j - 1]) + alpha * trend[j - 1];
    for (int j = 0; j < n; j++) detrended[j] = price[j] - trend[j];
}

// A2b L/WaveSpecZZ_gpu_wip.mq5:935-957
void oracle_mean_hann(const double* price, int n, double* out) {
    if (n <= 0) return;
    double mean = 0.0;
    for (int i = 0; i < n; ++i) mean += price[i];
    mean /= (double)n;
    if (n == 1) { out[0] = price[0] - mean; return; }
    const double denom = (double)(n - 1);
    for (int i = 0; i < n; ++i) {
        const double w = 0.5 - 0.5 * std::cos((2.0 * kPi * i) / denom);
        out[i] = (price[i] - mean) * w;
    }
}

// A5  L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3441-3445
void oracle_power(const double* re, const double* im, int n, double* spectrum) {
    for (int j = 0; j < n / 2; j++) spectrum[j] = (re[j] * re[j]) + (im[j] * im[j]);
}

// A6  L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:1183-1263
void oracle_phase_chain(const double* re, const double* im, int count, double* phase,
                        double* unwrapped, double* gd) {
    for (int i = 0; i < count; i++) phase[i] = std::atan2(im[i], re[i]);
    if (count == 0) return;
    unwrapped[0] = phase[0];
    for (int i = 1; i < count; i++) {
        double diff = phase[i] - phase[i - 1];
        double corr = 0;
        if (diff > kPi) corr = -2.0 * kPi;
        else if (diff < -kPi) corr = 2.0 * kPi;
        unwrapped[i] = unwrapped[i - 1] + diff + corr;
    }
    if (count < 3) { for (int i = 0; i < count; i++) gd[i] = 0.0; return; }
    gd[0] = -(unwrapped[1] - unwrapped[0]);
    for (int i = 1; i < count - 1; i++) gd[i] = -(unwrapped[i + 1] - unwrapped[i - 1]) / 2.0;
    gd[count - 1] = -(unwrapped[count - 1] - unwrapped[count - 2]);
    for (int i = 0; i < count; i++) {
        if (gd[i] > 100.0) gd[i] = 100.0;
        if (gd[i] < -100.0) gd[i] = -100.0;
    }
}

// ---------------------------------------------------------------------------------------------
// A7a L/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:537-554 (K generalised from 8)
void oracle_topk_insertion(const double* spectrum, int n, double min_period, double max_period,
                           int top_k, int* top_bin, double* top_pow) {
    const int bins = n / 2;
    for (int s = 0; s < top_k; s++) { top_pow[s] = -1.0; top_bin[s] = -1; }
    int min_index = (int)std::ceil((double)n / max_period);
    int max_index = (int)std::floor((double)n / min_period);
    if (max_index >= bins) max_index = bins - 1;
    if (min_index < 0) min_index = 0;   // guard only; the reference never reaches it
    for (int b = min_index; b <= max_index; b++) {
        double p = spectrum[b];
        for (int s = 0; s < top_k; s++) {
            if (p > top_pow[s]) {
                for (int t = top_k - 1; t > s; t--) { top_pow[t] = top_pow[t - 1]; top_bin[t] = top_bin[t - 1]; }
                top_pow[s] = p; top_bin[s] = b;
                break;
            }
        }
    }
}

// A7b L/WaveSpecZZ_1.0.4-kalman.mq5:143-180
int oracle_collect_sorted(const double* re, const double* im, int n, double min_period,
                          double max_period, int* idx, double* pow) {
    const int bins = n / 2;
    const int min_idx = (int)std::ceil((double)n / max_period);
    const int max_idx = (int)std::floor((double)n / min_period);
    int count = 0;
    for (int k = (min_idx > 1 ? min_idx : 1); k <= max_idx && k < bins; ++k) {
        const double r = re[k], i = im[k];
        idx[count] = k;
        pow[count] = r * r + i * i;
        count++;
    }
    for (int i = 0; i < count - 1; ++i) {
        int m = i;
        double mp = pow[i];
        for (int j = i + 1; j < count; ++j)
            if (pow[j] > mp) { m = j; mp = pow[j]; }
        if (m != i) {
            int ti = idx[i]; double tp = pow[i];
            idx[i] = idx[m]; pow[i] = pow[m];
            idx[m] = ti; pow[m] = tp;
        }
    }
    return count;
}

// A8a L/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:559-568
void oracle_recon_last(const double* re, const double* im, int n, int bin, double power,
                       double* wave, double* period) {
    double amp = 0.0, per = 0.0;
    if (bin > 0) {
        per = (double)n / (double)bin;
        double mag = std::sqrt(power);
        double ph = std::atan2(im[bin], re[bin]);
        double nn = (double)(n - 1);
        amp = (mag / (double)n) * std::cos(ph + 2.0 * kPi * (double)bin * nn / (double)n);
    }
    *wave = amp; *period = per;
}

// A8b L/WaveSpecZZ_1.0.4-kalman.mq5:182-192
double oracle_contribution(const double* re, const double* im, int n, int k) {
    const double r = re[k], i = im[k];
    const double n0 = n - 1;
    const double angle = 2.0 * kPi * k * n0 / n;
    const double ca = std::cos(angle), sa = std::sin(angle);
    const double scale = 2.0 / n;
    return scale * (r * ca - i * sa);
}

// ---------------------------------------------------------------------------------------------
// A9  L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:885-901 (defaults), :2015-2125
void oracle_kalman4d_defaults(oracle_kalman4d_params* p) {
    p->follow_strength = 1.0; p->q_pos = 0.01; p->q_vel = 0.003; p->q_acc = 0.0008;
    p->q_jerk = 0.0002; p->adapt_gain = 0.8; p->meas_noise = 1.0;
    p->init_var_pos = 16.0; p->init_var_vel = 9.0; p->init_var_acc = 4.0; p->init_var_jerk = 1.0;
    p->init_vel = 0.0; p->init_acc = 0.0; p->init_jerk = 0.0; p->clip_std = 6.0;
    p->ema_blend_period = 0.0;
}

void oracle_kalman4d_reset(oracle_kalman4d_state* s, const oracle_kalman4d_params* p, double first) {
    s->pos = first; s->vel = p->init_vel; s->acc = p->init_acc; s->jerk = p->init_jerk;
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) s->P[a][b] = 0.0;
    s->P[0][0] = std::fmax(1e-9, p->init_var_pos);
    s->P[1][1] = std::fmax(1e-9, p->init_var_vel);
    s->P[2][2] = std::fmax(1e-9, p->init_var_acc);
    s->P[3][3] = std::fmax(1e-9, p->init_var_jerk);
    s->ready = 1; s->ema_ready = 0; s->ema_prev = 0.0;
}

double oracle_kalman4d_step(oracle_kalman4d_state* s, const oracle_kalman4d_params* p, double z) {
    const double q_scale = std::fmax(0.05, p->follow_strength);
    double Qp = std::fmax(1e-9, p->q_pos * q_scale);
    double Qv = std::fmax(1e-9, p->q_vel * q_scale);
    double Qa = std::fmax(1e-9, p->q_acc * q_scale);
    double Qj = std::fmax(1e-9, p->q_jerk * q_scale);
    double R = std::fmax(1e-9, p->meas_noise);
    const double P00 = s->P[0][0], P01 = s->P[0][1], P02 = s->P[0][2], P03 = s->P[0][3];
    const double P10 = s->P[1][0], P11 = s->P[1][1], P12 = s->P[1][2], P13 = s->P[1][3];
    const double P20 = s->P[2][0], P21 = s->P[2][1], P22 = s->P[2][2], P23 = s->P[2][3];
    const double P30 = s->P[3][0], P31 = s->P[3][1], P32 = s->P[3][2], P33 = s->P[3][3];

    double x0p = s->pos + s->vel + 0.5 * s->acc + (1.0 / 6.0) * s->jerk;
    double x1p = s->vel + s->acc + 0.5 * s->jerk;
    double x2p = s->acc + s->jerk;
    double x3p = s->jerk;

    double P00p = P00 + P01 + 0.5 * P02 + (1.0 / 6.0) * P03
                + P10 + P11 + 0.5 * P12 + (1.0 / 6.0) * P13
                + 0.5 * P20 + 0.5 * P21 + 0.25 * P22 + (1.0 / 12.0) * P23
                + (1.0 / 6.0) * P30 + (1.0 / 6.0) * P31 + (1.0 / 12.0) * P32 + (1.0 / 36.0) * P33
                + Qp;
    double P01p = P01 + P02 + 0.5 * P03 + P11 + P12 + 0.5 * P13 + 0.5 * P21 + 0.5 * P22 + 0.25 * P23 + (1.0 / 6.0) * P31 + (1.0 / 6.0) * P32 + (1.0 / 12.0) * P33;
    double P02p = P02 + P03 + P12 + P13 + 0.5 * P22 + 0.5 * P23 + (1.0 / 6.0) * P32 + (1.0 / 6.0) * P33;
    double P03p = P03 + P13 + 0.5 * P23 + (1.0 / 6.0) * P33;
    double P11p = P11 + 2.0 * P12 + P13 + P21 + 2.0 * P22 + P23 + 0.5 * P31 + 0.5 * P32 + 0.25 * P33 + Qv;
    double P12p = P12 + P13 + P22 + P23 + 0.5 * P32 + 0.5 * P33;
    double P13p = P13 + P23 + 0.5 * P33;
    double P22p = P22 + 2.0 * P23 + P33 + Qa;
    double P23p = P23 + P33;
    double P33p = P33 + Qj;
    double P10p = P01p, P20p = P02p, P30p = P03p;
    double P21p = P12p, P31p = P13p, P32p = P23p;

    double y = z - x0p;
    double S = P00p + R;
    if (p->adapt_gain > 0.0) {
        double sigma = std::sqrt(S);
        double k = std::fmin(5.0, std::fabs(y) / sigma) * p->adapt_gain;
        double boost = 1.0 + k;
        P00p += (boost - 1.0) * Qp;
        P11p += (boost - 1.0) * Qv;
        P22p += (boost - 1.0) * Qa;
        P33p += (boost - 1.0) * Qj;
        S = P00p + R;
    }
    if (p->clip_std > 0.0) {
        double sigma = std::sqrt(S);
        double lim = p->clip_std * sigma;
        if (y > lim) y = lim;
        if (y < -lim) y = -lim;
    }
    double K0 = P00p / S, K1 = P10p / S, K2 = P20p / S, K3 = P30p / S;

    s->pos = x0p + K0 * y;
    s->vel = x1p + K1 * y;
    s->acc = x2p + K2 * y;
    s->jerk = x3p + K3 * y;

    double P00n = (1.0 - K0) * P00p, P01n = (1.0 - K0) * P01p, P02n = (1.0 - K0) * P02p, P03n = (1.0 - K0) * P03p;
    double P10n = P10p - K1 * P00p, P11n = P11p - K1 * P01p, P12n = P12p - K1 * P02p, P13n = P13p - K1 * P03p;
    double P20n = P20p - K2 * P00p, P21n = P21p - K2 * P01p, P22n = P22p - K2 * P02p, P23n = P23p - K2 * P03p;
    double P30n = P30p - K3 * P00p, P31n = P31p - K3 * P01p, P32n = P32p - K3 * P02p, P33n = P33p - K3 * P03p;

    s->P[0][0] = std::fmax(1e-12, P00n); s->P[0][1] = P01n; s->P[0][2] = P02n; s->P[0][3] = P03n;
    s->P[1][0] = P10n; s->P[1][1] = std::fmax(1e-12, P11n); s->P[1][2] = P12n; s->P[1][3] = P13n;
    s->P[2][0] = P20n; s->P[2][1] = P21n; s->P[2][2] = std::fmax(1e-12, P22n); s->P[2][3] = P23n;
    s->P[3][0] = P30n; s->P[3][1] = P31n; s->P[3][2] = P32n; s->P[3][3] = std::fmax(1e-12, P33n);

    double out = s->pos;
    if (p->ema_blend_period > 0.0) {
        double alpha = 2.0 / (p->ema_blend_period + 1.0);
        if (!s->ema_ready) { s->ema_prev = out; s->ema_ready = 1; }
        s->ema_prev = alpha * out + (1.0 - alpha) * s->ema_prev;
        out = s->ema_prev;
    }
    return out;
}

// bar-loop order L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3354-3360
void oracle_kalman4d_series(const double* z, int count, const oracle_kalman4d_params* p, double* out) {
    oracle_kalman4d_state s;
    std::memset(&s, 0, sizeof s);
    for (int i = 0; i < count; i++) {
        if (!s.ready) oracle_kalman4d_reset(&s, p, z[i]);
        out[i] = oracle_kalman4d_step(&s, p, z[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// A10 L/WaveSpecZZ_1.0.4-kalman.mq5:93-101 (reset), :194-231 (update)
void oracle_wkalman_reset(oracle_wkalman_state* s, double init_variance) {
    for (int i = 0; i < 32; i++) { s->weights[i] = 0.0; s->cov[i] = std::fmax(1e-6, init_variance); }
}

double oracle_wkalman_update(oracle_wkalman_state* s, const double* cycle_vals, int cycle_count,
                             double measurement, double q, double r) {
    const double Q = std::fmax(1e-9, q);
    const double R = std::fmax(1e-9, r);
    if (cycle_count > 32) cycle_count = 32;
    double residual = measurement;
    double innovation = R;
    double cov_tmp[32], weight_tmp[32];
    for (int i = 0; i < cycle_count; i++) {
        s->cov[i] += Q;
        cov_tmp[i] = s->cov[i];
        weight_tmp[i] = s->weights[i];
        residual -= cycle_vals[i] * weight_tmp[i];
        innovation += cycle_vals[i] * cycle_vals[i] * cov_tmp[i];
    }
    if (innovation < 1e-9) innovation = R;
    double blended = 0.0;
    for (int i = 0; i < cycle_count; i++) {
        const double H = cycle_vals[i];
        const double cov = cov_tmp[i];
        const double K = (cov * H) / innovation;
        const double new_w = weight_tmp[i] + K * residual;
        const double new_cov = (1.0 - K * H) * cov;
        s->weights[i] = new_w;
        s->cov[i] = std::fmax(new_cov, 1e-9);
        blended += new_w * H;
    }
    return blended;
}

// ---------------------------------------------------------------------------------------------
// A11 L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:387-502
namespace {
struct PlaCtx {
    const double* y;
    int max_segments;
    double max_error;
    int count;
    int* starts; int* ends; double* slopes; double* intercepts;
};

void pla_fit(const double* y, int s, int e, double& slope, double& intercept) {   // :387-417
    const int n = e - s + 1;
    if (n <= 1) { slope = 0.0; intercept = y[s]; return; }
    double sx = 0.0, sy = 0.0, sx2 = 0.0, sxy = 0.0;
    for (int i = s; i <= e; ++i) {
        const double x = (double)i, v = y[i];
        sx += x; sy += v; sx2 += x * x; sxy += x * v;
    }
    const double denom = (double)n * sx2 - sx * sx;
    if (std::fabs(denom) < 1e-9) { slope = 0.0; intercept = sy / (double)n; }
    else {
        slope = ((double)n * sxy - sx * sy) / denom;
        intercept = (sy - slope * sx) / (double)n;
    }
}

double pla_error(const double* y, int s, int e, double slope, double intercept, int& worst) { // :419-440
    double mx = 0.0;
    worst = s;
    for (int i = s; i <= e; ++i) {
        const double approx = slope * (double)i + intercept;
        const double err = std::fabs(y[i] - approx);
        if (err > mx) { mx = err; worst = i; }
    }
    return mx;
}

void pla_append(PlaCtx& c, int s, int e, double slope, double intercept) {   // :374-385
    if (e < s) return;
    c.starts[c.count] = s; c.ends[c.count] = e; c.slopes[c.count] = slope; c.intercepts[c.count] = intercept;
    c.count++;
}

void pla_split(PlaCtx& c, int s, int e) {   // :442-472
    if (s >= e) { pla_append(c, s, e, 0.0, c.y[s]); return; }
    double slope = 0.0, intercept = 0.0;
    pla_fit(c.y, s, e, slope, intercept);
    int worst = s;
    double error = pla_error(c.y, s, e, slope, intercept, worst);
    const bool can_split = (c.count + 2) <= c.max_segments && (e - s) > 1;
    if (can_split && error > c.max_error) {
        const int left_end = (s > worst - 1) ? s : worst - 1;
        const int right_start = (e < worst) ? e : worst;
        pla_split(c, s, left_end);
        pla_split(c, right_start, e);
    } else {
        pla_append(c, s, e, slope, intercept);
    }
}
}  // namespace

int oracle_pla_build(const double* window, int n, int max_segments, double max_error, double* line,
                     int* seg_starts, int* seg_ends, double* seg_slopes, double* seg_intercepts) {
    // :474-502.  The segment list can exceed max_segments by the recursion's look-ahead; the
    // reference sizes its arrays at 2*max_segments and grows on demand, so callers pass
    // scratch of at least n entries.
    std::vector<int> st(n + 4), en(n + 4);
    std::vector<double> sl(n + 4), ic(n + 4);
    PlaCtx c{window, max_segments < 1 ? 1 : max_segments, max_error < 1e-8 ? 1e-8 : max_error, 0,
             st.data(), en.data(), sl.data(), ic.data()};
    pla_split(c, 0, n - 1);
    if (c.count <= 0) return 0;
    for (int s = 0; s < c.count; ++s)
        for (int i = st[s]; i <= en[s] && i < n; ++i) line[i] = sl[s] * (double)i + ic[s];
    for (int s = 0; s < c.count; ++s) {
        if (seg_starts) seg_starts[s] = st[s];
        if (seg_ends) seg_ends[s] = en[s];
        if (seg_slopes) seg_slopes[s] = sl[s];
        if (seg_intercepts) seg_intercepts[s] = ic[s];
    }
    return c.count;
}

// ---------------------------------------------------------------------------------------------
// A12 R/WaveSpecZZ_1.1.0-gpuopt.mq5:393-451
void oracle_zigzag_feed_110(const double* main_ch, const double* high_ch, const double* low_ch,
                            int len, int mode, double high0, double low0, double* feed) {
    int last_ext = -1;
    double last_val = 0.0;
    for (int k = 0; k < len; k++) if (main_ch[k] != 0.0) { last_ext = k; last_val = main_ch[k]; break; }
    if (last_ext == -1) last_val = (high0 + low0) * 0.5;
    std::vector<int> ext_pos; std::vector<double> ext_val;
    for (int j = 0; j < len; ++j) {
        double v = last_val;
        switch (mode) {
            case 0:
                if (main_ch[j] != 0.0) { last_ext = j; last_val = main_ch[j]; }
                v = last_val;
                break;
            case 1: {
                ext_pos.clear(); ext_val.clear();
                for (int k = 0; k < len; k++) if (main_ch[k] != 0.0) { ext_pos.push_back(k); ext_val.push_back(main_ch[k]); }
                int n = (int)ext_pos.size();
                if (n == 0) v = last_val;
                else if (j <= ext_pos[0]) v = ext_val[0];
                else if (j >= ext_pos[n - 1]) v = ext_val[n - 1];
                else {
                    int kseg = -1;
                    for (int kk = 0; kk < n - 1; kk++) if (j >= ext_pos[kk] && j < ext_pos[kk + 1]) { kseg = kk; break; }
                    if (kseg == -1) v = ext_val[n - 1];
                    else {
                        int a = ext_pos[kseg], b = ext_pos[kseg + 1];
                        double va = ext_val[kseg], vb = ext_val[kseg + 1];
                        double t = (double)(j - a) / (double)(b - a);
                        v = va + (vb - va) * t;
                    }
                }
                break;
            }
            case 2:
                v = (high_ch[j] + low_ch[j]) * 0.5;
                break;
        }
        feed[j] = v;
    }
}

// A12 L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:237-357 (current-timeframe branch)
// Applied price, Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3308-3316 (per-bar form of the
// per-window loops: the value of bar start_pos + j does not depend on the window).
int oracle_applied_price(const double* open, const double* high, const double* low, const double* close,
                         long long n, int mode, double* out) {
    for (long long j = 0; j < n; j++) {
        switch (mode) {
            case 1: out[j] = close[j]; break;                                         // :3310 ArrayCopy
            case 2: out[j] = open[j]; break;                                          // :3311
            case 3: out[j] = high[j]; break;                                          // :3312
            case 4: out[j] = low[j]; break;                                           // :3313
            case 5: out[j] = (high[j] + low[j]) / 2.0; break;                         // :3314
            case 6: out[j] = (high[j] + low[j] + close[j]) / 3.0; break;              // :3315
            case 7: out[j] = (high[j] + low[j] + 2 * close[j]) / 4.0; break;          // :3316
            default: return -1;
        }
    }
    return 0;
}

int oracle_zigzag_series_legacy(const double* zz_main, const double* zz_high, const double* zz_low,
                                int n, int mode, double* price_data) {
    std::vector<int> pidx(n); std::vector<double> pval(n);
    int pc = 0;
    for (int j = 0; j < n; ++j) {
        double value = zz_main[j];
        if (value == 0.0 || !std::isfinite(value)) {
            if (zz_high[j] != 0.0 && std::isfinite(zz_high[j])) value = zz_high[j];
            else if (zz_low[j] != 0.0 && std::isfinite(zz_low[j])) value = zz_low[j];
        }
        if (value != 0.0 && std::isfinite(value)) { pidx[pc] = j; pval[pc] = value; ++pc; }
    }
    if (pc < 2) return 0;
    int first_idx = pidx[0];
    double first_val = pval[0];
    for (int j = 0; j <= first_idx && j < n; ++j) price_data[j] = first_val;
    if (mode == 0) {
        for (int p = 0; p < pc - 1; ++p) {
            int s = pidx[p], e = pidx[p + 1];
            double sv = pval[p], ev = pval[p + 1];
            int span = e - s;
            if (span <= 0) { price_data[s] = sv; continue; }
            for (int off = 0; off <= span && (s + off) < n; ++off) {
                double t = (double)off / (double)span;
                price_data[s + off] = sv + (ev - sv) * t;
            }
        }
    } else {
        for (int p = 0; p < pc - 1; ++p) {
            int s = pidx[p], e = pidx[p + 1];
            double plateau = pval[p];
            if (e < s) continue;
            for (int i = s; i <= e && i < n; ++i) price_data[i] = plateau;
        }
    }
    int last_idx = pidx[pc - 1];
    double last_val = pval[pc - 1];
    for (int j = last_idx; j < n; ++j) price_data[j] = last_val;
    return 1;
}

// ---------------------------------------------------------------------------------------------
// A8c R/WaveSpecZZ_1.1.0-gpuopt.mq5:1044-1099 + record layout of SaveCycleCache :294-324
void oracle_cycle_cache(const double* cycles, int out_len, int stride, int top_k, int window_len, int hop,
                        int got, double period_seconds, int music_only, int use_music_weights,
                        double min_coherence, double min_score, double min_snr_db, double* out) {
    const double EMPTY = 1.7976931348623157e308;   // EMPTY_VALUE
    // buffer b of slot-pair p lives at out[idx*20 + 2*b + p]  (Wave, Period, Eta, Phase, Energy, Coher, Snr, Score, Eigen, EtaConf)
    for (int64_t i = 0; i < (int64_t)got * 20; i++) out[i] = EMPTY;
    const double two_pi = 6.28318530717958647692;
    for (int c = 0; c < out_len; ++c) {
        const int64_t base = (int64_t)c * stride;
        int method_id = (stride > 14 ? (int)cycles[base + 14] : 0);
        if (music_only && method_id != 1) continue;
        double amp = cycles[base + 0], freq = cycles[base + 1], period = cycles[base + 2], phase = cycles[base + 3];
        double eta_sec = cycles[base + 5];
        double energy = cycles[base + 6], coher = cycles[base + 7], snr = cycles[base + 8], eigen = cycles[base + 10],
               score = cycles[base + 11], etac = cycles[base + 13];
        double w_energy = std::fmax(energy, 0.0);
        double w_coher = std::fmax(coher, 0.0);
        double w_score = std::fmax(score, 0.0);
        double snr_eff = std::fmax(snr, min_snr_db);
        double w_snr = 1.0 / (1.0 + std::pow(10.0, -snr_eff / 10.0));
        double weight_total = use_music_weights ? (w_energy * w_coher * w_score * w_snr) : 1.0;
        if (coher < min_coherence || score < min_score) weight_total = 0.0;
        int window_idx = c / top_k;
        int start_bar = window_idx * hop;
        if (start_bar >= got) continue;
        double omega = two_pi * freq;
        int recon_span = std::min(window_len - 1, got - start_bar - 1);
        int slot = c % top_k;
        const int p = (slot == 0) ? 0 : 1;
        for (int k = 0; k <= recon_span; ++k) {
            int idx = start_bar + k;
            double theta = phase - omega * k;
            double val = amp * weight_total * std::sin(theta);
            double* o = out + (int64_t)idx * 20;
            o[0 + p] = val; o[2 + p] = period; o[4 + p] = std::fmax(eta_sec - k * period_seconds, 0.0);
            o[6 + p] = theta; o[8 + p] = energy; o[10 + p] = coher; o[12 + p] = snr; o[14 + p] = score;
            o[16 + p] = eigen; o[18 + p] = etac;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// A13 L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:1415-1667, bar-loop driver :3450-3504
void oracle_tracker_reset(oracle_tracker_state* st) {      // first-run reset, :3245-3251
    st->count = 0;
    for (int s = 0; s < 12; s++) st->slot[s] = -1;
}

namespace {
bool same_period(double p1, double p2, double tol) {       // IsSamePeriod :1418-1427
    if (p1 <= 0 || p2 <= 0) return false;
    double diff = std::fabs(p1 - p2);
    double avg = (p1 + p2) / 2.0;
    double diff_pct = (diff / avg) * 100.0;
    return diff_pct <= tol;
}
int find_closest(const oracle_tracker_state* st, double period, double tol) {   // :1433-1456
    int best = -1;
    double smallest = 999999;
    for (int i = 0; i < st->count; i++) {
        if (st->bars_inactive[i] > 0) continue;
        double diff = std::fabs(st->period[i] - period);
        if (same_period(period, st->period[i], tol)) {
            if (diff < smallest) { smallest = diff; best = i; }
        }
    }
    return best;
}
}  // namespace

void oracle_tracker_step(oracle_tracker_state* st, const double* spectrum, int n, double min_period,
                         double max_period, double tolerance_pct, int max_inactive_bars,
                         int32_t* slot_index, double* slot_period) {
    const int spectrum_size = n / 2;
    int min_index = (int)std::ceil((double)n / max_period);
    int max_index = (int)std::floor((double)n / min_period);
    // 4c/5: every band bin ascending, match-or-add (:3450-3499)
    for (int j = min_index; j <= max_index && j < spectrum_size; j++) {
        double period = (j > 0) ? (double)n / j : 0;
        if (period <= 0) continue;
        double power = spectrum[j];
        int t = find_closest(st, period, tolerance_pct);
        if (t >= 0) {                                       // UpdateTracker :1461-1471
            st->period[t] = period; st->fft_index[t] = j; st->power[t] = power;
            st->is_active[t] = 1; st->bars_inactive[t] = 0;
        } else if (st->count < ORACLE_TRACKER_CAP) {        // AddTracker :1476-1498
            int c = st->count++;
            st->period[c] = period; st->fft_index[c] = j; st->power[c] = power;
            st->is_active[c] = 1; st->bars_inactive[c] = 0;
        }
    }
    // DeactivateUnseenTrackers :1504-1529
    for (int i = st->count - 1; i >= 0; i--) {
        if (!st->is_active[i]) {
            st->bars_inactive[i]++;
            if (st->bars_inactive[i] >= max_inactive_bars) {
                for (int j = i; j < st->count - 1; j++) {
                    st->period[j] = st->period[j + 1]; st->power[j] = st->power[j + 1];
                    st->fft_index[j] = st->fft_index[j + 1]; st->is_active[j] = st->is_active[j + 1];
                    st->bars_inactive[j] = st->bars_inactive[j + 1];
                }
                st->count--;
            }
        }
    }
    for (int i = 0; i < st->count; i++) st->is_active[i] = 0;
    // UpdateStableSlots :1582-1667
    for (int s = 0; s < 12; s++) {
        int t = st->slot[s];
        if (t < 0 || t >= st->count) st->slot[s] = -1;
    }
    std::vector<int> sorted(st->count), used(st->count, 0);
    for (int i = 0; i < st->count; i++) sorted[i] = i;
    for (int i = 0; i < st->count - 1; i++)
        for (int j = 0; j < st->count - i - 1; j++)
            if (st->power[sorted[j]] < st->power[sorted[j + 1]]) { int t = sorted[j]; sorted[j] = sorted[j + 1]; sorted[j + 1] = t; }
    for (int s = 0; s < 12; s++) {
        int t = st->slot[s];
        if (t >= 0 && t < st->count) {
            used[t] = 1;
            slot_period[s] = st->period[t]; slot_index[s] = st->fft_index[t];
        }
    }
    for (int s = 0; s < 12; s++) {
        if (st->slot[s] >= 0 && st->slot[s] < st->count) continue;
        int chosen = -1;
        for (int k = 0; k < st->count; k++) {
            int idx = sorted[k];
            if (used[idx]) continue;
            chosen = idx;
            break;
        }
        if (chosen != -1) {
            st->slot[s] = chosen; used[chosen] = 1;
            slot_period[s] = st->period[chosen]; slot_index[s] = st->fft_index[chosen];
        } else {
            st->slot[s] = -1;
            slot_period[s] = 0.0; slot_index[s] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Pipelines (bar loops).  Stage order follows L/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:3301-3460;
// the nodetrend/top-8/reconstruction tail follows L/...-gpuopt-nodetrend.mq5:515-568; the
// weight-Kalman tail follows L/WaveSpecZZ_1.0.4-kalman.mq5:254-286.
void oracle_default_cfg(oracle_pipeline_cfg* cfg, int window_len) {
    std::memset(cfg, 0, sizeof *cfg);
    cfg->window_len = window_len; cfg->hop = 1; cfg->top_k = 8; cfg->row_stride = 15;
    cfg->min_period = 18; cfg->max_period = 200; cfg->sample_rate_seconds = 60.0;
    cfg->feed = 0; cfg->detrend = 0; cfg->trend_period = 1024; cfg->window_type = 0;
    cfg->select = 0; cfg->pla_max_segments = 32; cfg->pla_max_error = 0.0005;
    cfg->outputs = 1 | 2 | 4;
    cfg->wk_process_noise = 0.25; cfg->wk_meas_noise = 9.0; cfg->wk_init_variance = 25.0;
    oracle_kalman4d_defaults(&cfg->kalman);
    cfg->tracker_tolerance = 5.0; cfg->tracker_max_inactive = 3; cfg->reserved0 = 0;   // :985-986
}

namespace {
struct WinScratch {
    std::vector<double> price, trend, det, re, im, spec, pw, ph, un, gd;
    std::vector<int> idx;
    explicit WinScratch(int n) : price(n), trend(n), det(n), re(n), im(n), spec(n / 2), pw(n / 2),
                                 ph(n / 2), un(n / 2), gd(n / 2), idx(n / 2) {}
};

// New-build row definition (DESIGN.md "result row"): chosen so that existing consumers,
// which plot amplitude*sin(phase - omega*k) (R/WaveSpecZZ_1.1.0-gpuopt.mq5:1087-1098, :1508-1541;
// L/WaveSpecZZ_1.0.4-old.mq5:2738-2741), render the A8b last-sample contribution.
void fill_row(double* row, int stride, int n, int bin, double power, double re, double im,
              double band_sum, double sample_rate_seconds) {
    double f[15];
    for (int i = 0; i < 15; i++) f[i] = 0.0;
    if (bin > 0) {
        const double nn = (double)(n - 1);
        f[0] = 2.0 * std::sqrt(power) / (double)n;
        f[1] = (double)bin / (double)n;
        f[2] = (double)n / (double)bin;
        double ph = std::atan2(im, re) + 2.0 * kPi * (double)bin * nn / (double)n + 0.5 * kPi;
        ph = std::remainder(ph, 2.0 * kPi);
        f[3] = ph;
        double d = std::fmod(0.5 * kPi - ph, kPi);
        if (d < 0.0) d += kPi;
        f[4] = d / (2.0 * kPi * f[1]);
        f[5] = f[4] * sample_rate_seconds;
        f[6] = band_sum > 0.0 ? power / band_sum : 0.0;
        f[14] = 0.0;
    }
    const int m = stride < 15 ? stride : 15;
    for (int i = 0; i < m; i++) row[i] = f[i];
    for (int i = m; i < stride; i++) row[i] = 0.0;
}

void process_window(const double* series, int64_t w, const oracle_pipeline_cfg* cfg, WinScratch& s,
                    oracle_kalman4d_state* kst, oracle_wkalman_state* wk,
                    double* spectra, double* rows, int32_t* bins, double* waves, double* kalman,
                    double* phase, double* wkalman, oracle_tracker_state* trk = nullptr,
                    int32_t* trk_index = nullptr, double* trk_period = nullptr) {
    const int n = cfg->window_len;
    const int K = cfg->top_k;
    const double* src = series + w * cfg->hop;
    // 1. feed
    bool have_pla = false;
    if (cfg->feed == 1) {
        std::memcpy(s.price.data(), src, sizeof(double) * n);
        std::vector<double> line(n);
        int c = oracle_pla_build(s.price.data(), n, cfg->pla_max_segments, cfg->pla_max_error,
                                 line.data(), nullptr, nullptr, nullptr, nullptr);
        if (c > 0) { std::memcpy(s.price.data(), line.data(), sizeof(double) * n); have_pla = true; }
    }
    if (!have_pla) std::memcpy(s.price.data(), src, sizeof(double) * n);
    // Kalman on the newest sample of the applied price
    if (kalman) {
        double meas = s.price[n - 1];
        if (!kst->ready) oracle_kalman4d_reset(kst, &cfg->kalman, meas);
        kalman[w] = oracle_kalman4d_step(kst, &cfg->kalman, meas);
    }
    // 2. detrend
    if (cfg->detrend == 1) {
        oracle_detrend_iir(s.price.data(), n, cfg->trend_period, s.trend.data(), s.det.data());
    } else if (cfg->detrend == 2) {
        double mean = 0.0;
        for (int i = 0; i < n; ++i) mean += s.price[i];
        mean /= (double)n;
        for (int i = 0; i < n; ++i) s.det[i] = s.price[i] - mean;
    } else {
        std::memcpy(s.det.data(), s.price.data(), sizeof(double) * n);
    }
    // 3. window
    oracle_apply_window(s.det.data(), n, cfg->window_type);
    // 4. FFT + unpack + power
    oracle_fft_forward(s.det.data(), n, s.re.data(), s.im.data());
    if (spectra) {
        double* o = spectra + w * (int64_t)n;
        for (int k = 0; k < n / 2; k++) { o[2 * k] = s.re[k]; o[2 * k + 1] = s.im[k]; }
    }
    oracle_power(s.re.data(), s.im.data(), n, s.spec.data());
    if (trk && trk_index && trk_period)
        oracle_tracker_step(trk, s.spec.data(), n, cfg->min_period, cfg->max_period, cfg->tracker_tolerance,
                            cfg->tracker_max_inactive, trk_index + w * 12, trk_period + w * 12);
    // 5. selection
    int sel_bin[32]; double sel_pow[32];
    for (int i = 0; i < 32; i++) { sel_bin[i] = -1; sel_pow[i] = -1.0; }
    double band_sum = 0.0;
    {
        int lo = (int)std::ceil((double)n / cfg->max_period);
        int hi = (int)std::floor((double)n / cfg->min_period);
        if (hi >= n / 2) hi = n / 2 - 1;
        if (cfg->select == 1 && lo < 1) lo = 1;
        if (lo < 0) lo = 0;
        for (int b = lo; b <= hi; b++) band_sum += s.spec[b];
    }
    if (cfg->select == 1) {
        int c = oracle_collect_sorted(s.re.data(), s.im.data(), n, cfg->min_period, cfg->max_period,
                                      s.idx.data(), s.pw.data());
        for (int i = 0; i < K && i < c; i++) { sel_bin[i] = s.idx[i]; sel_pow[i] = s.pw[i]; }
    } else {
        oracle_topk_insertion(s.spec.data(), n, cfg->min_period, cfg->max_period, K, sel_bin, sel_pow);
    }
    for (int k = 0; k < K; k++) {
        const int b = sel_bin[k];
        if (bins) bins[w * K + k] = b;
        if (rows)
            fill_row(rows + (w * K + k) * (int64_t)cfg->row_stride, cfg->row_stride, n, b, sel_pow[k],
                     b >= 0 ? s.re[b] : 0.0, b >= 0 ? s.im[b] : 0.0, band_sum, cfg->sample_rate_seconds);
        if (waves) {
            double wv = 0.0, per = 0.0;
            if (b >= 0) oracle_recon_last(s.re.data(), s.im.data(), n, b, sel_pow[k], &wv, &per);
            waves[w * K + k] = wv;
        }
    }
    if (wkalman) {
        double vals[32]; int use = 0;
        for (int k = 0; k < K && k < 32; k++)
            if (sel_bin[k] >= 0) vals[use++] = oracle_contribution(s.re.data(), s.im.data(), n, sel_bin[k]);
        if (use <= 0) wkalman[w] = 0.0;
        else wkalman[w] = oracle_wkalman_update(wk, vals, use, src[n - 1], cfg->wk_process_noise, cfg->wk_meas_noise);
    }
    if (phase) {
        const int h = n / 2;
        oracle_phase_chain(s.re.data(), s.im.data(), h, s.ph.data(), s.un.data(), s.gd.data());
        double* o = phase + w * (int64_t)(3 * h);
        std::memcpy(o, s.ph.data(), sizeof(double) * h);
        std::memcpy(o + h, s.un.data(), sizeof(double) * h);
        std::memcpy(o + 2 * h, s.gd.data(), sizeof(double) * h);
    }
}
}  // namespace

void oracle_pipeline_series(const double* series, int series_len, const oracle_pipeline_cfg* cfg,
                            double* spectra, double* rows, int32_t* bins, double* waves,
                            double* kalman, double* phase, double* wkalman) {
    oracle_pipeline_series_trk(series, series_len, cfg, spectra, rows, bins, waves, kalman, phase, wkalman,
                               nullptr, nullptr);
}

void oracle_pipeline_series_trk(const double* series, int series_len, const oracle_pipeline_cfg* cfg,
                                double* spectra, double* rows, int32_t* bins, double* waves,
                                double* kalman, double* phase, double* wkalman,
                                int32_t* trk_index, double* trk_period) {
    const int n = cfg->window_len;
    if (series_len < n) return;
    const int64_t nwin = 1 + (int64_t)(series_len - n) / cfg->hop;
    WinScratch s(n);
    oracle_kalman4d_state kst; std::memset(&kst, 0, sizeof kst);
    oracle_wkalman_state wk; oracle_wkalman_reset(&wk, cfg->wk_init_variance);
    std::vector<oracle_tracker_state> trk(1);
    oracle_tracker_reset(&trk[0]);
    for (int64_t w = 0; w < nwin; w++)
        process_window(series, w, cfg, s, &kst, &wk, spectra, rows, bins, waves, kalman, phase, wkalman,
                       &trk[0], trk_index, trk_period);
}

int64_t oracle_pipeline_batch_mt(const double* series, int n_series, int series_len,
                                 const oracle_pipeline_cfg* cfg, int threads, int64_t max_windows,
                                 double* spectra, double* rows, int32_t* bins, double* waves) {
    const int n = cfg->window_len;
    if (series_len < n) return 0;
    int64_t nwin = 1 + (int64_t)(series_len - n) / cfg->hop;
    const int64_t full_nwin = nwin;
    if (max_windows > 0 && max_windows < nwin) nwin = max_windows;
    const int64_t chunk = 2048;
    const int64_t chunks_per_series = (nwin + chunk - 1) / chunk;
    const int64_t total_chunks = chunks_per_series * n_series;
    std::atomic<int64_t> next{0};
    if (threads < 1) threads = 1;
    auto worker = [&]() {
        WinScratch s(n);
        for (;;) {
            int64_t c = next.fetch_add(1);
            if (c >= total_chunks) break;
            const int64_t sidx = c / chunks_per_series;
            const int64_t w0 = (c % chunks_per_series) * chunk;
            const int64_t w1 = (w0 + chunk < nwin) ? w0 + chunk : nwin;
            const double* ser = series + sidx * (int64_t)series_len;
            double* sp = spectra ? spectra + sidx * full_nwin * n : nullptr;
            double* rw = rows ? rows + sidx * full_nwin * cfg->top_k * cfg->row_stride : nullptr;
            int32_t* bn = bins ? bins + sidx * full_nwin * cfg->top_k : nullptr;
            double* wv = waves ? waves + sidx * full_nwin * cfg->top_k : nullptr;
            for (int64_t w = w0; w < w1; w++)
                process_window(ser, w, cfg, s, nullptr, nullptr, sp, rw, bn, wv, nullptr, nullptr, nullptr);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return nwin * n_series;
}

}  // extern "C"
