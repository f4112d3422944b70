// ws_cache.cu — device-side decode of the batch result into the per-bar cycle cache record
// (SURVEY.md section 8f rank 2).
//
// What the indicator does on the CPU after a batch warm-up (WaveSpecZZ_1.1.0-gpuopt.mq5:1067-1099):
// for every result row, in row order, it back-propagates the sine over up to N bars
//     Wave[start_bar + k] = amp * weight * sin(phase - 2 pi freq k),  k = 0 .. recon_span
// into per-bar buffers, later rows overwriting earlier ones — O(rows x N) sin() calls — and then
// SaveCycleCache (:294-324) writes 20 doubles per bar.  Only the LAST writer of a bar survives, and
// that writer is known in closed form: for buffer 1 it is slot 0 of the last window starting at or
// before the bar (k = bar - start), for buffer 2 the highest slot >= 1 of that window (every slot
// >= 1 lands in buffer 2, :1093-1094).  So the record of a bar is an O(1) function of at most two
// rows: one thread per bar, 20 coalesced doubles out, no back-propagation loop at all.  Rows the
// indicator would skip (InpMusicOnly and method != 1, :1071) are searched past exactly as the
// sequential loop would leave them: the previous window that still covers the bar wins.
#include "../../include/wavespec_abi.h"
#include "ws_common.cuh"
#include "ws_series.h"

namespace ws {

constexpr double kEmpty = 1.7976931348623157e308;     // MQL5 EMPTY_VALUE = DBL_MAX
constexpr double kTwoPi = 6.28318530717958647692;     // the literal of :1064

struct CacheRow { bool ok; double val, period, eta, theta, energy, coher, snr, score, eigen, etac; };

__device__ __forceinline__ bool row_skipped(const double* row, int stride, int music_only) {
    const int method_id = stride > 14 ? (int)row[14] : 0;
    return music_only && method_id != 1;
}

__device__ __forceinline__ CacheRow decode(const double* row, int k, double period_seconds,
                                           const wavespec_cache_params cp) {
    CacheRow r;
    const double amp = row[0], freq = row[1], period = row[2], phase = row[3];
    const double eta_sec = row[5];
    const double energy = row[6], coher = row[7], snr = row[8], eigen = row[10], score = row[11], etac = row[13];
    const double w_energy = fmax(energy, 0.0), w_coher = fmax(coher, 0.0), w_score = fmax(score, 0.0);
    const double snr_eff = fmax(snr, cp.min_snr_db);
    const double w_snr = 1.0 / (1.0 + pow(10.0, -snr_eff / 10.0));
    double weight_total = cp.use_music_weights ? (w_energy * w_coher * w_score * w_snr) : 1.0;
    if (coher < cp.min_coherence || score < cp.min_score) weight_total = 0.0;
    const double omega = kTwoPi * freq;
    const double theta = phase - omega * k;
    r.ok = true;
    r.val = amp * weight_total * sin(theta);
    r.period = period;
    r.eta = fmax(eta_sec - k * period_seconds, 0.0);
    r.theta = theta;
    r.energy = energy; r.coher = coher; r.snr = snr; r.score = score; r.eigen = eigen; r.etac = etac;
    return r;
}

__global__ void cycle_cache_kernel(const double* __restrict__ rows, int64_t n_windows, int32_t top_k, int32_t stride,
                                   int32_t N, int32_t hop, int64_t bars, int64_t bar0, int64_t nbars,
                                   double period_seconds, const wavespec_cache_params cp, double* __restrict__ out) {
    const int64_t idx = bar0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= bar0 + nbars || idx >= bars) return;
    CacheRow b1, b2;
    b1.ok = false; b2.ok = false;
    int64_t w = idx / hop;
    if (w > n_windows - 1) w = n_windows - 1;
    // windows that still cover this bar, newest first
    for (; w >= 0 && !(b1.ok && b2.ok); --w) {
        const int64_t start = w * hop;
        if (start >= bars) continue;
        int64_t span = bars - start - 1;
        if (span > N - 1) span = N - 1;
        const int64_t k = idx - start;
        if (k > span) break;                      // older windows start even earlier: none covers the bar
        const double* wr = rows + w * (int64_t)top_k * stride;
        if (!b1.ok && !row_skipped(wr, stride, cp.music_only)) b1 = decode(wr, (int)k, period_seconds, cp);
        if (!b2.ok)
            for (int s = top_k - 1; s >= 1; --s) {
                const double* r = wr + (int64_t)s * stride;
                if (!row_skipped(r, stride, cp.music_only)) { b2 = decode(r, (int)k, period_seconds, cp); break; }
            }
        if (top_k < 2 && b1.ok) break;
    }
    double* o = out + idx * 20;
    const double e = kEmpty;
    o[0] = b1.ok ? b1.val : e;      o[1] = b2.ok ? b2.val : e;
    o[2] = b1.ok ? b1.period : e;   o[3] = b2.ok ? b2.period : e;
    o[4] = b1.ok ? b1.eta : e;      o[5] = b2.ok ? b2.eta : e;
    o[6] = b1.ok ? b1.theta : e;    o[7] = b2.ok ? b2.theta : e;
    o[8] = b1.ok ? b1.energy : e;   o[9] = b2.ok ? b2.energy : e;
    o[10] = b1.ok ? b1.coher : e;   o[11] = b2.ok ? b2.coher : e;
    o[12] = b1.ok ? b1.snr : e;     o[13] = b2.ok ? b2.snr : e;
    o[14] = b1.ok ? b1.score : e;   o[15] = b2.ok ? b2.score : e;
    o[16] = b1.ok ? b1.eigen : e;   o[17] = b2.ok ? b2.eigen : e;
    o[18] = b1.ok ? b1.etac : e;    o[19] = b2.ok ? b2.etac : e;
}

cudaError_t launch_cycle_cache(const double* rows, int64_t n_windows, int32_t top_k, int32_t stride, int32_t N,
                               int32_t hop, int64_t bars, int64_t bar0, int64_t nbars, double period_seconds,
                               const wavespec_cache_params& cp, double* out, cudaStream_t stream) {
    if (nbars < 1) return cudaSuccess;
    const unsigned blocks = (unsigned)((nbars + 255) / 256);
    cycle_cache_kernel<<<blocks, 256, 0, stream>>>(rows, n_windows, top_k, stride, N, hop, bars, bar0, nbars,
                                                   period_seconds, cp, out);
    return cudaGetLastError();
}

}  // namespace ws
