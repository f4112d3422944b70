// ws_epilogue_cta.cuh — CTA-wide top-K epilogue of the sliding kernel (insertion rule A7a,
// Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:537-554) for bands of <= 128 bins.
//
// The captured band of the tile's T windows (xb[T][band], complex) is reduced in phases that keep
// every thread busy, with no dependent chains and no warp shuffles:
//   R0  powers pw[T][band];
//   R1  thread = (window, bin): RANK of the bin among the window's band under the reference's
//       order — a bin is beaten by every strictly larger power and by every EQUAL power at a lower
//       bin (strict '>' insertion keeps the earlier bin ahead).  rank < K  <=>  the insertion
//       loop of the reference leaves this bin in slot `rank`.  band independent compares per
//       thread, all loads are shared-memory broadcasts;
//   R2  thread = (window, slot): row arithmetic (sqrt, atan2), rows staged in shared memory and
//       streamed out as contiguous 128-bit stores.
#pragma once
#include "ws_common.cuh"
#include "ws_epilogue.cuh"

namespace ws {

struct CtaEpiLayout {          // byte offsets inside the overlay region (dead work area)
    int pw_off, selv_off, selp_off, bsum_off, stage_off, total;
};

__host__ __device__ inline CtaEpiLayout cta_epi_layout(int T, int band, int K, int row_stride, bool rows) {
    CtaEpiLayout l;
    int o = 0;
    l.pw_off = o;   o += T * band * 8;
    l.selv_off = o; o += T * K * 8;
    l.bsum_off = o; o += T * 8;
    l.selp_off = o; o += T * K * 4;
    o = (o + 15) & ~15;
    l.stage_off = o;
    if (rows && row_stride <= 16) o += T * K * row_stride * 8;
    l.total = (o + 15) & ~15;
    return l;
}

// xb      : shared, [T][band] complex captured bins (band-relative index), not overlaid
// ov      : shared overlay region of cta_epi_layout(...).total bytes (16-byte aligned)
// nthreads: blockDim.x
__device__ __forceinline__ void cta_select_emit(const Params& p, const double2* xb, int band, int lo, int T,
                                                int nvalid, int64_t gw_tile, unsigned char* ov, int nthreads) {
    const int tid = threadIdx.x;
    const int K = p.K, N = p.N;
    const CtaEpiLayout L = cta_epi_layout(T, band, K, p.row_stride, p.rows != nullptr);
    double* pw = reinterpret_cast<double*>(ov + L.pw_off);
    double* selv = reinterpret_cast<double*>(ov + L.selv_off);
    int* selp = reinterpret_cast<int*>(ov + L.selp_off);
    double* bsum = reinterpret_cast<double*>(ov + L.bsum_off);
    double* stage = reinterpret_cast<double*>(ov + L.stage_off);

    // ---- R0: powers; empty selection
    for (int i = tid; i < nvalid * band; i += nthreads) { const double2 v = xb[i]; pw[i] = v.x * v.x + v.y * v.y; }
    for (int i = tid; i < nvalid * K; i += nthreads) { selp[i] = -1; selv[i] = -1.0; }
    __syncthreads();

    // ---- R1: rank of every (window, bin); the first T threads also sum the band serially
    if (tid < nvalid) {
        const double* pwin = pw + tid * band;
        double bs = 0.0;
        for (int j = 0; j < band; j++) bs += pwin[j];
        bsum[tid] = bs;
    }
    for (int t = tid; t < nvalid * band; t += nthreads) {
        const int wl = t / band, e = t - wl * band;
        const double* pwin = pw + wl * band;
        const double v = pwin[e];
        int rank = 0;
        if (v == v) {                                  // NaN never enters the list
            for (int j = 0; j < e; j++) rank += (pwin[j] >= v) ? 1 : 0;         // equal power at a lower bin wins
            for (int j = e + 1; j < band; j++) rank += (pwin[j] > v) ? 1 : 0;
            if (rank < K) { selp[wl * K + rank] = e; selv[wl * K + rank] = v; }
        }
    }
    __syncthreads();

    // ---- R2: rows / bins / waves / contributions, one (window, slot) per thread
    const int rs = p.row_stride;
    const bool staged = p.rows && rs <= 16;
    for (int t = tid; t < nvalid * K; t += nthreads) {
        const int wl = t / K;
        const int pos = selp[t];
        const double pw = selv[t];
        const int bin = pos >= 0 ? lo + pos : -1;
        double re = 0.0, im = 0.0;
        if (pos >= 0) { const double2 x = xb[wl * band + pos]; re = x.x; im = x.y; }
        const int64_t slot = gw_tile * K + t;
        if (p.bins) p.bins[slot] = bin;
        const double nn = (double)(N - 1);
        if (p.waves) {
            double wv = 0.0;
            if (bin > 0) {
                double mag = sqrt(pw);
                double ph = atan2(im, re);
                wv = (mag / (double)N) * cos(ph + 2.0 * kPi * (double)bin * nn / (double)N);
            }
            p.waves[slot] = wv;
        }
        if (p.contrib) {
            double cv = 0.0;
            if (bin >= 0) {
                double sn, cs;
                sincos(2.0 * kPi * bin * nn / N, &sn, &cs);
                cv = (2.0 / N) * (re * cs - im * sn);
            }
            p.contrib[slot] = cv;
        }
        if (p.rows) {
            double f[kRowFields];
#pragma unroll
            for (int i = 0; i < kRowFields; i++) f[i] = 0.0;
            if (bin > 0) {
                f[0] = 2.0 * sqrt(pw) / (double)N;
                f[1] = (double)bin / (double)N;
                f[2] = (double)N / (double)bin;
                // phase at the newest sample: atan2 + 2 pi k (N-1)/N + pi/2 wrapped to [-pi, pi];
                // 2 pi k (N-1)/N == -2 pi k / N (mod 2 pi) keeps the wrap to a single step
                double ph = atan2(im, re) + (0.5 * kPi - 2.0 * kPi * (double)bin / (double)N);
                if (ph > kPi) ph -= 2.0 * kPi;
                if (ph < -kPi) ph += 2.0 * kPi;
                f[3] = ph;
                double d = 0.5 * kPi - ph;                  // bars to the next extremum of amp*sin
                if (d < 0.0) d += kPi;
                if (d >= kPi) d -= kPi;
                f[4] = d / (2.0 * kPi * f[1]);
                f[5] = f[4] * p.sample_rate_seconds;
                const double bs = bsum[wl];
                f[6] = bs > 0.0 ? pw / bs : 0.0;
            }
            double* row = staged ? stage + t * rs : p.rows + slot * (int64_t)rs;
#pragma unroll
            for (int i = 0; i < kRowFields; i++) if (i < rs) row[i] = f[i];
            for (int i = kRowFields; i < rs; i++) row[i] = 0.0;
        }
    }
    if (staged) {
        __syncthreads();
        const int total = nvalid * K * rs;
        const int64_t base = gw_tile * K * (int64_t)rs;
        double* dst = p.rows + base;
        if (((base | total) & 1) == 0) {
            const double2* s2 = reinterpret_cast<const double2*>(stage);
            double2* d2 = reinterpret_cast<double2*>(dst);
            for (int i = tid; i < (total >> 1); i += nthreads) __stcs(d2 + i, s2[i]);
        } else {
            for (int i = tid; i < total; i += nthreads) __stcs(dst + i, stage[i]);
        }
    }
}

}  // namespace ws
