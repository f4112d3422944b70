// ws_inverse.cu — batched inverse real FFT behind gpu_fft_real_inverse
// (declared Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:27, used Legacy/WaveSpecZZ_1.0.4-core.mq5:426,
// Legacy/WaveSpecZZ_1.0.4-parallel.mq5:1273) and the "inverse-FFT reconstruction of the selected
// cycles" of the north star: the same transform fed with a spectrum masked to the top-K bins.
//
// Input per window: N/2 interleaved bins as gpu_fft_real_forward / the spectra plane hold them
// (Nyquist bin absent -> 0, Im X[0] ignored).  Output: N real samples, normalised by 1/N (design
// decision, see include/wavespec_abi.h).
//
// One WARP per window, on the in-place radix-8 machinery of the forward per-window kernel
// (ws_warpfft_core.cuh): with M = N/2 and Z = E + iO the spectrum of z[m] = x[2m] + i x[2m+1],
//     E[k] = (X[k] + conj X[M-k]) / 2,   O[k] = (X[k] - conj X[M-k]) / 2 * e^{+2 pi i k / N},
// z = conj(FFT(conj Z)) / M.  A lane forms conj Z[k] and conj Z[M-k] from the pair (X[k], X[M-k]) it
// reads from HBM (both runs coalesced) straight into the warp-private swizzled array, the DIF passes
// run in place with one __syncwarp between them, and the digit-reversed result leaves as contiguous
// 16-byte stores.  HBM bound by construction: 8 N bytes in, 8 N bytes out per window.
#include "ws_common.cuh"
#include "ws_series.h"
#include "ws_warpfft_core.cuh"

namespace ws {

template <int LN>
__device__ __forceinline__ void inverse_window(const double2* __restrict__ X, const int32_t* __restrict__ bins, int K,
                                               const double2* __restrict__ tw, double2* Z, unsigned* mask,
                                               double2* __restrict__ out, int lane) {
    typedef ws_wf::Geo<LN> G;
    constexpr int M = G::M, H = M / 2;
    const bool masked = bins != nullptr;
    if (masked) {
        // bit b of the mask: bin b is one of the window's selected cycles
        for (int i = lane; i < (M + 31) / 32; i += 32) mask[i] = 0u;
        __syncwarp();
        if (lane < K) { const int b = bins[lane]; if (b >= 0 && b < M) atomicOr(&mask[b >> 5], 1u << (b & 31)); }
        __syncwarp();
    }
    auto keep = [&](int b) { return !masked || ((mask[b >> 5] >> (b & 31)) & 1u); };
    const double2 zero = make_double2(0.0, 0.0);
    // conj Z[k], conj Z[M-k] from the pair (X[k], X[M-k]); k = 0 and k = M/2 are self-paired
    for (int k = lane; k <= H; k += 32) {
        const int km = (M - k) & (M - 1);
        double2 xa = keep(k) ? X[k] : zero;
        const double2 xb = (k == 0) ? zero : (keep(km) ? X[km] : zero);      // X[M] (Nyquist) := 0
        if (k == 0) xa.y = 0.0;                                               // X[0] of a real signal is real
        ws_wf::inverse_unpack<LN>(k, xa, xb, tw, Z);
    }
    __syncwarp();
    ws_wf::dif_pass<LN, G::radix(0), G::stride(0), false>(lane, 0, Z, tw);
    ws_wf::later_passes<LN, 1>(lane, Z, tw, [] { __syncwarp(); });
    __syncwarp();
    for (int m = lane; m < M; m += 32) __stcs(out + m, ws_wf::inverse_pair<LN>(Z, m));
    __syncwarp();
}

template <int LN, int W>
__global__ void __launch_bounds__(32 * W)
inverse_real_warp_kernel(const double* __restrict__ spec, const int32_t* __restrict__ bins, int32_t K,
                         int64_t n_windows, const double2* __restrict__ tw, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = 1 << LN, M = N / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* Z = reinterpret_cast<double2*>(smem_raw) + (size_t)warp * M;
    unsigned* mask = reinterpret_cast<unsigned*>(smem_raw + (size_t)W * M * 16) + warp * ((M + 31) / 32);
    for (int64_t w = (int64_t)blockIdx.x * W + warp; w < n_windows; w += (int64_t)gridDim.x * W)
        inverse_window<LN>(reinterpret_cast<const double2*>(spec) + w * M, bins ? bins + w * K : nullptr, K, tw, Z, mask,
                           reinterpret_cast<double2*>(out) + w * M, lane);
}

template <int LN, int W>
static cudaError_t launch_inv(const double* d_spec, const int32_t* d_bins, int32_t K, int64_t n_windows,
                              const double2* tw, double* d_out, cudaStream_t stream) {
    constexpr int M = (1 << LN) / 2;
    const size_t smem = (size_t)W * M * 16 + (size_t)W * ((M + 31) / 32) * 4;
    static std::atomic<unsigned long long> attr_seen{0};
    if (first_launch_on_device(attr_seen)) {
        cudaError_t e = cudaFuncSetAttribute(inverse_real_warp_kernel<LN, W>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
    }
    int64_t blocks = (n_windows + W - 1) / W;
    const int64_t cap = 148 * 16;                    // grid-stride beyond a few waves of CTAs
    if (blocks > cap) blocks = cap;
    inverse_real_warp_kernel<LN, W><<<(unsigned)blocks, 32 * W, smem, stream>>>(d_spec, d_bins, K, n_windows, tw, d_out);
    return cudaGetLastError();
}

cudaError_t launch_inverse_real(const double* d_spec, int32_t N, int64_t n_windows, const double2* tw,
                                double* d_out, cudaStream_t stream, const int32_t* d_bins, int32_t K) {
    switch (N) {
        case 4:    return launch_inv<2, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 8:    return launch_inv<3, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 16:   return launch_inv<4, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 32:   return launch_inv<5, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 64:   return launch_inv<6, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 128:  return launch_inv<7, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 256:  return launch_inv<8, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 512:  return launch_inv<9, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 1024: return launch_inv<10, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 2048: return launch_inv<11, 8>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 4096: return launch_inv<12, 4>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        case 8192: return launch_inv<13, 2>(d_spec, d_bins, K, n_windows, tw, d_out, stream);
        default:   return cudaErrorInvalidValue;
    }
}

}  // namespace ws
