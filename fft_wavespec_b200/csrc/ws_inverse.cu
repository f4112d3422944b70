// ws_inverse.cu — inverse real FFT behind gpu_fft_real_inverse
// (declared Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:27, used Legacy/WaveSpecZZ_1.0.4-core.mq5:426).
// Input: N/2 interleaved bins as gpu_fft_real_forward writes them (Nyquist bin absent -> 0).
// Output: N real samples, normalised by 1/N (design decision, see include/wavespec_abi.h).
//
// One CTA per window: rebuild the N/2-point spectrum Z = E + iO of z[m] = x[2m] + i x[2m+1]
// from the Hermitian half spectrum, run a forward Stockham radix-2 FFT on conj(Z) in shared
// memory, conjugate and scale (IFFT(Z) = conj(FFT(conj Z))/M).
#include "ws_common.cuh"
#include "ws_series.h"

namespace ws {

__global__ void __launch_bounds__(128)
inverse_real_kernel(const double* __restrict__ spec, int32_t N, int32_t log2N,
                    const double2* __restrict__ tw, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = N >> 1;
    double2* A = reinterpret_cast<double2*>(smem_raw);
    double2* B = A + M;
    const int tid = threadIdx.x;
    const double2* X = reinterpret_cast<const double2*>(spec + (int64_t)blockIdx.x * N);
    double* o = out + (int64_t)blockIdx.x * N;

    for (int k = tid; k < M; k += blockDim.x) {
        double2 xk = X[k];
        double2 xm = (k == 0) ? make_double2(0.0, 0.0) : cconj(X[M - k]);   // conj X[M-k]; X[M] := 0
        double2 E = make_double2(0.5 * (xk.x + xm.x), 0.5 * (xk.y + xm.y));
        double2 D = make_double2(0.5 * (xk.x - xm.x), 0.5 * (xk.y - xm.y));
        double2 O = cmul(D, cconj(__ldg(tw + k)));                            // * e^{+2 pi i k/N}
        // Z = E + iO ; store conj(Z)
        A[k] = make_double2(E.x - O.y, -(E.y + O.x));
    }
    __syncthreads();
    double2* in = A; double2* ob = B;
    const int h = M >> 1;
    for (int Ns = 1; Ns < M; Ns <<= 1) {
        const int tstep = N / (2 * Ns);
        for (int j = tid; j < h; j += blockDim.x) {
            int k = j & (Ns - 1);
            double2 a0 = in[j];
            double2 a1 = cmul(in[j + h], __ldg(tw + tstep * k));
            int base = ((j - k) << 1) + k;
            ob[base] = cadd(a0, a1);
            ob[base + Ns] = csub(a0, a1);
        }
        __syncthreads();
        double2* t = in; in = ob; ob = t;
    }
    const double sc = 1.0 / (double)M;
    for (int m = tid; m < M; m += blockDim.x) {
        double2 z = in[m];
        o[2 * m] = z.x * sc;
        o[2 * m + 1] = -z.y * sc;
    }
}

cudaError_t launch_inverse_real(const double* d_spec, int32_t N, int32_t n_windows, const double2* tw,
                                double* d_out, cudaStream_t stream) {
    int log2N = 0;
    while ((1 << log2N) < N) log2N++;
    size_t smem = (size_t)N * 16;   // two buffers of N/2 double2
    static unsigned long long attr_seen = 0;
    if (first_launch_on_device(attr_seen)) {
        cudaError_t e = cudaFuncSetAttribute(inverse_real_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
    }
    if (smem > 232448) return cudaErrorInvalidValue;
    inverse_real_kernel<<<n_windows, 128, smem, stream>>>(d_spec, N, log2N, tw, d_out);
    return cudaGetLastError();
}

}  // namespace ws
