// Micro-benchmark: per-SM throughput of FP64 arithmetic vs FP64 compare/select vs 64-bit integer
// compare on B200, to size the top-K epilogue.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b = seed * 0.5;
    long long i0 = __double_as_longlong(a0), i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, ib = __double_as_longlong(b);
    int c = 0;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {          // DFMA
            a0 = fma(a0, b, a1); a1 = fma(a1, b, a2); a2 = fma(a2, b, a3); a3 = fma(a3, b, a4);
            a4 = fma(a4, b, a5); a5 = fma(a5, b, a6); a6 = fma(a6, b, a7); a7 = fma(a7, b, a0);
        } else if (MODE == 1) {   // DSETP + integer add (compare, count)
            c += (a0 > b); c += (a1 > b); c += (a2 > b); c += (a3 > b);
            c += (a4 > b); c += (a5 > b); c += (a6 > b); c += (a7 > b);
            b += 1e-300 * c;      // keep the compares live
        } else if (MODE == 2) {   // max via compare+select (DSETP + 2 SEL)
            a0 = a0 > a1 ? a0 : a1; a1 = a1 > a2 ? a1 : a2; a2 = a2 > a3 ? a2 : a3; a3 = a3 > a4 ? a3 : a4;
            a4 = a4 > a5 ? a4 : a5; a5 = a5 > a6 ? a5 : a6; a6 = a6 > a7 ? a6 : a7; a7 = a7 > b ? a7 : b;
            b = b + 1.0;
        } else if (MODE == 3) {   // 64-bit integer compare + count
            c += (i0 > ib); c += (i1 > ib); c += (i2 > ib); c += (i3 > ib);
            c += (i0 + 5 > ib); c += (i1 + 5 > ib); c += (i2 + 5 > ib); c += (i3 + 5 > ib);
            ib += c;
        } else {                  // DADD
            a0 += a1; a1 += a2; a2 += a3; a3 += a4; a4 += a5; a5 += a6; a6 += a7; a7 += b;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + b + c + (double)(i0 + ib);
}

template <int MODE> void run(const char* name, double* d) {
    const int iters = 20000, blocks = 148 * 4, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 100, 1.5); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters, 1.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 8.0 * iters * blocks * threads;
    printf("%-28s %8.3f ms  %8.1f G thread-ops/s  = %6.1f ops/clk/SM at 1.965 GHz\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    double* d; cudaMalloc(&d, 148 * 4 * 256 * 8);
    run<0>("DFMA", d); run<4>("DADD", d); run<1>("DSETP.GT + IADD", d); run<2>("DSETP + select (max)", d); run<3>("ISETP 64-bit + IADD", d);
    return 0;
}
