// Micro-benchmark: what a pure WRITE stream, a pure READ stream and a copy reach on this B200, to put
// a number next to MEASURED_PEAKS.json's copy figure (half read, half write) for the sliding kernel,
// which is a pure writer.  Variants: 16-byte streaming stores (what the kernel issues), default-policy
// stores, 32-byte stores, bulk async stores from shared memory (cp.async.bulk.global.shared::cta, the
// TMA engine), 16-byte streaming loads, copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bw stream_bw.cu && ./stream_bw [GiB]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void write_cs(double2* __restrict__ p, size_t n) {
    const double2 v = make_double2(1.0 + threadIdx.x, 2.0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, v);
}
__global__ void write_default(double2* __restrict__ p, size_t n) {
    const double2 v = make_double2(1.0 + threadIdx.x, 2.0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// 8 stores in flight per thread, strided like the chain of the sliding kernel (8 runs of 512 B per warp)
__global__ void write_cs_x8(double2* __restrict__ p, size_t n) {
    const double2 v = make_double2(1.0 + threadIdx.x, 2.0);
    const size_t per = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * per < n; i += 8 * per) {
#pragma unroll
        for (int j = 0; j < 8; j++) __stcs(p + i + j * per, v);
    }
    for (; i < n; i += per) __stcs(p + i, v);
}
__global__ void write_32B(double4* __restrict__ p, size_t n) {     // n in double4 units
    const double a = 1.0 + threadIdx.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double* q = reinterpret_cast<double*>(p + i);
        asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(q), "d"(a), "d"(a), "d"(a), "d"(a) : "memory");
    }
}
__global__ void read_cs(const double2* __restrict__ p, size_t n, double* sink) {
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = __ldcs(p + i);
        acc += v.x + v.y;
    }
    if (acc == 123.456) *sink = acc;
}
__global__ void copy_cs(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(b + i, __ldcs(a + i));
}

// bulk async store: every CTA owns a shared tile of TILE bytes and streams it to successive rows
template <int TILE, int DEPTH>
__global__ void write_bulk(unsigned char* __restrict__ p, size_t bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < TILE * DEPTH / 8; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = 1.0 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t tiles = bytes / TILE;
        int slot = 0;
        for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            const unsigned src = (unsigned)__cvta_generic_to_shared(sm + slot * TILE);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + t * TILE), "r"(src), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
            slot = (slot + 1) % DEPTH;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <class F> static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 8.0;
    const size_t bytes = (size_t)(gib * (1ull << 30)) & ~(size_t)0xffff;
    unsigned char *a, *b; double* sink;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"buffer_gib\": %.2f, \"results\": [\n", pr.name, sms, gib);
    const size_t n16 = bytes / 16, n32 = bytes / 32;
    bool first = true;
    auto report = [&](const char* name, int ctas_per_sm, int threads, float ms, double moved) {
        printf("%s {\"kernel\": \"%s\", \"ctas_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"gb_per_s\": %.1f}", first ? "" : ",\n", name,
               ctas_per_sm, threads, ms, moved / ms / 1e6);
        first = false;
    };
    for (int cps : {2, 4, 8, 16}) {
        const int grid = sms * cps;
        report("write_16B_cs", cps, 256, time_ms([&] { write_cs<<<grid, 256>>>((double2*)a, n16); }, 5), (double)bytes);
        report("write_16B_cs_x8", cps, 256, time_ms([&] { write_cs_x8<<<grid, 256>>>((double2*)a, n16); }, 5), (double)bytes);
        report("write_16B_default", cps, 256, time_ms([&] { write_default<<<grid, 256>>>((double2*)a, n16); }, 5), (double)bytes);
        report("write_32B_cs", cps, 256, time_ms([&] { write_32B<<<grid, 256>>>((double4*)a, n32); }, 5), (double)bytes);
        report("read_16B_cs", cps, 256, time_ms([&] { read_cs<<<grid, 256>>>((const double2*)a, n16, sink); }, 5), (double)bytes);
        report("copy_16B_cs", cps, 256, time_ms([&] { copy_cs<<<grid, 256>>>((const double2*)a, (double2*)b, n16); }, 5), 2.0 * bytes);
    }
    CK(cudaFuncSetAttribute(write_bulk<8192, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 4));
    CK(cudaFuncSetAttribute(write_bulk<32768, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 2));
    for (int cps : {1, 2, 4}) {
        const int grid = sms * cps;
        report("write_bulk_8KB_x4", cps, 64, time_ms([&] { write_bulk<8192, 4><<<grid, 64, 8192 * 4>>>(a, bytes); }, 5), (double)bytes);
        report("write_bulk_32KB_x2", cps, 64, time_ms([&] { write_bulk<32768, 2><<<grid, 64, 32768 * 2>>>(a, bytes); }, 5), (double)bytes);
    }
    printf("\n]}\n");
    return 0;
}
