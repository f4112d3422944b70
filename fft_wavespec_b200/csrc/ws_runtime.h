// ws_runtime.h — host runtime behind the C ABI: one Device per opened GPU (streams, coefficient
// tables, stream-ordered memory, pinned staging pool, one worker thread) and the job table of the
// imports.mqh submit / try_get / free calls (Include/imports.mqh:12-19).
//
// The reference's callers are single processes that call gpu_init once and then submit jobs
// (WaveSpecZZ_1.1.0-gpuopt.mq5:722-757, WaveCyclesBatchFetcher.mq5:105-133); SURVEY.md 8(e) shards
// series over GPUs with one host worker per device and no data-path collective.  So the unit here
// is the job: it is bound to one device at submit time (round robin over the open devices), its
// launches are issued by that device's worker thread, and its result leaves the device in window
// chunks that overlap the compute of the next chunk.
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wavespec_abi.h"
#include "ws_common.cuh"
#include "ws_series.h"

namespace wsrt {

// ---- errors ------------------------------------------------------------------------------------
extern thread_local std::string t_last_error;       // text behind gpu_get_last_error_w, per calling thread
extern std::atomic<int64_t> g_launches;             // kernels launched by this library
extern std::atomic<const char*> g_last_kernel;      // kernel family of the last FFT dispatch

int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define WS_CUDA(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ::wsrt::cuda_fail(e__, what); } while (0)

// Makes `dev` the calling thread's current device for the scope and restores the previous one.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) { cudaSetDevice(dev); changed = true; }
    }
    ~DeviceGuard() { if (changed && prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Stream-ordered device buffer (cudaMallocAsync on the device's pool, whose release threshold is
// unlimited: steady-state calls allocate nothing from the driver and never synchronise).  The free
// is ordered after the work already enqueued on `st`, so a buffer can be dropped while its kernels
// are still running.
struct AsyncBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t st = nullptr;
    AsyncBuf() = default;
    AsyncBuf(const AsyncBuf&) = delete;
    AsyncBuf& operator=(const AsyncBuf&) = delete;
    ~AsyncBuf() { release(); }
    cudaError_t alloc(size_t b, cudaStream_t s) {
        release();
        bytes = b; st = s;
        return b ? cudaMallocAsync(&p, b, s) : cudaSuccess;
    }
    void release() { if (p) { cudaFreeAsync(p, st); p = nullptr; } }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// Long-lived device buffer (tables, per-stream scratch)
struct DeviceBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DeviceBuf() = default;
    DeviceBuf(const DeviceBuf&) = delete;
    DeviceBuf& operator=(const DeviceBuf&) = delete;
    ~DeviceBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t b) {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = b;
        return b ? cudaMalloc(&p, b) : cudaSuccess;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// Pinned host staging, cached by size (cudaHostAlloc costs milliseconds per call).
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void*> free_list;
    size_t cached = 0;
    static constexpr size_t kMaxCached = (size_t)2 << 30;
    void* get(size_t bytes, size_t* got);
    void put(void* p, size_t bytes);
    void trim();
};

struct Job;

struct Device {
    int index = -1;
    std::mutex mu;                                   // tables and scratch maps
    std::vector<cudaStream_t> streams;               // compute streams handed out round robin
    std::vector<cudaStream_t> copy_streams;          // result copies (one per in-flight job, round robin)
    cudaStream_t side = nullptr;                     // Kalman4D beside the FFT kernels of the same call
    cudaStream_t h2d = nullptr;                      // series uploads of submitted jobs
    std::atomic<uint32_t> rr{0}, rr_copy{0};
    std::map<int, std::unique_ptr<DeviceBuf>> tw;                       // N -> exp(-2 pi i m/N)
    std::map<std::pair<int, int>, std::unique_ptr<DeviceBuf>> win;      // (N, type) -> w[i]
    std::map<std::pair<int, double>, std::unique_ptr<DeviceBuf>> apow;  // (N, alpha) -> alpha^j
    std::map<int, std::unique_ptr<DeviceBuf>> rowtab;                   // N -> per bin (N/k, N/(2 pi k))
    std::map<cudaStream_t, std::unique_ptr<DeviceBuf>> band_scratch;    // ws_sliding.cu -> ws_rows.cu hand-off
    std::map<cudaStream_t, std::unique_ptr<DeviceBuf>> phase_scratch;   // spectra of a window range (phase path)
    PinnedPool pinned;
    // worker: issues the launches of the jobs bound to this device, in submit order
    std::thread worker;
    std::mutex qmu;
    std::condition_variable qcv;
    std::deque<std::shared_ptr<Job>> queue;
    bool stop = false;

    cudaStream_t pick_stream() { return streams[rr.fetch_add(1) % streams.size()]; }
    cudaStream_t pick_copy_stream() { return copy_streams[rr_copy.fetch_add(1) % copy_streams.size()]; }
    int get_twiddles(int N, const double2** out);
    int get_window(int N, int type, const double** out);
    int get_apow(int N, double alpha, const double** out);
    int get_rowtab(int N, const double2** out);
};

enum JobKind { kJobWindow = 0, kJobBatchRows = 1, kJobCacheRecord = 2 };
enum JobState { kQueued = 0, kLaunched = 1, kFailed = 2, kCancelled = 3 };

// One piece of a job's product, in doubles of the product buffer; `done` fires when the kernels
// that produce it have finished.
struct Chunk { int64_t off = 0, elems = 0; cudaEvent_t done = nullptr; };

struct Job {
    std::mutex mu;
    Device* dev = nullptr;
    int kind = kJobBatchRows;
    wavespec_pipeline_cfg cfg;
    wavespec_cache_params cache;
    int32_t series_len = 0;
    int64_t nwin = 0;
    cudaStream_t st = nullptr, cst = nullptr;
    AsyncBuf d_series, d_rows, d_record;
    cudaEvent_t h2d_done = nullptr;
    std::vector<Chunk> chunks;
    std::atomic<int> state{kQueued};
    int status = WAVESPEC_OK;
    std::string error;
    int64_t rows = 0;                 // rows of the rows product (nwin * top_k)
    int64_t product_elems = 0;        // doubles of the product the caller fetches
    const double* d_product = nullptr;
    // delivery
    double* armed_out = nullptr;      // caller buffer the result is being copied into
    int64_t armed_elems = 0;
    bool armed_pinned = false;
    size_t next_chunk = 0;            // pageable delivery: chunks already copied
    cudaEvent_t copied = nullptr;     // pinned delivery: fires when the last chunk has landed
    bool delivered = false;
    // single-window jobs: one pinned host block owned by the job — the rows, then the window — that
    // the kernels read and write directly (zero-copy)
    void* h_rows = nullptr;
    size_t h_rows_bytes = 0;
    double* h_series = nullptr;
    ~Job();
};

struct Runtime {
    std::mutex mu;
    std::vector<std::unique_ptr<Device>> devs;       // open devices, in the order they were opened
    std::map<int64_t, std::shared_ptr<Job>> jobs;
    int64_t next_job = 1;
    std::atomic<uint32_t> rr_dev{0};
};
extern Runtime& g_rt;

// device lookup: nullptr (and last error set) when no device is open
Device* primary_device();
Device* device_by_index(int index);
Device* device_of_pointer(const void* device_ptr);
Device* next_device_round_robin();
int open_device(int device_index, int stream_count);
void close_all_devices();

// ws_pipeline.cu
int validate_cfg(const wavespec_pipeline_cfg* c, int32_t series_len);
struct Planes {
    double* spectra = nullptr; double* rows = nullptr; int32_t* bins = nullptr; double* waves = nullptr;
    double* contrib = nullptr; double* kalman = nullptr; double* phase = nullptr; double* wkalman = nullptr;
    int32_t* trk_index = nullptr; double* trk_period = nullptr;
};
// The whole per-bar pipeline on device pointers, enqueued on `st`.  [w_begin, w_begin + w_count) is the
// window range (w_count < 0: every window); ranges serve the stateless planes only.
int run_pipeline(Device& dev, const double* d_series, int32_t n_series, int32_t series_len,
                 const wavespec_pipeline_cfg* c, const Planes& out, cudaStream_t st,
                 int64_t w_begin = 0, int64_t w_count = -1);

// ws_jobs.cu
int submit_job(const double* series, int32_t series_len, const wavespec_pipeline_cfg& c, int kind,
               const wavespec_cache_params* cache, int64_t* job_id);
int try_get_job(int64_t job_id, double* out, int64_t out_cap, int32_t out_stride, int kind,
                int64_t* out_len, int32_t* ready);
int free_job(int64_t job_id);
void worker_main(Device* dev);

}  // namespace wsrt
