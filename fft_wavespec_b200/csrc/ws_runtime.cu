// ws_runtime.cu — devices, coefficient tables and pools of the host runtime (see ws_runtime.h).
#include "ws_runtime.h"
#include "ws_warpfft_core.cuh"

#include <cmath>
#include <cstring>

namespace wsrt {

thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};
std::atomic<const char*> g_last_kernel{"none"};
// Never destroyed: at process exit the CUDA runtime may already be gone when static destructors
// run, and the worker threads are still parked on their queues.  gpu_shutdown is the orderly way
// down; a process that exits without it simply drops everything.
Runtime& g_rt = *new Runtime;

int fail(int code, const std::string& msg) { t_last_error = msg; return code; }

int cuda_fail(cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    int code = (e == cudaErrorMemoryAllocation) ? WAVESPEC_NO_MEM
             : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
                   ? WAVESPEC_BACKEND_UNAVAILABLE : WAVESPEC_INTERNAL_ERROR;
    cudaGetLastError();            // the sticky-free errors must not leak into the next call's check
    return fail(code, m);
}

// ---- pinned staging ----------------------------------------------------------------------------
void* PinnedPool::get(size_t bytes, size_t* got) {
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = free_list.lower_bound(bytes);
        if (it != free_list.end() && it->first <= 2 * bytes + 4096) {
            void* p = it->second;
            *got = it->first;
            cached -= it->first;
            free_list.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    size_t want = bytes < 4096 ? 4096 : bytes;
    if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *got = want;
    return p;
}

void PinnedPool::put(void* p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu);
    if (cached + bytes > kMaxCached) { cudaFreeHost(p); return; }
    free_list.emplace(bytes, p);
    cached += bytes;
}

void PinnedPool::trim() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& kv : free_list) cudaFreeHost(kv.second);
    free_list.clear();
    cached = 0;
}

// ---- coefficient tables ------------------------------------------------------------------------
static const double kPi = 3.14159265358979323846;   // MQL5 M_PI

// Window coefficients with the reference's own expressions and evaluation order
// (Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast.mq5:1126-1156; type 5: Legacy/WaveSpecZZ_gpu_wip.mq5:954).
// Evaluated once per (N, type) on the host in IEEE double, i.e. the same values the MQL5 loop
// recomputes for every bar.
static void build_window(int n, int type, std::vector<double>& w) {
    w.resize(n);
    for (int i = 0; i < n; i++) {
        double v = 1.0;
        switch (type) {
            case WAVESPEC_WINDOW_HANN:     v = 0.5 * (1.0 - std::cos(2.0 * kPi * i / (n - 1))); break;
            case WAVESPEC_WINDOW_HAMMING:  v = 0.54 - 0.46 * std::cos(2.0 * kPi * i / (n - 1)); break;
            case WAVESPEC_WINDOW_BLACKMAN: v = 0.42 - 0.5 * std::cos(2.0 * kPi * i / (n - 1))
                                               + 0.08 * std::cos(4.0 * kPi * i / (n - 1)); break;
            case WAVESPEC_WINDOW_BARTLETT: v = 1.0 - std::fabs((2.0 * i - n + 1) / (n - 1)); break;
            case WAVESPEC_WINDOW_HANN_WIP: v = 0.5 - 0.5 * std::cos((2.0 * kPi * i) / (double)(n - 1)); break;
            default: break;
        }
        w[i] = v;
    }
}

static int upload_table(const std::vector<double>& h, std::unique_ptr<DeviceBuf>& buf, const char* what) {
    buf = std::make_unique<DeviceBuf>();
    WS_CUDA(buf->alloc(h.size() * 8), what);
    WS_CUDA(cudaMemcpy(buf->p, h.data(), h.size() * 8, cudaMemcpyHostToDevice), what);
    return WAVESPEC_OK;
}

int Device::get_twiddles(int N, const double2** out) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = tw.find(N);
    if (it == tw.end()) {
        std::vector<double> h(2 * (size_t)N);
        for (int m = 0; m < N; m++) {
            // exact table twiddles (long double evaluation, rounded once)
            long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)m / (long double)N;
            h[2 * m] = (double)cosl(a);
            h[2 * m + 1] = (double)sinl(a);
        }
        // exact values on the axes
        h[0] = 1.0; h[1] = 0.0;
        if (N >= 2) { h[2 * (N / 2)] = -1.0; h[2 * (N / 2) + 1] = 0.0; }
        if (N >= 4) { h[2 * (N / 4)] = 0.0; h[2 * (N / 4) + 1] = -1.0; h[2 * (3 * N / 4)] = 0.0; h[2 * (3 * N / 4) + 1] = 1.0; }
        // behind the N entries: the per-pass blocks of the warp-per-window transforms (ws_warpfft_core.cuh)
        int ln = 0;
        while ((1 << ln) < N) ln++;
        if ((1 << ln) == N && ln >= 4) {
            const int extra = ws_wf::pass_twiddle_count(ln);
            h.resize(2 * (size_t)(N + extra));
            ws_wf::fill_pass_twiddles(ln, reinterpret_cast<const double2*>(h.data()),
                                      reinterpret_cast<double2*>(h.data()) + N);
        }
        std::unique_ptr<DeviceBuf> buf;
        int rc = upload_table(h, buf, "twiddle table");
        if (rc) return rc;
        it = tw.emplace(N, std::move(buf)).first;
    }
    *out = it->second->as<double2>();
    return WAVESPEC_OK;
}

int Device::get_window(int N, int type, const double** out) {
    *out = nullptr;
    if (type == WAVESPEC_WINDOW_NONE) return WAVESPEC_OK;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(N, type);
    auto it = win.find(key);
    if (it == win.end()) {
        std::vector<double> h;
        build_window(N, type, h);
        std::unique_ptr<DeviceBuf> buf;
        int rc = upload_table(h, buf, "window table");
        if (rc) return rc;
        it = win.emplace(key, std::move(buf)).first;
    }
    *out = it->second->as<double>();
    return WAVESPEC_OK;
}

int Device::get_apow(int N, double alpha, const double** out) {
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(N, alpha);
    auto it = apow.find(key);
    if (it == apow.end()) {
        std::vector<double> h(N);
        for (int j = 0; j < N; j++) h[j] = std::pow(alpha, (double)j);
        std::unique_ptr<DeviceBuf> buf;
        int rc = upload_table(h, buf, "alpha^j table");
        if (rc) return rc;
        it = apow.emplace(key, std::move(buf)).first;
    }
    *out = it->second->as<double>();
    return WAVESPEC_OK;
}

// Per-bin constants of the result rows: .x = period N/k (the correctly rounded quotient the row
// carries), .y = N / (2 pi k) = 1 / (2 pi freq), the factor that turns a phase distance into bars.
int Device::get_rowtab(int N, const double2** out) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = rowtab.find(N);
    if (it == rowtab.end()) {
        std::vector<double> h((size_t)N, 0.0);          // N/2 bins x 2
        for (int k = 1; k < N / 2; k++) {
            h[2 * k] = (double)N / (double)k;
            h[2 * k + 1] = 1.0 / (2.0 * kPi * ((double)k / (double)N));
        }
        std::unique_ptr<DeviceBuf> buf;
        int rc = upload_table(h, buf, "row table");
        if (rc) return rc;
        it = rowtab.emplace(N, std::move(buf)).first;
    }
    *out = it->second->as<double2>();
    return WAVESPEC_OK;
}

// ---- devices -----------------------------------------------------------------------------------
Device* primary_device() {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    if (g_rt.devs.empty()) { fail(WAVESPEC_BACKEND_UNAVAILABLE, "gpu_init has not been called (or failed)"); return nullptr; }
    return g_rt.devs.front().get();
}

Device* device_by_index(int index) {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    for (auto& d : g_rt.devs) if (d->index == index) return d.get();
    return nullptr;
}

Device* device_of_pointer(const void* p) {
    cudaPointerAttributes a;
    if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) {
        if (Device* d = device_by_index(a.device)) return d;
        fail(WAVESPEC_BAD_ARGS, "the device that owns this pointer has not been opened with gpu_init");
        return nullptr;
    }
    cudaGetLastError();
    return primary_device();
}

Device* next_device_round_robin() {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    if (g_rt.devs.empty()) { fail(WAVESPEC_BACKEND_UNAVAILABLE, "gpu_init has not been called (or failed)"); return nullptr; }
    return g_rt.devs[g_rt.rr_dev.fetch_add(1) % g_rt.devs.size()].get();
}

static int open_one(int device_index, int stream_count) {
    // caller holds g_rt.mu
    for (auto& d : g_rt.devs) if (d->index == device_index) return WAVESPEC_OK;      // idempotent
    DeviceGuard guard(device_index);
    cudaDeviceProp prop;
    WS_CUDA(cudaGetDeviceProperties(&prop, device_index), "cudaGetDeviceProperties");
    if (prop.major < 10)
        return fail(WAVESPEC_BACKEND_UNAVAILABLE, "this library is built for sm_100a (B200) only");
    auto dev = std::make_unique<Device>();
    dev->index = device_index;
    const int n = stream_count < 1 ? 1 : (stream_count > 32 ? 32 : stream_count);   // more CUDA streams buy nothing
    dev->streams.resize(n);
    for (int i = 0; i < n; i++)
        WS_CUDA(cudaStreamCreateWithFlags(&dev->streams[i], cudaStreamNonBlocking), "cudaStreamCreate");
    dev->copy_streams.resize(n < 4 ? n : 4);
    for (auto& s : dev->copy_streams) WS_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate(copy)");
    WS_CUDA(cudaStreamCreateWithFlags(&dev->side, cudaStreamNonBlocking), "cudaStreamCreate(side)");
    WS_CUDA(cudaStreamCreateWithFlags(&dev->h2d, cudaStreamNonBlocking), "cudaStreamCreate(h2d)");
    // stream-ordered allocations stay cached in the device's pool between calls
    cudaMemPool_t pool;
    WS_CUDA(cudaDeviceGetDefaultMemPool(&pool, device_index), "cudaDeviceGetDefaultMemPool");
    uint64_t keep = UINT64_MAX;
    WS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep), "cudaMemPoolSetAttribute");
    Device* raw = dev.get();
    dev->worker = std::thread(worker_main, raw);
    g_rt.devs.push_back(std::move(dev));
    return WAVESPEC_OK;
}

int open_device(int device_index, int stream_count) {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(WAVESPEC_BACKEND_UNAVAILABLE,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if (device_index == -1) {                       // every device of the box (one session per device)
        for (int d = 0; d < count; d++) {
            int rc = open_one(d, stream_count);
            if (rc) return rc;
        }
        return WAVESPEC_OK;
    }
    if (device_index < 0 || device_index >= count) return fail(WAVESPEC_BAD_ARGS, "device_index out of range");
    return open_one(device_index, stream_count);
}

void close_all_devices() {
    std::vector<std::unique_ptr<Device>> devs;
    std::map<int64_t, std::shared_ptr<Job>> jobs;
    {
        std::lock_guard<std::mutex> lk(g_rt.mu);
        devs.swap(g_rt.devs);
        jobs.swap(g_rt.jobs);
    }
    for (auto& d : devs) {
        { std::lock_guard<std::mutex> lk(d->qmu); d->stop = true; }
        d->qcv.notify_all();
        if (d->worker.joinable()) d->worker.join();
    }
    for (auto& d : devs) {
        DeviceGuard guard(d->index);
        cudaDeviceSynchronize();
        for (auto it = jobs.begin(); it != jobs.end();)
            if (it->second->dev == d.get()) it = jobs.erase(it); else ++it;
        d->queue.clear();
        d->band_scratch.clear(); d->phase_scratch.clear();
        d->tw.clear(); d->win.clear(); d->apow.clear(); d->rowtab.clear();
        d->pinned.trim();
        for (auto s : d->streams) cudaStreamDestroy(s);
        for (auto s : d->copy_streams) cudaStreamDestroy(s);
        if (d->side) cudaStreamDestroy(d->side);
        if (d->h2d) cudaStreamDestroy(d->h2d);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d->index) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        cudaGetLastError();
    }
}

}  // namespace wsrt
