// ws_jobs.cu — the asynchronous job table behind gpu_submit_extract_cycles[_batch] /
// gpu_try_get_cycles[_batch] / gpu_free_job (Include/imports.mqh:12-19) and the cycle-cache record
// job (wavespec_submit_cycle_cache_batch).
//
// Life of a job
//   submit (caller's thread): bind the job to a device (round robin), upload the series on that
//       device's upload stream and wait for the upload only — the caller may reuse its buffer as
//       soon as submit returns (WaveSpecZZ_1.1.0-gpuopt.mq5:1313-1339) — then hand the job to the
//       device's worker thread.  Single-window jobs skip the upload: their window and their rows live
//       in one pinned host block that the kernels address directly.
//   worker: allocates the product from the stream-ordered pool and enqueues the kernels in window
//       chunks, recording an event per chunk.
//   try_get (any thread): never waits for the device.  The first poll "arms" the caller's buffer.
//       If it is page-locked every chunk's copy is enqueued at once on the job's copy stream, each
//       behind its chunk event, straight into the caller's memory (no staging, the copy of chunk i
//       overlaps the kernels of chunk i+1); later polls only query one event.  If it is pageable,
//       each poll copies the chunks that have completed since the previous poll.
//   free: drops the job; device memory returns to the pool in stream order, so an in-flight job
//       can be freed without waiting for its kernels (WaveSpecZZ_1.1.0-gpuopt.mq5:705-714).
#include <cstdlib>
#include <cstring>

#include "ws_runtime.h"

namespace wsrt {

// windows per chunk of a batch job (WAVESPEC_JOB_CHUNK overrides it: test hook)
static int64_t chunk_windows() {
    const char* e = getenv("WAVESPEC_JOB_CHUNK");
    const long long x = e ? atoll(e) : 0;
    return (int64_t)(x > 0 ? x : 262144);
}

Job::~Job() {
    if (!dev) return;
    DeviceGuard guard(dev->index);
    if (copied) { cudaEventSynchronize(copied); cudaEventDestroy(copied); }   // nothing may land in a freed caller buffer
    if (h_rows) {
        // the pinned block of a single-window job is read and written by the kernels on `st`: wait for
        // them before it goes back to the pool
        if (state.load() == kLaunched && !chunks.empty() && chunks.back().done) cudaEventSynchronize(chunks.back().done);
        else cudaStreamSynchronize(st);
        dev->pinned.put(h_rows, h_rows_bytes);
    }
    for (auto& c : chunks) if (c.done) cudaEventDestroy(c.done);
    if (h2d_done) cudaEventDestroy(h2d_done);
    d_series.release(); d_rows.release(); d_record.release();
    cudaGetLastError();
}

static std::shared_ptr<Job> find_job(int64_t id) {
    std::lock_guard<std::mutex> lk(g_rt.mu);
    auto it = g_rt.jobs.find(id);
    return it == g_rt.jobs.end() ? nullptr : it->second;
}

int submit_job(const double* series, int32_t series_len, const wavespec_pipeline_cfg& c, int kind,
               const wavespec_cache_params* cache, int64_t* job_id) {
    if (!job_id) return fail(WAVESPEC_BAD_ARGS, "job_id is null");
    *job_id = 0;
    Device* dev = next_device_round_robin();
    if (!dev) return WAVESPEC_BACKEND_UNAVAILABLE;
    if (!series) return fail(WAVESPEC_BAD_ARGS, "series is null");
    int rc = validate_cfg(&c, series_len);
    if (rc) return rc;
    if (kind == kJobCacheRecord && (!cache || c.row_stride < 14))
        return fail(WAVESPEC_BAD_ARGS, "cycle-cache jobs need cache parameters and a row stride >= 14");
    DeviceGuard guard(dev->index);
    auto job = std::make_shared<Job>();
    job->dev = dev; job->kind = kind; job->cfg = c; job->series_len = series_len;
    if (cache) job->cache = *cache; else std::memset(&job->cache, 0, sizeof job->cache);
    job->nwin = 1 + (int64_t)(series_len - c.window_len) / c.hop;
    job->rows = job->nwin * c.top_k;
    job->product_elems = kind == kJobCacheRecord ? (int64_t)series_len * 20 : job->rows * c.row_stride;
    job->st = dev->pick_stream();
    job->cst = dev->pick_copy_stream();
    if (kind == kJobWindow) {
        // Single-window jobs (one per bar in the 1.1.0 live loop, up to InpAsyncDepth in flight) run
        // zero-copy: one pinned block holds the rows, then the window; the kernels address it directly
        // (cudaHostAlloc memory is mapped under UVA).  No device allocation, no copy commands, nothing
        // to wait for in submit — the caller's buffer is free as soon as the memcpy returns.
        const size_t rows_bytes = (((size_t)job->rows * c.row_stride * 8) + 255) & ~(size_t)255;
        size_t got = 0;
        job->h_rows = dev->pinned.get(rows_bytes + (size_t)series_len * 8, &got);
        if (!job->h_rows) return fail(WAVESPEC_NO_MEM, "pinned block of a window job");
        job->h_rows_bytes = got;
        job->h_series = reinterpret_cast<double*>(static_cast<char*>(job->h_rows) + rows_bytes);
        std::memcpy(job->h_series, series, (size_t)series_len * 8);
        int64_t id;
        {
            std::lock_guard<std::mutex> lk(g_rt.mu);
            id = g_rt.next_job++;
            g_rt.jobs[id] = job;
        }
        {
            std::lock_guard<std::mutex> lk(dev->qmu);
            dev->queue.push_back(job);
        }
        dev->qcv.notify_one();
        *job_id = id;
        return WAVESPEC_OK;
    }
    // upload: the device buffer is released (stream-ordered) on the job's compute stream, so it is
    // allocated there; the copy itself runs on the upload stream and never queues behind kernels
    WS_CUDA(job->d_series.alloc((size_t)series_len * 8, dev->h2d), "cudaMallocAsync(series)");
    job->d_series.st = job->st;
    WS_CUDA(cudaMemcpyAsync(job->d_series.p, series, (size_t)series_len * 8, cudaMemcpyHostToDevice, dev->h2d),
            "cudaMemcpyAsync(series)");
    WS_CUDA(cudaEventCreateWithFlags(&job->h2d_done, cudaEventDisableTiming), "cudaEventCreate");
    WS_CUDA(cudaEventRecord(job->h2d_done, dev->h2d), "cudaEventRecord(upload)");
    int64_t id;
    {
        std::lock_guard<std::mutex> lk(g_rt.mu);
        id = g_rt.next_job++;
        g_rt.jobs[id] = job;
    }
    {
        std::lock_guard<std::mutex> lk(dev->qmu);
        dev->queue.push_back(job);
    }
    dev->qcv.notify_one();
    // pageable sources have left the caller's buffer when cudaMemcpyAsync returns; page-locked ones
    // when the event fires
    WS_CUDA(cudaEventSynchronize(job->h2d_done), "cudaEventSynchronize(upload)");
    *job_id = id;
    return WAVESPEC_OK;
}

// worker side: everything a job needs from the device, enqueued in one go
static int launch_job(Job& job) {
    Device& dev = *job.dev;
    const wavespec_pipeline_cfg& c = job.cfg;
    cudaStream_t st = job.st;
    if (job.kind == kJobWindow) {
        Planes out;
        out.rows = static_cast<double*>(job.h_rows);
        int rc = run_pipeline(dev, job.h_series, 1, job.series_len, &c, out, st, 0, job.nwin);
        if (rc) return rc;
        Chunk ch;
        ch.off = 0; ch.elems = job.rows * c.row_stride;
        WS_CUDA(cudaEventCreateWithFlags(&ch.done, cudaEventDisableTiming), "cudaEventCreate");
        job.chunks.push_back(ch);
        WS_CUDA(cudaEventRecord(ch.done, st), "cudaEventRecord(chunk)");
        return WAVESPEC_OK;
    }
    WS_CUDA(cudaStreamWaitEvent(st, job.h2d_done, 0), "cudaStreamWaitEvent(upload)");
    WS_CUDA(job.d_rows.alloc((size_t)job.rows * c.row_stride * 8, st), "cudaMallocAsync(rows)");
    if (job.kind == kJobCacheRecord)
        WS_CUDA(job.d_record.alloc((size_t)job.series_len * 20 * 8, st), "cudaMallocAsync(cache record)");
    job.d_product = job.kind == kJobCacheRecord ? job.d_record.as<double>() : job.d_rows.as<double>();
    const int64_t nwin = job.nwin;
    const int64_t per = chunk_windows();
    Planes out;
    out.rows = job.d_rows.as<double>();
    for (int64_t wa = 0; wa < nwin; wa += per) {
        const int64_t cn = wa + per <= nwin ? per : nwin - wa;
        int rc = run_pipeline(dev, job.d_series.as<double>(), 1, job.series_len, &c, out, st, wa, cn);
        if (rc) return rc;
        Chunk ch;
        if (job.kind == kJobCacheRecord) {
            // bars [wa*hop, (wa+cn)*hop); the last chunk also owns the bars past the last window start
            const int64_t b0 = wa * c.hop;
            const int64_t b1 = (wa + cn >= nwin) ? job.series_len : (wa + cn) * c.hop;
            WS_CUDA(ws::launch_cycle_cache(job.d_rows.as<double>(), nwin, c.top_k, c.row_stride, c.window_len, c.hop,
                                           job.series_len, b0, b1 - b0, c.sample_rate_seconds, job.cache,
                                           job.d_record.as<double>(), st), "cycle_cache kernel");
            g_launches++;
            ch.off = b0 * 20; ch.elems = (b1 - b0) * 20;
        } else {
            ch.off = wa * c.top_k * c.row_stride; ch.elems = cn * c.top_k * c.row_stride;
        }
        WS_CUDA(cudaEventCreateWithFlags(&ch.done, cudaEventDisableTiming), "cudaEventCreate");
        job.chunks.push_back(ch);
        WS_CUDA(cudaEventRecord(ch.done, st), "cudaEventRecord(chunk)");
    }
    return WAVESPEC_OK;
}

void worker_main(Device* dev) {
    cudaSetDevice(dev->index);
    for (;;) {
        std::shared_ptr<Job> job;
        {
            std::unique_lock<std::mutex> lk(dev->qmu);
            dev->qcv.wait(lk, [&] { return dev->stop || !dev->queue.empty(); });
            if (dev->queue.empty()) return;        // stop requested and nothing left
            job = std::move(dev->queue.front());
            dev->queue.pop_front();
        }
        std::lock_guard<std::mutex> lk(job->mu);
        if (job->state.load() == kCancelled) continue;
        int rc = launch_job(*job);
        if (rc) {
            job->status = rc;
            job->error = t_last_error;
            job->state.store(kFailed);
        } else {
            job->state.store(kLaunched);
        }
    }
}

// Copies (or schedules the copies of) what the caller may have: `elems` doubles from the start of
// the product.  Returns 1 when everything has landed in `out`, 0 when not yet, < 0 on error.
static int deliver(Job& job, double* out, int64_t elems) {
    if (elems <= 0) return 1;
    if (job.delivered && job.armed_out == out && job.armed_elems == elems) return 1;
    if (job.armed_out != out || job.armed_elems != elems) {
        // (re)arm: a copy into a previous buffer must be over before we forget about it
        if (job.copied) { cudaEventSynchronize(job.copied); cudaEventDestroy(job.copied); job.copied = nullptr; }
        job.armed_out = out; job.armed_elems = elems; job.next_chunk = 0; job.delivered = false;
        cudaPointerAttributes a;
        job.armed_pinned = cudaPointerGetAttributes(&a, out) == cudaSuccess && a.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (job.armed_pinned) {
            for (const Chunk& ch : job.chunks) {
                if (ch.off >= elems) break;
                const int64_t n = ch.off + ch.elems <= elems ? ch.elems : elems - ch.off;
                WS_CUDA(cudaStreamWaitEvent(job.cst, ch.done, 0), "cudaStreamWaitEvent(chunk)");
                WS_CUDA(cudaMemcpyAsync(out + ch.off, job.d_product + ch.off, (size_t)n * 8, cudaMemcpyDeviceToHost,
                                        job.cst), "cudaMemcpyAsync(result chunk)");
            }
            WS_CUDA(cudaEventCreateWithFlags(&job.copied, cudaEventDisableTiming), "cudaEventCreate");
            WS_CUDA(cudaEventRecord(job.copied, job.cst), "cudaEventRecord(copied)");
        }
    }
    if (job.armed_pinned) {
        cudaError_t q = cudaEventQuery(job.copied);
        if (q == cudaErrorNotReady) return 0;
        if (q != cudaSuccess) return cuda_fail(q, "job failed on the device");
        job.delivered = true;
        return 1;
    }
    // pageable destination: take the chunks that are complete, in order
    while (job.next_chunk < job.chunks.size()) {
        const Chunk& ch = job.chunks[job.next_chunk];
        if (ch.off >= elems) { job.next_chunk = job.chunks.size(); break; }
        cudaError_t q = cudaEventQuery(ch.done);
        if (q == cudaErrorNotReady) return 0;
        if (q != cudaSuccess) return cuda_fail(q, "job failed on the device");
        const int64_t n = ch.off + ch.elems <= elems ? ch.elems : elems - ch.off;
        WS_CUDA(cudaMemcpy(out + ch.off, job.d_product + ch.off, (size_t)n * 8, cudaMemcpyDeviceToHost),
                "cudaMemcpy(result chunk)");
        job.next_chunk++;
    }
    job.delivered = true;
    return 1;
}

// out_cap: rows for window jobs, doubles for batch and cache-record jobs.  *out_len: rows (window and
// batch jobs) or bars (cache-record jobs).
int try_get_job(int64_t job_id, double* out, int64_t out_cap, int32_t out_stride, int kind,
                int64_t* out_len, int32_t* ready) {
    if (out_len) *out_len = 0;
    if (ready) *ready = 0;
    if (!out || !out_len || !ready) return fail(WAVESPEC_BAD_ARGS, "null output pointer");
    auto job = find_job(job_id);
    if (!job) return fail(WAVESPEC_BAD_ARGS, "unknown job id");
    if (job->kind != kind) return fail(WAVESPEC_BAD_ARGS, "job id belongs to another job kind");
    // "still running": the async single-window loop of 1.1.0 (:1342-1374) only accepts NOT_READY;
    // WaveCyclesBatchFetcher.mq5:127-132 sleeps only on OK with ready == 0, and the 1.1.0 batch loop
    // (:1029-1040) takes both — so window jobs answer NOT_READY, batch jobs OK / ready 0.
    const int running = kind == kJobWindow ? WAVESPEC_NOT_READY : WAVESPEC_OK;
    const int s = job->state.load();
    if (s == kQueued) return running;
    std::unique_lock<std::mutex> lk(job->mu, std::try_to_lock);
    if (!lk.owns_lock()) return running;            // the worker (or another poller) holds it right now
    if (job->state.load() == kFailed) return fail(job->status, job->error);
    DeviceGuard guard(job->dev->index);
    if (kind == kJobWindow) {
        cudaError_t q = cudaEventQuery(job->chunks.back().done);
        if (q == cudaErrorNotReady) return running;
        if (q != cudaSuccess) return cuda_fail(q, "job failed on the device");
        int64_t rows = job->rows < out_cap ? job->rows : out_cap;
        const int js = job->cfg.row_stride;
        const int m = out_stride < js ? out_stride : js;
        const double* h = static_cast<const double*>(job->h_rows);
        for (int64_t r = 0; r < rows; r++) {
            for (int i = 0; i < m; i++) out[r * out_stride + i] = h[r * js + i];
            for (int i = m; i < out_stride; i++) out[r * out_stride + i] = 0.0;
        }
        *out_len = rows;
        *ready = 1;
        return WAVESPEC_OK;
    }
    int64_t units, elems;
    if (kind == kJobCacheRecord) {
        units = out_cap / 20 < job->series_len ? out_cap / 20 : job->series_len;      // bars
        elems = units * 20;
    } else {
        const int64_t cap_rows = out_cap / job->cfg.row_stride;
        units = job->rows;
        if (units > cap_rows) units = cap_rows - cap_rows % job->cfg.top_k;           // whole windows only
        elems = units * job->cfg.row_stride;
    }
    const int d = deliver(*job, out, elems);
    if (d < 0) return d;
    if (d == 0) return running;
    *out_len = units;
    *ready = 1;
    return WAVESPEC_OK;
}

int free_job(int64_t job_id) {
    std::shared_ptr<Job> job;
    {
        std::lock_guard<std::mutex> lk(g_rt.mu);
        auto it = g_rt.jobs.find(job_id);
        if (it == g_rt.jobs.end()) return fail(WAVESPEC_BAD_ARGS, "unknown job id");
        job = it->second;
        g_rt.jobs.erase(it);
    }
    // a job the worker has not reached yet is skipped; one that is in flight keeps running, and its
    // memory goes back to the pool in stream order when the last reference drops
    int expect = kQueued;
    job->state.compare_exchange_strong(expect, kCancelled);
    return WAVESPEC_OK;
}

}  // namespace wsrt
