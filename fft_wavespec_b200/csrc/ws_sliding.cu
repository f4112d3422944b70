// ws_sliding.cu — shared-butterfly sliding real FFT for the plain hop-1 path (sm_100a).
//
// This is the headline kernel: the per-bar spectrum of every window of a series with no
// detrend and no window function (WaveSpecZZ_1.1.0-gpuopt.mq5:1239-1241 and
// Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:515-533), followed by the
// band-limited top-K selection and row pack (:537-568 of the latter).
//
// One CTA owns a tile of T consecutive windows of one series:
//   1. the T + N - 1 samples of the tile are staged in shared memory — by bulk async copy
//      (cp.async.bulk + mbarrier) in the producer / consumer kernel, with 128-bit loads that are all
//      issued before the first store elsewhere — together with the N/4 twiddles every pass uses
//      (each sample is read from HBM once per tile; nothing after this step waits for memory);
//   2. the deepest decimation level is computed directly from the samples (2..16-point DFTs);
//   3. fused radix-8 passes (ws_sliding_core.cuh) build the level-3 vectors in shared memory,
//      sharing every sub-transform between the overlapping windows of the tile (items numbered
//      position-fastest over odd position strides: no bank conflicts, warp-uniform twiddles);
//   4. the top pass produces the N/2 complex bins of each window in registers and streams them
//      to HBM with 128-bit stores (a warp writes contiguous 512-byte runs); in-band bins are
//      also captured in shared memory;
//   5. the warps run the batched top-K epilogue (ws_epilogue.cuh: four windows per warp, eight
//      lanes each at K <= 8 — a sorting / merge network on 32-bit keys with the exact (double, int)
//      network behind it) on the captured band and stream the rows out; wide bands take one pass
//      with per-lane register lists, the selection-sort rule and K > 8 the one-warp-per-window form.  Alternatively (WAVESPEC_SPLIT=1, and
//      always for the tracker plane) the in-band bins go to a compact global buffer consumed by
//      ws_rows.cu / the tracker kernel.
//
// Roofline: the only mandatory HBM traffic is 8*hop bytes in + 8*N bytes out per spectrum; the
// arithmetic is ~N/2 packed butterflies per window (~4.6 kflop-instr at N = 1024), far below the
// FP64 pipe limit, so the kernel is bound by the spectrum store.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ws_common.cuh"
#include "ws_epilogue.cuh"
#include "ws_series.h"
#include "ws_sliding_core.cuh"

namespace ws {

using ws_slide::Plan;

constexpr int kSlideThreads = 256;
// N = 4096 has 255 chains per segment and room for one CTA per SM only (210 KB of shared memory):
// when rows are produced it runs two segments on 512 threads so that 16 warps are resident instead
// of 8 (rows only 92 -> 118, spectra + rows 71 -> 82 M spectra/s); the pure spectra writer is
// faster with 8 warps (105 vs 96).
__host__ __device__ constexpr int slide_threads(int n, bool captures) { return (n >= 4096 && captures) ? 512 : kSlideThreads; }

struct SlideLayout {
    int x_doubles;      // staged samples (even count)
    int arena_off;      // byte offsets inside dynamic shared memory
    int xb_off;         // band capture [T][band] complex
    int ov_off;         // epilogue overlay (CTA epilogue buffers, or pw + ord of the per-window path)
    int band;           // captured bins per window
    int epi_mode;        // 2: batched warp epilogue (insertion rule); 0: warp-per-window path (sort rule)
    int Lg;             // lanes per window of the batched warp epilogue
    int overlap;        // 1: producer/consumer kernel (selection runs beside the top pass)
    int stage_off;      // overlap: consumer warps' row staging (outside the live work area)
    int staged;         // 1: spectra rows collected in a shared-memory ring and drained by bulk async stores
    int ring_off;       // staged: S rings of `ring_slots` spectrum rows (N/2 double2 each)
    int ring_slots;
    int special_off;    // staged: [T][8] bins of the packed slot 0 (multiples of N/16)
    int tw_off;         // the N/4 twiddles every pass of the tile uses, staged beside the samples (0: read from global);
                        // the producer/consumer kernel keeps the mbarrier of its bulk staging right behind them
    int fastsel;        // overlap: 32-bit key selection 0 never, 1 always, 2 for a tile's last groups only
    int total_bytes;
};

// Receives the bins of the top pass: streams them to HBM and captures the in-band ones.
// The eight bins of slot k are +-k + C_J Q: two moving pointers plus compile-time offsets.
// CAP: 0 no capture, 1 capture in shared memory (fused epilogue), 2 capture to the global band
// buffer consumed by the rows kernel
// TOP: levels fused in the top pass (3: eight bins per slot, 2: four bins per slot)
template <int J, int TOP> struct SlotOf { static constexpr int c = ws_slide::SlotOfs<J>::c, sgn = ws_slide::SlotOfs<J>::sgn; };
template <int J> struct SlotOf<J, 2> { static constexpr int c = ws_slide::SlotOfs4<J & 3>::c, sgn = ws_slide::SlotOfs4<J & 3>::sgn; };

template <int N, bool SPEC, int CAP, int TOP>
struct TopSink {
    static constexpr bool SEL = CAP != 0;
    double2* g;             // spectra of the tile's first window
    double2* xb;            // band capture of the tile's first window (shared or global)
    int lo, hi, band;
    int nvalid;             // windows of this tile that exist
    // per-thread state
    int k;
    unsigned inband;
    double2 *gpP, *gpM, *xpP, *xpM;
    bool ok, first;
    static constexpr int Q = N >> (TOP + 1), N2 = N / 2;
    template <int J> __device__ __forceinline__ void mark() {
        const int idx = SlotOf<J, TOP>::c * Q + SlotOf<J, TOP>::sgn * k;
        if (idx >= lo && idx <= hi) inband |= 1u << J;
    }
    // The windows of a chain follow consecutively from m0.  With a band capture the two spectrum
    // pointers and the two capture pointers MOVE from window to window (478 -> 437 instructions per four
    // windows of the producer/consumer kernel, 1.89 -> 1.87 ms per 1.2 M windows); the pure spectra writer
    // rebuilds them from g + m * N2 — the moving form makes that instance spill (1.65 -> 1.73 ms).
    static constexpr bool MOVING = CAP != 0;
    __device__ __forceinline__ void bind(int kk, int m0) {
        k = kk;
        inband = 0;
        if (SEL) {
            mark<0>(); mark<1>(); mark<2>(); mark<3>();
            if (TOP == 3) { mark<4>(); mark<5>(); mark<6>(); mark<7>(); }
        }
        if (MOVING) {
            if (SPEC) { gpP = g + (int64_t)m0 * N2 + k; gpM = g + (int64_t)m0 * N2 - k; }
            if (SEL) { xpP = xb + (m0 * band - lo) + k; xpM = xb + (m0 * band - lo) - k; }
            first = true;
        }
    }
    __device__ __forceinline__ void begin(int m) {
        ok = m < nvalid;
        if (MOVING) {
            if (!first) {
                if (SPEC) { gpP += N2; gpM += N2; }
                if (SEL) { xpP += band; xpM += band; }
            }
            first = false;
        } else {
            if (SPEC) { gpP = g + m * N2 + k; gpM = g + m * N2 - k; }
            if (SEL) { xpP = xb + (m * band - lo) + k; xpM = xb + (m * band - lo) - k; }
        }
    }
    // The validity test costs a branch per store group (10 % of the chain's instructions).  Two ways of
    // removing it were measured — predicated stores, and a second chain instance without the test for
    // full tiles (398 instead of 485 instructions per four windows) — and both made the top pass of a
    // 40-window tile SLOWER (38.0k and 35.5k against 31.9k cycles): without the reconvergence points
    // ptxas schedules more values live across the stores and the chain spills under the 128-register
    // cap.  The chain is bound by its stores' register hand-over to the LSU, not by instruction count.
    template <int J> __device__ __forceinline__ void put(double2 v) {
        if (!ok) return;
        constexpr int c = SlotOf<J, TOP>::c * Q;
        constexpr bool plus = SlotOf<J, TOP>::sgn > 0;
        // streaming stores; default, write-through and L2-only policies measured the same or slower
        if (SPEC) __stcs((plus ? gpP : gpM) + c, v);
        if (SEL) if ((inband >> J) & 1u) (plus ? xpP : xpM)[c] = v;
    }
    __device__ __forceinline__ void put0(int m, int i, double2 v) {
        if (m >= nvalid) return;
        if (i == 0) v.y = 0.0;       // slot 0 carries the Nyquist bin in .y: the contract drops it
        if (SPEC) __stcs(g + m * N2 + i, v);
        if (SEL) if (i >= lo && i <= hi) xb[m * band + (i - lo)] = v;
    }
};

template <int N, bool SPEC, int CAP, int TOP>
__global__ void __launch_bounds__(slide_threads(N, CAP != 0), N >= 4096 ? 1 : (TOP == 3 ? 2 : 3))
sliding_shared_kernel(const Params p, const Plan pl, const SlideLayout lay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = slide_threads(N, CAP != 0);
    double* x = reinterpret_cast<double*>(smem_raw);
    double2* arena = reinterpret_cast<double2*>(smem_raw + lay.arena_off);
    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const int64_t w0 = p.win_offset + (int64_t)blockIdx.x * pl.T;
    const int64_t wend = p.win_offset + p.chunk_nwin;
    const double* src = p.series + (int64_t)s * p.series_stride;
    const int nvalid = (int)((wend - w0) < pl.T ? (wend - w0) : pl.T);

    // 1. stage the tile (+ halo); beyond the series the samples only feed windows that are never stored.
    // The N/4 twiddles the passes use come along when the layout has room for them (see
    // sliding_overlap_kernel): no pass waits for the memory system after this point.  Measured: the
    // pure spectra writer at N <= 1024 — two chain warps per CTA, bound by the store stream alone — is
    // 9-14 % SLOWER with the table and the compile-time level structure (1.80-1.90 against 1.65 ms per
    // 1.2 M windows at N = 1024), every other instance 5-15 % faster; it keeps the plain form.
    constexpr bool kLean = SPEC && CAP == 0 && N <= 1024;
    if (p.prefetch_tiles > 0)
        prefetch_future_tile(src, p.series_len, w0 + (int64_t)p.prefetch_tiles * pl.T + (N - 1), pl.T, tid);
    const double2* twu = p.tw;
    if constexpr (kLean) {
        const int64_t left = (int64_t)p.series_len - w0;
        stage_samples(src + w0, pl.x_len, left > pl.x_len ? pl.x_len : (int)left, x, tid, NT);
        __syncthreads();
        ws_slide::bottom_level(tid, NT, x, pl, twu, arena);
        __syncthreads();
        for (int i = pl.nst; i >= 2; i--) {
            ws_slide::SmemSink sink{arena + pl.off[i - 1], pl.stride[i - 1]};
            ws_slide::direct_pass(tid, NT, arena + pl.off[i], pl.stride[i], pl.Q[i], 1 << pl.lev[i - 1],
                                  pl.P[i - 1], twu, pl.N, pl.lev[i - 1], sink);
            __syncthreads();
        }
    } else {
        constexpr int TWN = (N / 4 + NT - 1) / NT;
        double2 twv[TWN];
        if (lay.tw_off) {
#pragma unroll
            for (int u = 0; u < TWN; u++) if (tid + u * NT < N / 4) twv[u] = p.tw[tid + u * NT];
        }
        const int64_t left = (int64_t)p.series_len - w0;
        stage_samples(src + w0, pl.x_len, left > pl.x_len ? pl.x_len : (int)left, x, tid, NT);
        if (lay.tw_off) {
            double2* tws = reinterpret_cast<double2*>(smem_raw + lay.tw_off);
#pragma unroll
            for (int u = 0; u < TWN; u++) if (tid + u * NT < N / 4) tws[tid + u * NT] = twv[u];
            twu = tws;
        }
        __syncthreads();
        // 2. deepest level straight from the samples
        constexpr int NST = ws_slide::LevelsOf<N, TOP>::nst;      // compile-time level structure
        ws_slide::bottom_level<NST, ws_slide::LevelsOf<N, TOP>::Lb>(tid, NT, x, pl, twu, arena);
        __syncthreads();
        // 3. chain-free radix-8 passes down to level 3
#pragma unroll
        for (int i = NST; i >= 2; i--) {
            ws_slide::SmemSink sink{arena + pl.off[i - 1], pl.stride[i - 1]};
            ws_slide::direct_pass(tid, NT, arena + pl.off[i], pl.stride[i], pl.Q[i], 1 << pl.lev[i - 1],
                                  pl.P[i - 1], twu, pl.N, pl.lev[i - 1], sink);
            __syncthreads();
        }
    }
    // 4. top pass: level 3 -> full spectra, streamed to HBM
    const bool want_sel = CAP == 1;
    TopSink<N, SPEC, CAP, TOP> top;
    top.g = SPEC ? reinterpret_cast<double2*>(p.spectra) + ((int64_t)s * p.spec_nwin + (w0 - p.spec_w0)) * (N / 2) : nullptr;
    top.xb = nullptr;
    if (CAP == 1 && lay.band > 0) top.xb = reinterpret_cast<double2*>(smem_raw + lay.xb_off);
    if (CAP == 2) top.xb = p.band_buf + ((int64_t)s * p.chunk_nwin + (w0 - p.win_offset)) * lay.band;
    top.lo = p.band_lo; top.hi = p.band_hi; top.band = lay.band;
    if (p.select == 1 && top.lo < 1) top.lo = 1;
    if (lay.band <= 0) { top.lo = 1; top.hi = 0; }      // empty band: capture nothing
    top.nvalid = nvalid;
    if (TOP == 3) ws_slide::chain_pass<N>(tid, NT, arena + pl.off[1], pl.T, pl.S, twu, top);
    else ws_slide::chain_pass4<N>(tid, NT, arena + pl.off[1], pl.T, pl.S, twu, top);
    if (!want_sel) return;
    __syncthreads();
    // 5. selection + rows
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t gw_tile = (int64_t)s * p.nwin + w0;
    if (lay.band <= 0) {
        for (int wl = warp; wl < nvalid; wl += NT / 32)    // empty band: every slot absent
            warp_select_emit(p, nullptr, nullptr, nullptr, gw_tile + wl);
        return;
    }
    const int band = lay.band, lo = top.lo;
    unsigned char* ov = smem_raw + lay.ov_off;
    if (lay.epi_mode == 2) {
        // insertion rule, several windows per warp (ws_epilogue.cuh)
        double* pwa = reinterpret_cast<double*>(ov);
        const int pws = lay.Lg == 8 ? 64 : band;
        if (lay.Lg != 8) {
            // run-time group width: powers in shared memory for the K scan rounds (the 8-lane
            // network of the common case reads the captured bins directly)
            for (int i = tid; i < nvalid * band; i += NT) { double2 v = top.xb[i]; pwa[i] = v.x * v.x + v.y * v.y; }
            __syncthreads();
        }
        const int wpb = 32 / lay.Lg;
        double* stage = reinterpret_cast<double*>(ov + ((pl.T * pws * 8 + 15) & ~15)) + warp * (wpb * p.K * 16);   // wpb windows x K rows x <= 16 doubles
        for (int b0 = warp * wpb; b0 < nvalid; b0 += (NT / 32) * wpb) {
            const int nb = (nvalid - b0) < wpb ? (nvalid - b0) : wpb;
            if (lay.Lg == 8)
                warp_select_emit_batch<8>(p, nullptr, top.xb + b0 * band, band, lo, 8, nb, gw_tile + b0, stage);
            else
                warp_select_emit_batch<0>(p, pwa + b0 * band, top.xb + b0 * band, band, lo, lay.Lg, nb, gw_tile + b0, stage);
        }
        return;
    }
    // selection-sort rule (A7b) or K > 8: one warp per window
    double* pw = reinterpret_cast<double*>(ov);
    for (int i = tid; i < nvalid * band; i += NT) { double2 v = top.xb[i]; pw[i] = v.x * v.x + v.y * v.y; }
    __syncthreads();
    int* ord = reinterpret_cast<int*>(ov + pl.T * band * 8) + warp * band;
    for (int wl = warp; wl < nvalid; wl += NT / 32) {
        warp_select_emit(p, pw + wl * band - lo, top.xb + wl * band - lo, ord, gw_tile + wl);
        __syncwarp();
    }
    (void)lane;
}

// ---- producer / consumer form of the fused kernel -------------------------------------------------
// Measured on B200: two warps per CTA already keep the HBM store stream of the top pass saturated,
// while the selection + row arithmetic is latency bound.  So when a tile's chains fit four warps
// (S (Q - 1) <= 128) the CTA splits after the lower passes: the first ceil(S (Q - 1) / 32) warps run
// the chains and stream the spectra; the other warps first emit the packed-slot bins, then pick up
// each group of four windows as soon as the chains have produced it (named barriers 1 + it:
// producers arrive, the consumer warps that own a group of iteration `it` wait) and run the top-K
// network and the row stores beside the producers.  Nothing is waited for by the producers, and the band capture needs no double
// buffering because a group's rows are only written before its barrier.
__device__ __forceinline__ void named_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int N, bool SPEC>
__global__ void __launch_bounds__(kSlideThreads, 2)
sliding_overlap_kernel(const Params p, const Plan pl, const SlideLayout lay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // phase stamps of the first CTAs (debug, WAVESPEC_TIMING_FILE): start, lower passes done, last
    // producer warp done, last consumer warp done
    long long* dbg = (p.dbg && blockIdx.y == 0 && blockIdx.x < 1024) ? p.dbg + 4 * blockIdx.x : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();
    double* x = reinterpret_cast<double*>(smem_raw);
    double2* arena = reinterpret_cast<double2*>(smem_raw + lay.arena_off);
    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const int64_t w0 = p.win_offset + (int64_t)blockIdx.x * pl.T;
    const int64_t wend = p.win_offset + p.chunk_nwin;
    const double* src = p.series + (int64_t)s * p.series_stride;
    const int nvalid = (int)((wend - w0) < pl.T ? (wend - w0) : pl.T);

    // Every twiddle the tile needs is one of the first N/4 table entries (all passes use W_N^j with
    // j < N/4): they are fetched together with the samples, so that the passes never wait for the
    // memory system again — a global load issued under the other CTAs' spectrum stores takes
    // thousands of cycles, and there were five of them in a row between here and the top pass.
    double2* tws = reinterpret_cast<double2*>(smem_raw + lay.tw_off);
    const double2* twu = tws;
    if (p.prefetch_tiles > 0)
        prefetch_future_tile(src, p.series_len, w0 + (int64_t)p.prefetch_tiles * pl.T + (N - 1), pl.T, tid);
    // The tile and the twiddles arrive through the bulk-copy engine (cp.async.bulk + mbarrier, SASS
    // UBLKCP): one issuing thread instead of 256 threads' LDG + STS.  (Measured equal in time to plain
    // 128-bit loads that are all in flight before the first store — the ~4 000 cycles are an L2 round
    // trip under the chip-wide store stream, whatever path asks for it.)
    // The engine needs 16-byte aligned addresses and sizes: the sample array is shifted by one slot when
    // the tile starts on an odd sample, the copy may take one sample past the tile when the series has
    // it, and whatever is left at either end (odd first sample, end of the series) moves by plain loads.
    {
        const double* g = src + w0;
        const int64_t left = (int64_t)p.series_len - w0;
        const int head = (int)((reinterpret_cast<unsigned long long>(g) >> 3) & 1ull);
        x += head;                                                   // x[head] is 16-byte aligned
        const int want = (pl.x_len - head + 1) & ~1;
        const int64_t avail = left - head;
        const int nb = avail >= want ? want : (avail > 0 ? (int)(avail & ~1LL) : 0);
        const unsigned bar = (unsigned)__cvta_generic_to_shared(smem_raw + lay.tw_off + (N / 4) * 16);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nb * 8 + (N / 4) * 16) : "memory");
            if (nb > 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"((unsigned)__cvta_generic_to_shared(x + head)), "l"(g + head), "r"(nb * 8), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((unsigned)__cvta_generic_to_shared(tws)), "l"(p.tw), "r"((N / 4) * 16), "r"(bar) : "memory");
        }
        if (head && tid == 32) x[0] = left > 0 ? g[0] : 0.0;
        for (int i = head + nb + tid; i < pl.x_len; i += kSlideThreads) x[i] = i < left ? g[i] : 0.0;
        __syncthreads();                                             // the barrier is initialised for everybody
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    constexpr int NST = ws_slide::LevelsOf<N>::nst;          // the plan's top is 3 here: compile-time level structure
    ws_slide::bottom_level<NST, ws_slide::LevelsOf<N>::Lb>(tid, kSlideThreads, x, pl, twu, arena);
    __syncthreads();
#pragma unroll
    for (int i = NST; i >= 2; i--) {
        ws_slide::SmemSink sink{arena + pl.off[i - 1], pl.stride[i - 1]};
        ws_slide::direct_pass(tid, kSlideThreads, arena + pl.off[i], pl.stride[i], pl.Q[i], 1 << pl.lev[i - 1],
                              pl.P[i - 1], twu, pl.N, pl.lev[i - 1], sink);
        __syncthreads();
    }

    if (dbg && tid == 0) dbg[1] = clock64();
    TopSink<N, SPEC, 1, 3> top;
    top.g = SPEC ? reinterpret_cast<double2*>(p.spectra) + ((int64_t)s * p.spec_nwin + (w0 - p.spec_w0)) * (N / 2) : nullptr;
    top.xb = reinterpret_cast<double2*>(smem_raw + lay.xb_off);
    top.lo = p.band_lo; top.hi = p.band_hi; top.band = lay.band;
    top.nvalid = nvalid;
    constexpr int gen = ws_slide::TopGeom<N>::Q - 1;
    const int per = pl.T / pl.S;                 // windows per chain, a multiple of 4
    const int iters = per >> 2;
    const int nprod = ((pl.S * gen + 31) >> 5) << 5;          // producer threads: whole warps
    const int bar_count = nprod + pl.S * 32;     // an iteration is awaited by one consumer warp per segment
    const double2* lvl3 = arena + pl.off[1];

    if (tid < nprod) {
        const bool active = tid < pl.S * gen;
        const int sub = active ? tid / gen : 0;
        const int k = active ? 1 + tid - sub * gen : 1;
        ws_slide::chain_single<N>(active, k, sub * per, per, lvl3, twu, top, [&](int it) {
            __threadfence_block();               // the group's captured bins before the arrival
            __syncwarp();
            named_arrive(1 + it, bar_count);
        });
        if (dbg && (tid & 31) == 0) atomicMax((unsigned long long*)&dbg[2], (unsigned long long)clock64());
        return;
    }

    // consumers: packed-slot bins of all windows first (they may lie inside the band)
    const int ncons = kSlideThreads - nprod;
    const int ctid = tid - nprod;
    const int cw = ctid >> 5, ncw = ncons >> 5;
    ws_slide::special_pass<N>(ctid, ncons, lvl3, pl.T, twu, top);
    named_sync(15, ncons);
    const int64_t gw_tile = (int64_t)s * p.nwin + w0;
    double* stage = reinterpret_cast<double*>(smem_raw + lay.stage_off) + cw * 512;
    // groups of four windows in completion order: b = it * S + segment; consumer warp cw takes
    // b = cw, cw + ncw, ...  (S <= ncw, so a warp waits on an iteration's barrier at most once)
    for (int b = cw; b < iters * pl.S; b += ncw) {
        const int it = b / pl.S, seg = b - it * pl.S;
        named_sync(1 + it, bar_count);
        const int b0 = seg * per + 4 * it;
        if (b0 < nvalid) {
            const int nb = (nvalid - b0) < 4 ? (nvalid - b0) : 4;
            const bool fast = lay.fastsel == 1 || (lay.fastsel == 2 && it == iters - 1);
            warp_select_emit_batch<8>(p, nullptr, top.xb + b0 * lay.band, lay.band, top.lo, 8, nb, gw_tile + b0, stage, fast);
        }
    }
    if (dbg && (tid & 31) == 0) atomicMax((unsigned long long*)&dbg[3], (unsigned long long)clock64());
}

// ---- staged form: spectra rows leave through the TMA engine -----------------------------------------
// Measured on this B200 (profiles/r02_stream_bw.json): a pure write stream reaches 6.3-6.9 TB/s, and
// ONE thread per SM issuing 8 KB bulk async stores from shared memory already reaches 6.4 TB/s, while
// the chains of the kernel above stop at 5.5 TB/s: their st.global keep the data registers busy until
// the memory system has taken the data (long-scoreboard stalls 5.4 warps per issue), so compute and
// store issue serialise inside each of the few chain warps.  Here the chains write their eight bins
// per window into a shared-memory row instead (st.shared never waits for HBM); when the 64 threads of
// a segment have completed a row, one of them hands it to the bulk-copy engine
// (cp.async.bulk.global.shared::cta, 8 KB per row) and the chains go on with the next window while the
// row drains.  A ring of RING rows per segment decouples the two; the only wait is the issuer's
// wait_group.read before a slot is rewritten.
//
// Thread map: 64 threads per segment, thread kk of a segment owns chain slot k = kk (kk = 0 has no
// chain: it issues the bulk stores); threads kk < 8 also drop the window's packed-slot bins (computed
// for the whole tile before the split) into the row.  Consumers as in sliding_overlap_kernel.
template <int N, int RING>
struct StageSink {
    static constexpr int Q = N >> 4, N2 = N / 2;
    double2* ring;              // this segment's ring
    double2* g;                 // spectra row of the tile's first window (global)
    double2* xb;                // band capture of the tile's first window (shared)
    const double2* special;     // [T][8]
    int lo, hi, band, nvalid, bar_id, kk;
    int k;
    unsigned inband;
    double2 *spP, *spM, *xpP, *xpM;
    template <int J> __device__ __forceinline__ void mark() {
        const int idx = ws_slide::SlotOfs<J>::c * Q + ws_slide::SlotOfs<J>::sgn * k;
        if (idx >= lo && idx <= hi) inband |= 1u << J;
    }
    __device__ __forceinline__ void bind(int k_, int) {
        k = k_;
        inband = 0;
        mark<0>(); mark<1>(); mark<2>(); mark<3>(); mark<4>(); mark<5>(); mark<6>(); mark<7>();
    }
    __device__ __forceinline__ void begin(int m) {
        double2* slot = ring + (m & (RING - 1)) * N2;
        spP = slot + k; spM = slot - k;
        xpP = xb + (m * band - lo) + k; xpM = xb + (m * band - lo) - k;
    }
    template <int J> __device__ __forceinline__ void put(double2 v) {
        constexpr int c = ws_slide::SlotOfs<J>::c * Q;
        constexpr bool plus = ws_slide::SlotOfs<J>::sgn > 0;
        (plus ? spP : spM)[c] = v;
        if ((inband >> J) & 1u) (plus ? xpP : xpM)[c] = v;
    }
    // after every window, all 64 threads of the segment
    __device__ __forceinline__ void end(int m) {
        double2* slot = ring + (m & (RING - 1)) * N2;
        if (kk < 8) slot[kk * Q] = special[m * 8 + kk];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // my row bytes -> visible to the copy engine
        // the slot the NEXT window writes was last read by row m + 1 - RING: at most RING - 2 of the
        // rows committed so far (.. m - 1) may still be draining when the segment passes the barrier
        if (kk == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(RING - 2) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        if (kk == 0 && m < nvalid) {
            const unsigned src = (unsigned)__cvta_generic_to_shared(slot);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(g + (int64_t)m * N2), "r"(src), "n"(N2 * 16) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
};

// fills the [T][8] table of packed-slot bins and captures the in-band ones
template <int N>
struct SpecialSink {
    static constexpr int Q = N >> 4;
    double2* table; double2* xb; int lo, hi, band;
    __device__ __forceinline__ void put0(int m, int i, double2 v) {
        if (i == 0) v.y = 0.0;                       // slot 0 carries the Nyquist bin in .y: the contract drops it
        table[m * 8 + i / Q] = v;
        if (i >= lo && i <= hi) xb[m * band + (i - lo)] = v;
    }
};

template <int N, int RING>
__global__ void __launch_bounds__(kSlideThreads, 2)
sliding_staged_kernel(const Params p, const Plan pl, const SlideLayout lay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* x = reinterpret_cast<double*>(smem_raw);
    double2* arena = reinterpret_cast<double2*>(smem_raw + lay.arena_off);
    const int tid = threadIdx.x;
    const int s = blockIdx.y;
    const int64_t w0 = p.win_offset + (int64_t)blockIdx.x * pl.T;
    const int64_t wend = p.win_offset + p.chunk_nwin;
    const double* src = p.series + (int64_t)s * p.series_stride;
    const int nvalid = (int)((wend - w0) < pl.T ? (wend - w0) : pl.T);

    {
        const int64_t left = (int64_t)p.series_len - w0;
        stage_samples(src + w0, pl.x_len, left > pl.x_len ? pl.x_len : (int)left, x, tid, kSlideThreads);
    }
    __syncthreads();
    ws_slide::bottom_level(tid, kSlideThreads, x, pl, p.tw, arena);
    __syncthreads();
    for (int i = pl.nst; i >= 2; i--) {
        ws_slide::SmemSink sink{arena + pl.off[i - 1], pl.stride[i - 1]};
        ws_slide::direct_pass(tid, kSlideThreads, arena + pl.off[i], pl.stride[i], pl.Q[i], 1 << pl.lev[i - 1],
                              pl.P[i - 1], p.tw, pl.N, pl.lev[i - 1], sink);
        __syncthreads();
    }
    const double2* lvl3 = arena + pl.off[1];
    double2* xb = reinterpret_cast<double2*>(smem_raw + lay.xb_off);
    double2* special = reinterpret_cast<double2*>(smem_raw + lay.special_off);
    {
        // packed-slot bins of every window of the tile (8 per window), one window per thread
        SpecialSink<N> sp{special, xb, p.band_lo, p.band_hi, lay.band};
        ws_slide::special_pass<N>(tid, kSlideThreads, lvl3, pl.T, p.tw, sp);
    }
    __syncthreads();

    const int per = pl.T / pl.S;                 // windows per chain, a multiple of 4
    const int iters = per >> 2;
    const int nprod = pl.S * 64;                 // 64 threads per segment
    const int bar_count = nprod + pl.S * 32;     // an iteration is awaited by one consumer warp per segment

    if (tid < nprod) {
        const int sub = tid >> 6, kk = tid & 63;
        StageSink<N, RING> top;
        top.ring = reinterpret_cast<double2*>(smem_raw + lay.ring_off) + (size_t)sub * RING * (N / 2);
        top.g = reinterpret_cast<double2*>(p.spectra) + ((int64_t)s * p.spec_nwin + (w0 - p.spec_w0)) * (N / 2);
        top.xb = xb; top.special = special;
        top.lo = p.band_lo; top.hi = p.band_hi; top.band = lay.band; top.nvalid = nvalid;
        top.bar_id = 14 - sub; top.kk = kk;
        ws_slide::chain_single_stepwise<N>(kk != 0, kk != 0 ? kk : 1, sub * per, per, lvl3, p.tw, top, [&](int it) {
            __threadfence_block();               // the group's captured bins before the arrival
            __syncwarp();
            named_arrive(1 + it, bar_count);
        });
        // the rows still draining read this CTA's shared memory
        if (kk == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        return;
    }

    const int ncons = kSlideThreads - nprod;
    const int ctid = tid - nprod;
    const int cw = ctid >> 5, ncw = ncons >> 5;
    const int64_t gw_tile = (int64_t)s * p.nwin + w0;
    double* stage = reinterpret_cast<double*>(smem_raw + lay.stage_off) + cw * 512;
    for (int b = cw; b < iters * pl.S; b += ncw) {
        const int it = b / pl.S, seg = b - it * pl.S;
        named_sync(1 + it, bar_count);
        const int b0 = seg * per + 4 * it;
        if (b0 < nvalid) {
            const int nb = (nvalid - b0) < 4 ? (nvalid - b0) : 4;
            warp_select_emit_batch<8>(p, nullptr, xb + b0 * lay.band, lay.band, p.band_lo, 8, nb, gw_tile + b0, stage);
        }
    }
}


// Staged (bulk-store) form: insertion rule with the 8-lane network, spectra + selection outputs.
// Tile: the spectrum ring (S x RING x 8 N bytes... N/2 double2 per row) has to fit beside the level-3
// array and the band capture with two CTAs per SM, so the tile is shorter than the direct-store one.
static int staged_ring_slots() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("WAVESPEC_RING"); v = (e && atoi(e) == 4) ? 4 : 2; }
    return v;
}
static bool try_staged_plan(const Params& p, Plan& pl, SlideLayout& lay) {
    static int want = -1;
    // measured slower than the direct-store form on B200 (2.68 vs 2.39 ms per 1.2 M windows,
    // profiles/README.md): opt-in
    if (want < 0) { const char* e = getenv("WAVESPEC_STAGED"); want = (e && e[0] == '1') ? 1 : 0; }
    const bool sel = (p.bins || p.rows || p.waves || p.contrib) && !p.band_buf;
    if (!want || !sel || !p.spectra || p.select != 0 || p.K > 8) return false;
    if (p.band_hi < p.band_lo || p.band_hi - p.band_lo + 1 > 64) return false;
    int T, S = 2;
    switch (p.N) {
        case 1024: T = 24; break;
        default: return false;
    }
    if (const char* ov = getenv("WAVESPEC_STAGED_TILE")) {       // tuning hook: "T,S"
        int t = 0, sc = 0;
        if (sscanf(ov, "%d,%d", &t, &sc) == 2 && t > 0 && sc > 0) { T = t; S = sc; }
    }
    if (S > 2 || !ws_slide::plan_make(pl, p.N, T, S, 3)) return false;
    const int per = T / S;
    if (per % 4 || per / 4 > 12) return false;                   // named barriers 1..12, 13 and 14 for the segments
    const int ring = staged_ring_slots();
    std::memset(&lay, 0, sizeof lay);
    lay.band = p.band_hi - p.band_lo + 1;
    lay.x_doubles = (pl.x_len + 1) & ~1;
    lay.arena_off = lay.x_doubles * 8;
    const int below3 = lay.arena_off + pl.off[1] * 16;           // dead once level 3 is complete
    const int work_end = lay.arena_off + pl.arena_slots * 16;
    const int ncw = (kSlideThreads - S * 64) / 32;
    const int stage_bytes = ncw * 512 * 8;
    const int special_bytes = T * 8 * 16;
    int tail = (work_end + 127) & ~127;
    int low = 0;
    auto place = [&](int bytes) {                                // in the dead region if it fits, else at the tail
        int off;
        if (low + bytes <= below3) { off = low; low = (low + bytes + 15) & ~15; }
        else { off = tail; tail = (tail + bytes + 127) & ~127; }
        return off;
    };
    lay.stage_off = place(stage_bytes);
    lay.special_off = place(special_bytes);
    lay.xb_off = tail; tail = (tail + T * lay.band * 16 + 127) & ~127;
    lay.ring_off = tail; tail += S * ring * (p.N / 2) * 16;
    lay.ring_slots = ring;
    lay.epi_mode = 2; lay.Lg = 8; lay.overlap = 1; lay.staged = 1;
    lay.total_bytes = (tail + 15) & ~15;
    return lay.total_bytes <= 113 * 1024;                        // two CTAs per SM
}


static bool pick_plan(const Params& p, Plan& pl, SlideLayout& lay) {
    if (try_staged_plan(p, pl, lay)) return true;
    std::memset(&lay, 0, sizeof lay);
    // Tile shapes measured on B200 (profiles/README.md).  With spectra AND selection outputs the
    // chains of a tile are kept on four warps (S (N/16 - 1) <= 128) so that the producer /
    // consumer kernel can run the selection beside them; a pure spectra writer is HBM bound with
    // any S (fewest chains win), a pure row producer wants all eight warps on the chains.
    const bool any_sel = (p.bins || p.rows || p.waves || p.contrib) && !p.band_buf;
    int T, S;
    switch (p.N) {
        case 256:  T = 128; S = (any_sel && p.spectra) ? 4 : 16; break;
        case 512:  T = 64;  S = (any_sel && p.spectra) ? 4 : 8; break;
        // spectra + rows: 40 windows per tile (two segments of 20) is what fits two CTAs per SM with
        // the consumers' row staging in the dead part of the work area; a longer tile spreads the lower
        // passes and the consumers' tail over more windows (32 -> 40: 2.28 -> 2.13 ms per 1.2 M windows)
        case 1024: T = (any_sel && p.spectra) ? 40 : 32;  S = any_sel ? (p.spectra ? 2 : 4) : 1; break;
        case 2048: T = 16;  S = 2;  break;
        case 4096: T = 16;  S = (any_sel || p.band_buf) ? 2 : 1; break;
        default: return false;
    }
    int top = 3;
    if (const char* ov = getenv("WAVESPEC_TILE")) {       // tuning hook: "T,S[,top]"
        int t = 0, sc = 0, tp = 3;
        int n = sscanf(ov, "%d,%d,%d", &t, &sc, &tp);
        if (n >= 2 && t > 0 && sc > 0) { T = t; S = sc; }
        if (n == 3) top = tp;
    }
    if (!ws_slide::plan_make(pl, p.N, T, S, top)) return false;
    int lo = p.band_lo, hi = p.band_hi;
    if (p.select == 1 && lo < 1) lo = 1;
    const bool sel = (p.bins || p.rows || p.waves || p.contrib) && !p.band_buf;
    lay.band = ((sel || p.band_buf) && hi >= lo) ? hi - lo + 1 : 0;
    lay.x_doubles = (pl.x_len + 4) & ~1;        // room for the 16-byte aligned bulk staging (one sample of shift, one of over-read)
    lay.arena_off = lay.x_doubles * 8;
    const int below3 = lay.arena_off + pl.off[1] * 16;               // bytes below the level-3 array
    const int work_end = lay.arena_off + pl.arena_slots * 16;
    const int xb_bytes = pl.T * lay.band * 16;
    // The band capture is written while level 3 is still being read: it may only reuse what lies
    // below level 3, otherwise it gets its own space after the work area.  The epilogue buffers
    // are used after the barrier that follows the top pass: they overlay the (dead) work area.
    const int warps = slide_threads(p.N, true) / 32;
    lay.epi_mode = 0;
    lay.Lg = 32;
    if (p.select == 0) {
        int need = p.K > (lay.band + 7) / 8 ? p.K : (lay.band + 7) / 8;   // >= K lanes, <= 8 bins per lane
        lay.Lg = 1;
        while (lay.Lg < need && lay.Lg < 32) lay.Lg <<= 1;
        lay.epi_mode = 2;
    }
    int ov_bytes;
    if (lay.epi_mode == 2) {
        ov_bytes = ((pl.T * (lay.Lg == 8 ? 64 : lay.band) * 8 + 15) & ~15) + warps * (32 / lay.Lg) * p.K * 16 * 8;
    } else {
        ov_bytes = pl.T * lay.band * 8 + warps * lay.band * 4;
    }
    ov_bytes = (ov_bytes + 15) & ~15;
    int end;
    if (xb_bytes <= below3) {                  // capture under level 3, overlay right after it
        lay.xb_off = 0;
        lay.ov_off = (xb_bytes + 15) & ~15;
        end = lay.ov_off + ov_bytes;
    } else if (ov_bytes <= work_end) {         // overlay on the work area, capture above it
        lay.ov_off = 0;
        lay.xb_off = work_end;
        end = work_end + xb_bytes;
    } else {                                   // overlay larger than the work area
        lay.ov_off = 0;
        lay.xb_off = ov_bytes;
        end = ov_bytes + xb_bytes;
    }
    if (!sel || lay.band == 0) end = 0;
    // producer / consumer form: insertion rule with the 8-lane network, chains on four warps, one
    // consumer warp per segment (S divides the four consumer warps), at most 14 iterations
    // (named barriers 1..14)
    lay.overlap = 0;
    lay.stage_off = 0;
    {
        const int gen = (p.N >> 4) - 1;
        const int per = pl.T / pl.S;
        static int want = -1;
        if (want < 0) { const char* e = getenv("WAVESPEC_OVERLAP"); want = (e && e[0] == '0') ? 0 : 1; }
        if (want && sel && lay.band > 0 && lay.epi_mode == 2 && lay.Lg == 8 && pl.top == 3 && p.N <= 2048 &&
            p.spectra && pl.S * gen <= 128 && pl.S <= (kSlideThreads - ((pl.S * gen + 31) & ~31)) / 32 &&
            per % 4 == 0 && per / 4 <= 14) {
            // The capture lives beside the work area (level 3 is read throughout).  The consumers' row
            // staging goes where the samples and the deeper levels were: they are dead once level 3 is
            // complete, which is before the CTA splits.
            const int xo = (work_end + 15) & ~15;
            const int stage_bytes = (kSlideThreads - ((pl.S * gen + 31) & ~31)) / 32 * 512 * 8;   // per consumer warp
            int so = 0, total = xo + xb_bytes;
            if (stage_bytes > below3) { so = (xo + xb_bytes + 15) & ~15; total = so + stage_bytes; }
            const int two = (total + 15) & ~15;          // twiddles: live from staging to the end of the top pass
            total = two + (p.N / 4) * 16 + 16;           // + the mbarrier of the bulk staging
            if (total <= 113 * 1024) {                   // keep two CTAs per SM
                lay.overlap = 1;
                // 32-bit key selection in the consumers: 0 never, 1 always, 2 for a tile's last groups only
                // (WAVESPEC_FASTSEL).  Measured, ms per 1.2 M windows at 1 965 MHz, never / always / last
                // groups: N = 512 1.25 / 1.12 / 1.21, N = 1024 1.87 / 1.97 / 1.87 — at N = 1024 the chains'
                // store stream is the limit there, and consumers that issue faster take issue slots from the
                // chains.  But under sustained load the boxes run into their 1 kW power cap (SM clock
                // 1 670 - 1 750 MHz), and there the 18 % fewer instructions win: bench.py 611 - 637 M spectra/s
                // without the keys depending on the box, 634 M on every box with them (654 / 628 M for one
                // unthrottled launch of the same shape).  Sustained load is the operating point: always on.
                { static int fs = -1; if (fs < 0) { const char* e = getenv("WAVESPEC_FASTSEL"); fs = e ? atoi(e) : -1; }
                  lay.fastsel = fs >= 0 ? fs : 1; }
                lay.xb_off = xo;
                lay.stage_off = so;
                lay.tw_off = two;
                end = total;
            }
        }
    }
    lay.total_bytes = end > work_end ? end : work_end;
    lay.total_bytes = (lay.total_bytes + 15) & ~15;
    if (!lay.overlap) {
        // twiddle table behind everything else, unless it would cost a resident CTA (or not fit at all)
        const int with_tw = lay.total_bytes + (p.N / 4) * 16;
        const int cap = lay.total_bytes <= 113 * 1024 ? 113 * 1024 : 232448;
        static int want_tw = -1;
        if (want_tw < 0) { const char* e = getenv("WAVESPEC_TWS"); want_tw = (e && e[0] == '0') ? 0 : 1; }
        const bool lean = p.spectra && !sel && !p.band_buf && p.N <= 1024;      // kLean instance of the kernel
        if (want_tw && !lean && lay.total_bytes > 0 && with_tw <= cap) { lay.tw_off = lay.total_bytes; lay.total_bytes = with_tw; }
    }
    return lay.total_bytes <= 232448;
}

bool sliding_shared_supported(const Params& p) {
    if (p.hop != 1 || p.detrend != 0 || p.has_window || p.feed || p.phase) return false;
    Plan pl; SlideLayout lay;
    return pick_plan(p, pl, lay);
}

template <int N, bool SPEC, int CAP, int TOP>
static cudaError_t launch_top(const Params& p, const Plan& pl, const SlideLayout& lay, cudaStream_t stream) {
    static std::atomic<unsigned long long> attr_seen{0};
    if (first_launch_on_device(attr_seen)) {
        cudaError_t e = cudaFuncSetAttribute(sliding_shared_kernel<N, SPEC, CAP, TOP>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)((p.chunk_nwin + pl.T - 1) / pl.T), (unsigned)p.n_series);
    sliding_shared_kernel<N, SPEC, CAP, TOP><<<grid, slide_threads(N, CAP != 0), lay.total_bytes, stream>>>(p, pl, lay);
    return cudaGetLastError();
}

template <int N, bool SPEC>
static cudaError_t launch_overlap(const Params& p, const Plan& pl, const SlideLayout& lay, cudaStream_t stream) {
    if (N > 2048) return cudaErrorInvalidValue;          // chains of N = 4096 need all eight warps
    static std::atomic<unsigned long long> attr_seen{0};
    if (first_launch_on_device(attr_seen)) {
        cudaError_t e = cudaFuncSetAttribute(sliding_overlap_kernel<N, SPEC>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)((p.chunk_nwin + pl.T - 1) / pl.T), (unsigned)p.n_series);
    sliding_overlap_kernel<N, SPEC><<<grid, kSlideThreads, lay.total_bytes, stream>>>(p, pl, lay);
    return cudaGetLastError();
}

template <int N>
static cudaError_t launch_staged(const Params& p, const Plan& pl, const SlideLayout& lay, cudaStream_t stream) {
    if constexpr (N != 1024) return cudaErrorInvalidValue;
    else {
    dim3 grid((unsigned)((p.chunk_nwin + pl.T - 1) / pl.T), (unsigned)p.n_series);
    if (lay.ring_slots == 4) {
        static std::atomic<unsigned long long> attr_seen{0};
        if (first_launch_on_device(attr_seen)) {
            cudaError_t e = cudaFuncSetAttribute(sliding_staged_kernel<N, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
            if (e != cudaSuccess) return e;
        }
        sliding_staged_kernel<N, 4><<<grid, kSlideThreads, lay.total_bytes, stream>>>(p, pl, lay);
    } else {
        static std::atomic<unsigned long long> attr_seen{0};
        if (first_launch_on_device(attr_seen)) {
            cudaError_t e = cudaFuncSetAttribute(sliding_staged_kernel<N, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
            if (e != cudaSuccess) return e;
        }
        sliding_staged_kernel<N, 2><<<grid, kSlideThreads, lay.total_bytes, stream>>>(p, pl, lay);
    }
    return cudaGetLastError();
    }
}

template <int N, bool SPEC, int CAP>
static cudaError_t launch_one(const Params& p, const Plan& pl, const SlideLayout& lay, cudaStream_t stream) {
    return pl.top == 3 ? launch_top<N, SPEC, CAP, 3>(p, pl, lay, stream) : launch_top<N, SPEC, CAP, 2>(p, pl, lay, stream);
}

template <int N>
static cudaError_t launch_n(const Params& p, const Plan& pl, const SlideLayout& lay, cudaStream_t stream) {
    const bool spec = p.spectra != nullptr;
    const bool sel = p.bins || p.rows || p.waves || p.contrib;
    if (p.band_buf) {                          // hand the band to the rows kernel
        if (spec) return launch_one<N, true, 2>(p, pl, lay, stream);
        return launch_one<N, false, 2>(p, pl, lay, stream);
    }
    if (sel && lay.staged) return launch_staged<N>(p, pl, lay, stream);
    if (sel && lay.overlap) return spec ? launch_overlap<N, true>(p, pl, lay, stream) : launch_overlap<N, false>(p, pl, lay, stream);
    if (spec && sel) return launch_one<N, true, 1>(p, pl, lay, stream);
    if (spec) return launch_one<N, true, 0>(p, pl, lay, stream);
    if (sel) return launch_one<N, false, 1>(p, pl, lay, stream);
    return cudaSuccess;
}

cudaError_t launch_sliding_shared(Params p, cudaStream_t stream, const char** which) {
    Plan pl; SlideLayout lay;
    if (!pick_plan(p, pl, lay)) return cudaErrorInvalidValue;
    if (which && lay.overlap && !p.band_buf) *which = lay.staged ? "sliding_staged" : "sliding_overlap";
    p.tile_windows = pl.T;
    {
        // tiles ahead whose new samples a CTA pulls into L2: about one wave of resident CTAs
        static int pf = -2;
        if (pf == -2) { const char* e = getenv("WAVESPEC_PREFETCH"); pf = e ? atoi(e) : 296; }
        p.prefetch_tiles = pf;
    }
    switch (p.N) {
        case 256: return launch_n<256>(p, pl, lay, stream);
        case 512: return launch_n<512>(p, pl, lay, stream);
        case 1024: return launch_n<1024>(p, pl, lay, stream);
        case 2048: return launch_n<2048>(p, pl, lay, stream);
        case 4096: return launch_n<4096>(p, pl, lay, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace ws
