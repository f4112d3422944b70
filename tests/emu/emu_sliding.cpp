// emu_sliding.cpp — CPU emulation of the shared-butterfly sliding FFT kernel: runs the SAME
// host/device arithmetic (fft_wavespec_b200/csrc/ws_sliding_core.cuh) thread by thread, pass by
// pass, exactly as ws_sliding.cu schedules it on the GPU.  Test infrastructure: lets the index
// math and the tile/halo logic be checked on the CPU box against the oracle.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../fft_wavespec_b200/csrc/ws_sliding_core.cuh"

using namespace ws_slide;

namespace {
template <int N, int TOP>
struct GlobalSink {
    double* out; int64_t w0, nwin;
    int k; int pos;
    void bind(int kk, int) { k = kk; }
    void begin(int p) { pos = p; }
    template <int J> void put(double2 v) {
        if (TOP == 3) store(pos, SlotOfs<J>::c * (N / 16) + SlotOfs<J>::sgn * k, v);
        else store(pos, SlotOfs4<J & 3>::c * (N / 8) + SlotOfs4<J & 3>::sgn * k, v);
    }
    void put0(int p, int i, double2 v) { if (i == 0) v.y = 0.0; store(p, i, v); }   // Nyquist dropped
    void store(int p, int i, double2 v) {
        int64_t w = w0 + p;
        if (w >= nwin) return;
        double* o = out + w * N + 2 * (int64_t)i;
        o[0] = v.x; o[1] = v.y;
    }
};

template <int N>
void top(const Plan& pl, const std::vector<double2>& arena, const std::vector<double2>& tw, int nthreads,
         double* out, int64_t w0, int64_t nwin) {
    if (pl.top == 3) {
        GlobalSink<N, 3> gs{out, w0, nwin, 0, 0};
        for (int t = 0; t < nthreads; t++)
            chain_pass<N>(t, nthreads, arena.data() + pl.off[1], pl.T, pl.S, tw.data(), gs);
    } else {
        GlobalSink<N, 2> gs{out, w0, nwin, 0, 0};
        for (int t = 0; t < nthreads; t++)
            chain_pass4<N>(t, nthreads, arena.data() + pl.off[1], pl.T, pl.S, tw.data(), gs);
    }
}
}  // namespace

extern "C" int emu_sliding_top(const double* series, int series_len, int N, int T, int S, int top_levels,
                               int nthreads, double* out);

extern "C" int emu_sliding(const double* series, int series_len, int N, int T, int S, int nthreads,
                           double* out) {
    return emu_sliding_top(series, series_len, N, T, S, 3, nthreads, out);
}

extern "C" int emu_sliding_top(const double* series, int series_len, int N, int T, int S, int top_levels,
                               int nthreads, double* out) {
    Plan pl;
    if (!plan_make(pl, N, T, S, top_levels)) return -1;
    const int64_t nwin = series_len - N + 1;
    if (nwin < 1) return -2;
    std::vector<double2> tw(N);
    for (int m = 0; m < N; m++) {
        long double a = -2.0L * 3.141592653589793238462643383279502884L * m / N;
        tw[m] = make_double2((double)cosl(a), (double)sinl(a));
    }
    std::vector<double> x(pl.x_len);
    std::vector<double2> arena(pl.arena_slots);
    for (int64_t w0 = 0; w0 < nwin; w0 += T) {
        for (int i = 0; i < pl.x_len; i++) x[i] = (w0 + i < series_len) ? series[w0 + i] : 0.0;
        for (auto& a : arena) a = make_double2(NAN, NAN);
        for (int t = 0; t < nthreads; t++) bottom_level(t, nthreads, x.data(), pl, tw.data(), arena.data());
        for (int i = pl.nst; i >= 2; i--) {
            SmemSink sink{arena.data() + pl.off[i - 1], pl.stride[i - 1]};
            for (int t = 0; t < nthreads; t++)
                direct_pass(t, nthreads, arena.data() + pl.off[i], pl.stride[i], pl.Q[i], 1 << pl.lev[i - 1],
                            pl.P[i - 1], tw.data(), N, pl.lev[i - 1], sink);
        }
        switch (N) {
            case 256: top<256>(pl, arena, tw, nthreads, out, w0, nwin); break;
            case 512: top<512>(pl, arena, tw, nthreads, out, w0, nwin); break;
            case 1024: top<1024>(pl, arena, tw, nthreads, out, w0, nwin); break;
            case 2048: top<2048>(pl, arena, tw, nthreads, out, w0, nwin); break;
            default: top<4096>(pl, arena, tw, nthreads, out, w0, nwin); break;
        }
    }
    return 0;
}


// ---- staged form (sliding_staged_kernel): 64 threads per segment, thread kk owns chain slot kk
// (kk = 0: none), threads kk < 8 drop the packed-slot bins of the window into the row; a row leaves
// as ONE contiguous block.  The emulation keeps one row per window (no ring wrap: threads run one
// after the other here) and counts the writers of every bin.
namespace {
template <int N>
struct EmuStageSink {
    static constexpr int Q = N / 16, N2 = N / 2;
    double2* rows; int* hits; const double2* special;
    int k, kk;
    double2* slot; int* hslot;
    void bind(int k_, int) { k = k_; }
    void begin(int m) { slot = rows + (size_t)m * N2; hslot = hits + (size_t)m * N2; }
    template <int J> void put(double2 v) {
        const int idx = SlotOfs<J>::c * Q + SlotOfs<J>::sgn * k;
        slot[idx] = v; hslot[idx]++;
    }
    void end(int m) {
        if (kk < 8) { rows[(size_t)m * N2 + kk * Q] = special[m * 8 + kk]; hits[(size_t)m * N2 + kk * Q]++; }
    }
};
template <int N>
struct EmuSpecialSink {
    static constexpr int Q = N / 16;
    double2* table;
    void put0(int m, int i, double2 v) { if (i == 0) v.y = 0.0; table[m * 8 + i / Q] = v; }
};

template <int N>
int staged_top(const Plan& pl, const std::vector<double2>& arena, const std::vector<double2>& tw, double* out,
               int64_t w0, int64_t nwin) {
    constexpr int N2 = N / 2;
    std::vector<double2> rows((size_t)pl.T * N2, make_double2(NAN, NAN)), special((size_t)pl.T * 8);
    std::vector<int> hits((size_t)pl.T * N2, 0);
    EmuSpecialSink<N> sp{special.data()};
    for (int t = 0; t < 256; t++) special_pass<N>(t, 256, arena.data() + pl.off[1], pl.T, tw.data(), sp);
    const int per = pl.T / pl.S;
    for (int tid = 0; tid < pl.S * 64; tid++) {
        const int sub = tid >> 6, kk = tid & 63;
        EmuStageSink<N> sink{rows.data(), hits.data(), special.data(), 0, kk, nullptr, nullptr};
        chain_single_stepwise<N>(kk != 0, kk != 0 ? kk : 1, sub * per, per, arena.data() + pl.off[1], tw.data(), sink,
                                 [](int) {});
    }
    for (int m = 0; m < pl.T; m++) {
        for (int b = 0; b < N2; b++) if (hits[(size_t)m * N2 + b] != 1) return -3;     // every bin exactly once
        if (w0 + m >= nwin) continue;
        std::memcpy(out + (w0 + m) * N, rows.data() + (size_t)m * N2, sizeof(double2) * N2);   // the bulk store
    }
    return 0;
}
}  // namespace

extern "C" int emu_sliding_staged(const double* series, int series_len, int N, int T, int S, double* out) {
    Plan pl;
    if (!plan_make(pl, N, T, S, 3)) return -1;
    const int64_t nwin = series_len - N + 1;
    if (nwin < 1) return -2;
    std::vector<double2> tw(N);
    for (int m = 0; m < N; m++) {
        long double a = -2.0L * 3.141592653589793238462643383279502884L * m / N;
        tw[m] = make_double2((double)cosl(a), (double)sinl(a));
    }
    std::vector<double> x(pl.x_len);
    std::vector<double2> arena(pl.arena_slots);
    for (int64_t w0 = 0; w0 < nwin; w0 += T) {
        for (int i = 0; i < pl.x_len; i++) x[i] = (w0 + i < series_len) ? series[w0 + i] : 0.0;
        for (auto& a : arena) a = make_double2(NAN, NAN);
        for (int t = 0; t < 256; t++) bottom_level(t, 256, x.data(), pl, tw.data(), arena.data());
        for (int i = pl.nst; i >= 2; i--) {
            SmemSink sink{arena.data() + pl.off[i - 1], pl.stride[i - 1]};
            for (int t = 0; t < 256; t++)
                direct_pass(t, 256, arena.data() + pl.off[i], pl.stride[i], pl.Q[i], 1 << pl.lev[i - 1],
                            pl.P[i - 1], tw.data(), N, pl.lev[i - 1], sink);
        }
        int rc;
        switch (N) {
            case 512: rc = staged_top<512>(pl, arena, tw, out, w0, nwin); break;
            case 1024: rc = staged_top<1024>(pl, arena, tw, out, w0, nwin); break;
            case 2048: rc = staged_top<2048>(pl, arena, tw, out, w0, nwin); break;
            default: return -4;
        }
        if (rc) return rc;
    }
    return 0;
}
