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
    void bind(int kk) { k = kk; }
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
