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
struct GlobalSink {
    double* out; int N; int64_t w0, nwin;
    void put(int pos, int idx, double2 v) {
        int64_t w = w0 + pos;
        if (w >= nwin) return;
        double* o = out + w * N + 2 * (int64_t)idx;
        o[0] = v.x;
        o[1] = idx == 0 ? 0.0 : v.y;     // slot 0 carries the Nyquist bin in .y: dropped
    }
};
}  // namespace

extern "C" int emu_sliding(const double* series, int series_len, int N, int T, int S, int nthreads,
                           double* out) {
    Plan pl;
    if (!plan_make(pl, N, T, S)) return -1;
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
                fused_pass(t, nthreads, arena.data() + pl.off[i], pl.stride[i], pl.Q[i], 1 << (3 * (i - 1)),
                           pl.P[i - 1], 1, tw.data(), N, 3 * (i - 1), sink);
        }
        GlobalSink gs{out, N, w0, nwin};
        for (int t = 0; t < nthreads; t++)
            fused_pass(t, nthreads, arena.data() + pl.off[1], pl.stride[1], pl.Q[1], 1, pl.T, pl.S, tw.data(),
                       N, 0, gs);
    }
    return 0;
}
