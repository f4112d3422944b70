// emu_warpfft.cpp — CPU emulation of the warp-per-window FFT (ws_window_fft_warp.cu): runs the
// SAME host/device arithmetic (fft_wavespec_b200/csrc/ws_warpfft_core.cuh) lane by lane, phase by
// phase (a phase = the code between two __syncwarp), and records the shared-memory bank groups
// every quarter warp touches so the swizzle can be checked without a GPU.  Test infrastructure.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../fft_wavespec_b200/csrc/ws_warpfft_core.cuh"

using namespace ws_wf;

namespace {

template <int LN>
int run(const double* v, double* out) {
    typedef Geo<LN> G;
    std::vector<double2> tw(G::N + pass_twiddle_count(LN));
    for (int m = 0; m < G::N; m++) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)G::N;
        tw[m] = make_double2((double)cosl(a), (double)sinl(a));
    }
    fill_pass_twiddles(LN, tw.data(), tw.data() + G::N);
    std::vector<double2> Z(G::M, make_double2(NAN, NAN));
    auto load = [&](int m) { return make_double2(v[2 * m], v[2 * m + 1]); };
    for (int lane = 0; lane < 32; lane++)
        dif_pass<LN, G::radix(0), G::stride(0), true>(lane, load, Z.data(), tw.data());
    // later passes: emulate the __syncwarp by running each pass for all lanes before the next
    if constexpr (G::P > 1) for (int lane = 0; lane < 32; lane++) dif_pass<LN, G::radix(1), G::stride(1), false>(lane, 0, Z.data(), tw.data());
    if constexpr (G::P > 2) for (int lane = 0; lane < 32; lane++) dif_pass<LN, G::radix(2), G::stride(2), false>(lane, 0, Z.data(), tw.data());
    if constexpr (G::P > 3) for (int lane = 0; lane < 32; lane++) dif_pass<LN, G::radix(3), G::stride(3), false>(lane, 0, Z.data(), tw.data());
    if constexpr (G::P > 4) for (int lane = 0; lane < 32; lane++) dif_pass<LN, G::radix(4), G::stride(4), false>(lane, 0, Z.data(), tw.data());
    std::vector<int> hits(G::M, 0);
    for (int lane = 0; lane < 32; lane++)
        split_phase<LN>(lane, Z.data(), tw.data(), [&](int k, double2 x) {
            out[2 * k] = x.x; out[2 * k + 1] = x.y; hits[k]++;
        });
    for (int k = 0; k < G::M; k++) if (hits[k] != 1) return -2;
    // split_bin must reproduce the split bit for bit
    for (int k = 0; k < G::M; k++) {
        double2 x = split_bin<LN>(Z.data(), tw.data(), k);
        if (std::memcmp(&x.x, &out[2 * k], 8) || std::memcmp(&x.y, &out[2 * k + 1], 8)) return -3;
    }
    return 0;
}

// worst number of lanes of one quarter warp that share a 16-byte bank group, over every
// shared-memory access instruction of the transform
template <int LN, int R, int S>
int pass_conflicts() {
    typedef Geo<LN> G;
    constexpr int NB = G::M / R;
    int worst = 0;
    for (int b = 0; b < (NB + 31) / 32; b++)
        for (int r = 0; r < R; r++)
            for (int q = 0; q < 4; q++) {
                int cnt[8] = {0};
                for (int l = 8 * q; l < 8 * q + 8; l++) {
                    int j = l + 32 * b;
                    if (j >= NB) continue;
                    int i = j & (S - 1), base = (j - i) * R + i;
                    cnt[(swz(base) ^ swz(r * S)) & 7]++;
                }
                for (int c : cnt) worst = c > worst ? c : worst;
            }
    return worst;
}

template <int LN>
int conflicts(int* gather_worst) {
    typedef Geo<LN> G;
    int worst = 0;
    auto upd = [&](int c) { worst = c > worst ? c : worst; };
    upd(pass_conflicts<LN, G::radix(0), G::stride(0)>());
    if constexpr (G::P > 1) upd(pass_conflicts<LN, G::radix(1), G::stride(1)>());
    if constexpr (G::P > 2) upd(pass_conflicts<LN, G::radix(2), G::stride(2)>());
    if constexpr (G::P > 3) upd(pass_conflicts<LN, G::radix(3), G::stride(3)>());
    if constexpr (G::P > 4) upd(pass_conflicts<LN, G::radix(4), G::stride(4)>());
    int gw = 0;
    for (int it = 0; it < (G::M / 2 + 31) / 32; it++)
        for (int side = 0; side < 2; side++)
            for (int q = 0; q < 4; q++) {
                int cnt[8] = {0};
                for (int l = 8 * q; l < 8 * q + 8; l++) {
                    int k = l + 32 * it;
                    if (k >= G::M / 2) continue;
                    int x = side ? ((G::M - k) & (G::M - 1)) : k;
                    cnt[swz(G::rev(x)) & 7]++;
                }
                for (int c : cnt) gw = c > gw ? c : gw;
            }
    *gather_worst = gw;
    return worst;
}

// inverse transform as ws_inverse.cu runs it: unpack pairs (phase), forward passes (one phase each),
// digit-reversed read-out
template <int LN>
int run_inverse(const double* spec, const unsigned char* keep, double* out) {
    typedef Geo<LN> G;
    std::vector<double2> tw(G::N + pass_twiddle_count(LN));
    for (int m = 0; m < G::N; m++) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)G::N;
        tw[m] = make_double2((double)cosl(a), (double)sinl(a));
    }
    fill_pass_twiddles(LN, tw.data(), tw.data() + G::N);
    std::vector<double2> Z(G::M, make_double2(NAN, NAN));
    auto X = [&](int k) { return (keep && !keep[k]) ? make_double2(0.0, 0.0) : make_double2(spec[2 * k], spec[2 * k + 1]); };
    for (int lane = 0; lane < 32; lane++)
        for (int k = lane; k <= G::M / 2; k += 32) {
            const int km = (G::M - k) & (G::M - 1);
            double2 xa = X(k);
            const double2 xb = k == 0 ? make_double2(0.0, 0.0) : X(km);
            if (k == 0) xa.y = 0.0;
            inverse_unpack<LN>(k, xa, xb, tw.data(), Z.data());
        }
    for (int lane = 0; lane < 32; lane++) dif_pass<LN, G::radix(0), G::stride(0), false>(lane, 0, Z.data(), tw.data());
    if constexpr (G::P > 1) for (int lane = 0; lane < 32; lane++) dif_pass<LN, G::radix(1), G::stride(1), false>(lane, 0, Z.data(), tw.data());
    if constexpr (G::P > 2) for (int lane = 0; lane < 32; lane++) dif_pass<LN, G::radix(2), G::stride(2), false>(lane, 0, Z.data(), tw.data());
    if constexpr (G::P > 3) for (int lane = 0; lane < 32; lane++) dif_pass<LN, G::radix(3), G::stride(3), false>(lane, 0, Z.data(), tw.data());
    if constexpr (G::P > 4) for (int lane = 0; lane < 32; lane++) dif_pass<LN, G::radix(4), G::stride(4), false>(lane, 0, Z.data(), tw.data());
    for (int lane = 0; lane < 32; lane++)
        for (int m = lane; m < G::M; m += 32) {
            const double2 v = inverse_pair<LN>(Z.data(), m);
            out[2 * m] = v.x; out[2 * m + 1] = v.y;
        }
    return 0;
}

}  // namespace

extern "C" int emu_warp_ifft(const double* spec, int n, const unsigned char* keep, double* out) {
    switch (n) {
        case 4: return run_inverse<2>(spec, keep, out);
        case 8: return run_inverse<3>(spec, keep, out);
        case 16: return run_inverse<4>(spec, keep, out);
        case 32: return run_inverse<5>(spec, keep, out);
        case 64: return run_inverse<6>(spec, keep, out);
        case 128: return run_inverse<7>(spec, keep, out);
        case 256: return run_inverse<8>(spec, keep, out);
        case 512: return run_inverse<9>(spec, keep, out);
        case 1024: return run_inverse<10>(spec, keep, out);
        case 2048: return run_inverse<11>(spec, keep, out);
        case 4096: return run_inverse<12>(spec, keep, out);
        case 8192: return run_inverse<13>(spec, keep, out);
        default: return -1;
    }
}

namespace {
}  // namespace

extern "C" int emu_warpfft(const double* v, int N, double* out) {
    switch (N) {
        case 16: return run<4>(v, out);
        case 32: return run<5>(v, out);
        case 64: return run<6>(v, out);
        case 128: return run<7>(v, out);
        case 256: return run<8>(v, out);
        case 512: return run<9>(v, out);
        case 1024: return run<10>(v, out);
        case 2048: return run<11>(v, out);
        case 4096: return run<12>(v, out);
        case 8192: return run<13>(v, out);
    }
    return -1;
}

extern "C" int emu_warpfft_conflicts(int N, int* gather_worst) {
    switch (N) {
        case 256: return conflicts<8>(gather_worst);
        case 512: return conflicts<9>(gather_worst);
        case 1024: return conflicts<10>(gather_worst);
        case 2048: return conflicts<11>(gather_worst);
        case 4096: return conflicts<12>(gather_worst);
    }
    return -1;
}
