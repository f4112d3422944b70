"""The top-8 selection of the sliding kernels' batched epilogue (fft_wavespec_b200/csrc/ws_epilogue.cuh)
is a sorting network: each of the 8 lanes of a window sorts its 8 band entries, then three bitonic
merges with the lanes at distance 1, 2, 4 keep the best eight.  This test reads the compare-exchange
sequences out of the CUDA source and checks, on the CPU, that they sort (zero-one principle) and
that the merged result is exactly the reference's insertion top-K (A7a,
Legacy/WaveSpecZZ_1.0.3-pla-kalman-fast-gpuopt-nodetrend.mq5:537-554) — including ties, which the
reference resolves towards the lower bin."""
import itertools
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.path.dirname(HERE), "fft_wavespec_b200", "csrc", "ws_epilogue.cuh")


def networks():
    text = open(SRC).read()
    body = text[text.index("#define WS_CE"):text.index("#undef WS_CE")]
    body = body[body.index("WS_CE(0, 1)"):]                       # skip the macro definition itself
    pairs = [(int(a), int(b)) for a, b in re.findall(r"WS_CE\((\d), (\d)\)", body)]
    assert len(pairs) == 19 + 12, "expected Batcher's 19-exchange sort and a 12-exchange bitonic merge"
    return pairs[:19], pairs[19:]


def better(a, b):                      # (power, position): power desc, position asc — ws_common.cuh::better
    return a[0] > b[0] or (a[0] == b[0] and a[1] < b[1])


def ce(v, i, j):
    if better(v[j], v[i]):
        v[i], v[j] = v[j], v[i]


def test_sort_network_sorts_every_zero_one_input():
    sort8, _ = networks()
    for bits in itertools.product([0, 1], repeat=8):
        v = [(b, 0) for b in bits]
        for i, j in sort8:
            ce(v, i, j)
        assert [x[0] for x in v] == sorted(bits, reverse=True)


def test_bitonic_merge_sorts_every_bitonic_zero_one_input():
    _, merge = networks()
    for up in range(9):
        for down in range(9 - up):
            bits = [0] * up + [1] * down + [0] * (8 - up - down)          # rises then falls
            for seq in (bits, [1 - b for b in bits]):
                v = [(b, 0) for b in seq]
                for i, j in merge:
                    ce(v, i, j)
                assert [x[0] for x in v] == sorted(seq, reverse=True)


def test_lane_merges_reproduce_the_insertion_rule_with_ties(oracle):
    sort8, merge = networks()
    rng = np.random.default_rng(5)
    for trial in range(300):
        band = int(rng.integers(1, 65))
        # coarse values: many exact ties, some zeros (flat market), some large
        vals = rng.choice([0.0, 1.0, 2.0, 3.5], size=band) if trial % 2 else np.round(rng.random(band), 1)
        lanes = []
        for l in range(8):
            v = [((vals[l + 8 * i] if l + 8 * i < band else -2.0), l + 8 * i) for i in range(8)]
            for i, j in sort8:
                ce(v, i, j)
            lanes.append(v)
        for d in (1, 2, 4):
            nxt = []
            for l in range(8):
                a, b = lanes[l], lanes[l ^ d]
                c = [b[7 - i] if better(b[7 - i], a[i]) else a[i] for i in range(8)]
                for i, j in merge:
                    ce(c, i, j)
                nxt.append(c)
            lanes = nxt
        assert all(lanes[l] == lanes[0] for l in range(8))            # every lane ends with the same list
        got = [(e if p >= 0 else -1) for p, e in lanes[0]]
        # the reference: insertion scan over a spectrum whose band is bins lo .. lo + band - 1
        n, lo = 256, 4
        hi = lo + band - 1
        spec = np.full(n // 2, 99.0)                                  # out-of-band bins must be ignored
        spec[lo:hi + 1] = vals
        tb, tp = oracle.topk_insertion(spec, n, n / (hi + 0.5), n / lo, 8)    # band = [ceil(n/maxP), floor(n/minP)]
        want = [(b - lo if b >= 0 else -1) for b in tb]
        assert got == want, (trial, band)


# ---- the 32-bit key fast path in front of the exact network ---------------------------------------
def key_networks():
    text = open(SRC).read()
    body = text[text.index("#define WS_KE"):text.index("#undef WS_KE")]
    body = body[body.index("WS_KE(0, 1)"):]
    return [(int(a), int(b)) for a, b in re.findall(r"WS_KE\((\d), (\d)\)", body)]


def test_key_path_uses_the_same_networks():
    sort8, merge = networks()
    pairs = key_networks()
    assert pairs[:19] == sort8 and pairs[19:] == merge


def f32_rz_bits(p):
    """bits of the double p rounded towards zero to float (p >= 0), as __double2float_rz gives them"""
    f = np.float32(p)
    if float(f) > p:
        f = np.nextafter(f, np.float32(0.0))
    return int(np.array([f], dtype=np.float32).view(np.uint32)[0])


def key_path(vals, K):
    """Python statement of the fast path: returns (selected band positions, needs_exact)"""
    sort8, merge = networks()
    band = len(vals)
    lanes, lost = [], [0] * 8
    for l in range(8):
        k = []
        for i in range(8):
            e = l + 8 * i
            k.append(((f32_rz_bits(vals[e]) & ~127) | (64 - e)) if e < band and vals[e] >= 0 else 0)
        for i, j in sort8:
            k[i], k[j] = max(k[i], k[j]), min(k[i], k[j])
        lanes.append(k)
    for d in (1, 2, 4):
        nxt = []
        for l in range(8):
            a, b = lanes[l], lanes[l ^ d]
            c = []
            for i in range(8):
                lost[l] = max(lost[l], min(a[i], b[7 - i]))
                c.append(max(a[i], b[7 - i]))
            for i, j in merge:
                c[i], c[j] = max(c[i], c[j]), min(c[i], c[j])
            nxt.append(c)
        lanes = nxt
    assert all(lanes[l] == lanes[0] for l in range(8))
    key, worst = lanes[0], max(lost)
    bad = K == 8 and worst != 0 and (worst >> 7) == (key[7] >> 7)
    for i in range(7):
        bad = bad or (i < K and key[i + 1] != 0 and (key[i] >> 7) == (key[i + 1] >> 7))
    return [(64 - (k & 127)) if k else -1 for k in key[:K]], bad


def test_key_path_is_exact_whenever_it_does_not_ask_for_the_exact_network(oracle):
    rng = np.random.default_rng(11)
    taken = fallbacks = 0
    for trial in range(600):
        band = int(rng.integers(1, 65))
        K = int(rng.integers(1, 9))
        kind = trial % 4
        if kind == 0:
            vals = rng.random(band) * 10.0 ** rng.integers(-12, 12)            # generic powers
        elif kind == 1:
            vals = rng.random(band)
            i, j = rng.integers(0, band, 2)
            vals[i] = vals[j] * (1.0 + 1e-9)                                   # closer than a float can tell
        elif kind == 2:
            vals = rng.choice([0.0, 1.0, 2.0, 3.5], size=band)                 # exact ties
        else:
            vals = np.abs(rng.standard_normal(band)) ** 8                      # heavy tail
        got, bad = key_path(vals, K)
        if bad:
            fallbacks += 1
            continue
        taken += 1
        n, lo = 256, 4
        hi = lo + band - 1
        spec = np.full(n // 2, 1e300)
        spec[lo:hi + 1] = vals
        tb, tp = oracle.topk_insertion(spec, n, n / (hi + 0.5), n / lo, K)
        want = [(b - lo if b >= 0 else -1) for b in tb]
        assert got == want, (trial, band, K)
    assert taken > 250 and fallbacks > 100        # both outcomes exercised
