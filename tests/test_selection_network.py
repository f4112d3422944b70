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
