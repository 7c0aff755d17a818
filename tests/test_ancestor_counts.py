"""The method of the ancestors kernel (csrc/kernels.cuh: k_ancestors), restated in Python
integers and IEEE doubles and checked on the CPU against the oracle's search
(orc_resample_weight, EXACT mode).

Output slot i continues the first particle whose cumulative weight reaches
ceil(uu0 + i du).  The kernel turns that round: particle j owns the slots
[c(C_{j-1}), c(C_j)), c(C) = #{i : ceil(uu0 + i du) <= C}, and c(C) comes from the estimate
x = (C - uu0) / du: floor(x) + 1 whenever x is further than 2^-12 from an integer (the
certificate), the thresholds themselves otherwise.  Checked here: (1) whenever the certificate
holds, the estimate IS the exact count; (2) ranges -> first-slot entries and entries at the
multiples of 32 -> running maximum per 32 slots gives the oracle's ancestors, for every weight
pattern the GPU tests use and tile offsets as the kernel forms them."""
import math

import numpy as np
import pytest

TILE = 1024          # kTile: particles per weight tile (cumulative weights are per tile)
EPS = 2.0 ** -12     # kCountEps


def _fixed(w):
    return [int(np.rint(x * 2.0 ** 52)) for x in w]


def _threshold(uu0, du, i):
    return math.ceil(uu0 + float(i) * du)


def ancestors_by_counts(w, u, stats):
    n = len(w)
    f = _fixed(w)
    tot = sum(f)
    tot_d = float(tot)
    uu0, du, inv_du = tot_d * u / float(n), tot_d / float(n), float(n) / tot_d
    heads = [-1] * n
    prev = 0
    offset = 0
    for t0 in range(0, n, TILE):
        d0 = float(offset) - uu0
        q = 0
        for j in range(t0, min(t0 + TILE, n)):
            q += f[j]                                    # in-tile inclusive cumulative weight
            C = offset + q
            x = (float(q) + d0) * inv_du
            xc = min(max(x, -1.0), float(n))
            k = min(int(math.floor(xc)) + 1, n)
            fr = xc - math.floor(xc)
            exact = k
            while exact < n and _threshold(uu0, du, exact) <= C:
                exact += 1
            while exact > 0 and _threshold(uu0, du, exact - 1) > C:
                exact -= 1
            if EPS <= fr <= 1.0 - EPS:
                stats["certified"] += 1
                assert k == exact, (j, x, k, exact)       # the certificate's claim
            else:
                stats["exact"] += 1
                k = exact
            if j == n - 1:
                k = n                                     # the last particle takes what is left
            if k > prev:
                heads[prev] = j
                m = (prev | 31) + 1
                while m < k:
                    heads[m] = j
                    m += 32
            prev = k
        offset += q
    out = []
    for i in range(n):                                    # ancestor_of: running maximum per warp
        if i % 32 == 0:
            assert heads[i] >= 0, i
            cur = -1
        cur = max(cur, heads[i])
        out.append(cur)
    return np.array(out)


def _weights(kind, n, rng):
    if kind == "uniform":
        return np.ones(n)
    if kind == "onehot":
        w = np.zeros(n); w[(2 * n) // 3] = 1.0
        return w
    if kind == "onehot_last":
        w = np.zeros(n); w[n - 1] = 1.0
        return w
    if kind == "sparse":
        w = rng.uniform(size=n); w[rng.uniform(size=n) < 0.9] = 0.0; w[n - 1] = 1.0
        return w
    if kind == "heavy":
        w = np.exp(rng.normal(0, 10, n))
        return w / w.max()
    if kind == "tiny":
        w = np.full(n, 1e-300); w[0] = 1.0
        return w
    if kind == "band":
        w = 2.0 ** rng.uniform(-46, -40, n); w[n // 2] = 1.0
        return w
    lw = -0.5 * rng.normal(0, 3, n) ** 2
    return np.exp(lw - lw.max())


@pytest.mark.parametrize("kind", ["uniform", "onehot", "onehot_last", "sparse", "heavy", "tiny",
                                  "band", "real"])
def test_counts_give_the_oracles_ancestors(orc, kind):
    rng = np.random.default_rng(3)
    stats = {"certified": 0, "exact": 0}
    for n in (1, 2, 31, 32, 33, 1000, 1024, 1025, 5000):
        w = _weights(kind, n, rng)
        for u in (0.0, 0.37, rng.uniform(), 1.0 - 2.0 ** -53):
            got = ancestors_by_counts(w, u, stats)
            want = orc.resample_weight(w, u, orc.EXACT)
            assert np.array_equal(got, want), (kind, n, u)
    assert stats["certified"] > 0


def test_thresholds_on_a_cumulative_weight_take_the_exact_path(orc):
    """Equal weights and u = 0: every threshold sits exactly on a cumulative weight, the
    estimate is an integer, and the count must come from the thresholds themselves."""
    stats = {"certified": 0, "exact": 0}
    for n in (8, 64, 2048, 3000):
        w = np.ones(n)
        got = ancestors_by_counts(w, 0.0, stats)
        assert np.array_equal(got, orc.resample_weight(w, 0.0, orc.EXACT))
    assert stats["exact"] > 0 and stats["certified"] == 0
