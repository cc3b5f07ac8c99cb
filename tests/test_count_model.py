"""CPU model of the region-query kernel's "W form" (cloops_b200/csrc/index.cu:count_kernel_w), checked
against the oracle's neighbour counts.  It restates, tile by tile, exactly the integer algebra the kernel
relies on -- padded u', W = (relative strip << bu) | u', one-compare window tests, guard words, the
staged strip table, the hashed 3-cell occupancy bitmap -- so the algebra is verified without a GPU.
(The CUDA kernel itself is compared with the oracle in the -m gpu tests.)"""
import numpy as np
import pytest

from oracle import spec

TILE, RMAX, G, SMAX, BMW = 256, 2048, 8, 512, 512


def _bits(v):
    return int(v).bit_length()


def build_index(X, Y, eps):
    """mirror of index_build(): padded u', packed (strip,u') word and vmod, sorted by (strip,u')"""
    u = X - Y
    v = X + Y
    ubase = (u.min() // eps) * eps - eps
    vbase = (v.min() // eps) * eps
    uspan = int(u.max() - ubase + eps)
    be, bu = _bits(eps - 1), max(1, _bits(uspan))
    up = u - ubase
    vp = v - vbase
    strip = vp // eps
    vm = vp - strip * eps
    ks = (strip << bu) | up
    order = np.argsort(ks, kind="stable")
    ns = int(strip.max()) + 1
    sstart = np.searchsorted(strip[order], np.arange(-1, ns + 2), side="left")  # entry k = strip k-1
    return ks[order], vm[order], order, sstart, be, bu


def lower_bound(W, lo, hi, t):
    while lo < hi:
        mid = (lo + hi) >> 1
        if W[mid] < t:
            lo = mid + 1
        else:
            hi = mid
    return lo


def upper_bound(W, lo, hi, t):
    while lo < hi:
        mid = (lo + hi) >> 1
        if W[mid] <= t:
            lo = mid + 1
        else:
            hi = mid
    return lo


def model_counts(X, Y, eps, cap, bitmap=True, tile=TILE):
    ks, vmod, order, sstart, be, bu = build_index(X, Y, eps)
    n = len(ks)
    strip_of = ks >> bu
    out = np.zeros(n, dtype=np.int64)
    one = 1 << bu
    M = BMW * 32
    stats = {"fallback": 0, "pruned": 0, "searched": 0}

    def h(w):
        return ((w >> be) + (w >> bu) * 1237) & 0xFFFFFFFF

    for t0 in range(0, n, tile):
        t1 = min(t0 + tile, n)
        sA, sB = int(strip_of[t0]), int(strip_of[t1 - 1])
        nse = sB - sA + 4
        r0, r1 = int(sstart[sA]), int(sstart[sB + 3])
        ln = r1 - r0
        if ln > RMAX or nse > SMAX or (nse << bu) > 0xFFFFFFFF:
            stats["fallback"] += 1
            xs, ys = X[order], Y[order]
            for i in range(t0, t1):
                d = np.abs(xs - xs[i]) + np.abs(ys - ys[i])
                out[i] = min(int((d <= eps).sum()), cap)
            continue
        base = (sA - 1) << bu
        W = np.empty(ln + 2 * G, dtype=np.int64)
        V = np.zeros(ln + 2 * G, dtype=np.int64)
        W[:G] = 0
        W[G + ln:] = 0xFFFFFFFF
        W[G:G + ln] = ks[r0:r1] - base
        assert W[G:G + ln].min() >= 0 and W[G:G + ln].max() < 0xFFFFFFFF
        V[G:G + ln] = vmod[r0:r1]
        S = sstart[sA:sA + nse] - r0 + G
        BM = np.zeros(M, dtype=bool)
        BM[h(W[G:G + ln]) & (M - 1)] = True
        for i in range(t0, t1):
            me = G + i - r0
            wp = int(W[me])
            lo, hi = wp - eps, wp + eps
            assert lo >= 0
            if cap <= G + 1:
                c = 1
                for k in range(1, cap):
                    c += int(W[me - k] >= lo) + int(W[me + k] <= hi)
            else:
                srel = wp >> bu
                c = upper_bound(W, me + 1, int(S[srel + 1]), hi) - lower_bound(W, int(S[srel]), me, lo)
            dirs = 3
            if c < cap and bitmap:
                dirs = 0
                for bit, wq in ((1, wp - one), (2, wp + one)):
                    h0 = (h(wq) - 1) & (M - 1)
                    if BM[h0] or BM[(h0 + 1) & (M - 1)] or BM[(h0 + 2) & (M - 1)]:
                        dirs |= bit
                stats["pruned"] += 2 - bin(dirs).count("1")
            if c < cap and dirs:
                vm = int(V[me])
                srel = wp >> bu
                for bit, a, b, tlo, prev in ((1, S[srel - 1], S[srel], wp - one - eps, True),
                                             (2, S[srel + 1], S[srel + 2], wp + one - eps, False)):
                    if not (dirs & bit) or c >= cap:
                        continue
                    stats["searched"] += 1
                    thi = tlo + 2 * eps
                    assert tlo >= 0 and thi < 0xFFFFFFFF
                    j = lower_bound(W, int(a), int(b), tlo)
                    for k in range(4):
                        c += int(W[j + k] <= thi and (V[j + k] >= vm if prev else V[j + k] <= vm))
                    if W[j + 3] <= thi:
                        jj = j + 4
                        while c < cap and W[jj] <= thi:
                            c += int(V[jj] >= vm if prev else V[jj] <= vm)
                            jj += 1
            out[i] = min(c, cap)
    res = np.empty(n, dtype=np.int64)
    res[order] = out
    return res, stats


def _cases():
    rng = np.random.default_rng(11)
    # sparse background + dense clumps + duplicates + negative coordinates
    for n, span, eps in ((1500, 200_000, 1000), (1200, 30_000, 500), (900, 5_000_000, 1000), (700, 4000, 7), (600, 2000, 1),
                         (3000, 60_000, 1024)):
        X = rng.integers(0, span, n)
        d = np.exp(rng.uniform(np.log(10), np.log(max(20, span // 2)), n)).astype(np.int64)
        Y = X + d
        k = n // 5
        cx, cy = rng.integers(0, span, 8), rng.integers(0, span, 8)
        pick = rng.integers(0, 8, k)
        X[:k] = cx[pick] + rng.normal(0, eps / 2, k).astype(np.int64)
        Y[:k] = cy[pick] + rng.normal(0, eps / 2, k).astype(np.int64)
        X[k:k + 20] = X[0]
        Y[k:k + 20] = Y[0]
        if eps == 500:
            X -= span // 2
            Y -= span
        yield X.astype(np.int64), Y.astype(np.int64), eps


@pytest.mark.parametrize("bitmap", [True, False])
def test_w_form_model_matches_oracle(bitmap):
    searched = pruned = 0
    for X, Y, eps in _cases():
        want = spec.neighbour_counts(X, Y, eps)
        for cap in (2, 5, 9, 10, 40, 1 << 30):
            for tile in (256, 64):
                got, st = model_counts(X, Y, eps, cap, bitmap=bitmap, tile=tile)
                assert np.array_equal(got, np.minimum(want, cap)), (eps, cap, tile, np.flatnonzero(got != np.minimum(want, cap))[:5])
                searched += st["searched"]
                pruned += st["pruned"]
    assert searched > 0
    if bitmap:
        assert pruned > 0
