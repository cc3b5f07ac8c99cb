"""CPU model of the region-query kernel (cloops_b200/csrc/region_query.cu:count_kernel_tiled), checked against
the oracle's neighbour counts.  It restates, tile by tile, exactly the integer algebra the kernel relies
on -- padded u', W = (relative strip << bu) | u' with one-compare window tests, guard words instead of
bounds tests, slots that keep the global index modulo 4, the per-tile header, the uniform (fixed trip
count) binary search that may run past a strip's end, 4 probes + tail -- so that algebra is verified
without a GPU.  (The CUDA kernel itself is compared with the oracle in the -m gpu tests.)"""
import numpy as np
import pytest

from oracle import spec

G, GR, SMAX = 8, 12, 1024


def _bits(v):
    return int(v).bit_length()


def build_index(X, Y, eps):
    """mirror of index_build(): padded u', (strip,u') word and vmod, sorted by (strip,u'), dense strip table"""
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


def tile_info(ks, sstart, bu, t0, t1, rmax):
    """mirror of tile_info_kernel"""
    sA, sB = int(ks[t0] >> bu), int(ks[t1 - 1] >> bu)
    nse = sB - sA + 4
    r0, r1 = int(sstart[sA]), int(sstart[sB + 3])
    if r1 - r0 > rmax or nse > SMAX or (nse << bu) > 0xFFFFFFFF:
        return sA, r0, r1, 0, 0
    maxlen = int(np.diff(sstart[sA:sA + nse]).max())
    return sA, r0, r1, nse, maxlen.bit_length()


def uniform_lower_bound(W, pa, t, nsteps, last):
    """pa = slot of the largest word known to be < t; returns the first slot whose word is >= t"""
    sb = 2 << nsteps                       # in words * 4 like the kernel's byte steps: compare against 32 bytes
    while sb > 32:
        na = min(pa + sb // 4, last)
        if W[na] < t:
            pa = na
        sb >>= 1
    for step in (8, 4, 2, 1):
        if W[pa + step] < t:
            pa += step
    return pa + 1


def adjacent_count(W, V, sa, tlo, thi, vm, nsteps, last, room, nxt):
    j = uniform_lower_bound(W, sa, tlo, nsteps, last)
    f = 0
    if W[j] <= thi:
        for k in range(4):
            f += int(W[j + k] <= thi and (V[j + k] <= vm if nxt else V[j + k] >= vm))
        if W[j + 3] <= thi:
            a = j + 4
            while f < room and W[a] <= thi:
                f += int(V[a] <= vm if nxt else V[a] >= vm)
                a += 1
    return f


def model_counts(X, Y, eps, cap, tile=1024, rmax=2560):
    ks, vmod, order, sstart, be, bu = build_index(X, Y, eps)
    n = len(ks)
    out = np.zeros(n, dtype=np.int64)
    one = 1 << bu
    stats = {"fallback": 0, "tiles": 0, "queued": 0}
    templated = 2 <= cap <= 9              # CAPT > 0: probing form
    for t0 in range(0, n, tile):
        t1 = min(t0 + tile, n)
        sA, r0, r1, nse, nsteps = tile_info(ks, sstart, bu, t0, t1, rmax)
        stats["tiles"] += 1
        if nse == 0:
            stats["fallback"] += 1
            xs, ys = X[order], Y[order]
            for i in range(t0, t1):
                out[i] = min(int((np.abs(xs - xs[i]) + np.abs(ys - ys[i]) <= eps).sum()), cap)
            continue
        sl0 = G - (r0 & ~3)                # slot of global index j = sl0 + j
        size = sl0 + r1 + GR
        W = np.full(size, -1, dtype=np.int64)      # -1 = never written: reading it is a model error
        V = np.full(size, -1, dtype=np.int64)
        W[sl0 + r0:sl0 + r1] = ks[r0:r1] - ((sA - 1) << bu)
        V[sl0 + r0:sl0 + r1] = vmod[r0:r1]
        assert W[sl0 + r0:sl0 + r1].min() >= 0 and W[sl0 + r0:sl0 + r1].max() < 0xFFFFFFFF
        W[sl0 + r0 - G:sl0 + r0] = 0
        V[sl0 + r0 - G:sl0 + r0] = 0
        W[sl0 + r1:sl0 + r1 + GR] = 0xFFFFFFFF
        V[sl0 + r1:sl0 + r1 + GR] = 0
        assert sl0 + r0 - G >= 0
        S = sstart[sA:sA + nse] + sl0
        last = sl0 + r1
        queue = []
        for i in range(t0, t1):
            me = sl0 + i
            wp = int(W[me])
            lo, hi = wp - eps, wp + eps
            assert lo >= 0 and hi < 0xFFFFFFFF
            if templated:
                c = 1
                for q in range(1, cap):
                    assert W[me - q] >= 0 and W[me + q] >= 0          # probes stay inside data + guards
                    c += int(W[me - q] >= lo) + int(W[me + q] <= hi)
            elif cap > 1:
                sa = int(S[wp >> bu]) - 1
                c = uniform_lower_bound(W, sa, hi + 1, nsteps, last) - uniform_lower_bound(W, sa, lo, nsteps, last)
            else:
                c = 1
            out[i] = min(c, cap)
            if c < cap:
                queue.append((i, me, c))
        stats["queued"] += len(queue)
        for i, me, c in queue:
            wp, vm = int(W[me]), int(V[me])
            srel = wp >> bu
            c += adjacent_count(W, V, int(S[srel - 1]) - 1, wp - one - eps, wp - one + eps, vm, nsteps, last, cap - c, False)
            if c < cap:
                c += adjacent_count(W, V, int(S[srel + 1]) - 1, wp + one - eps, wp + one + eps, vm, nsteps, last, cap - c, True)
            out[i] = min(c, cap)
    res = np.empty(n, dtype=np.int64)
    res[order] = out
    return res, stats


def _cases():
    rng = np.random.default_rng(11)
    # sparse background + dense clumps + duplicates + negative coordinates
    for n, span, eps in ((1500, 200_000, 1000), (1200, 30_000, 500), (900, 5_000_000, 1000), (700, 4000, 7), (600, 2000, 1),
                         (3000, 60_000, 1024), (2500, 20_000, 3000)):
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


@pytest.mark.parametrize("tile,rmax", [(1024, 2560), (256, 640), (64, 4864)])
def test_tiled_region_query_model_matches_oracle(tile, rmax):
    queued = tiles = fallback = 0
    for X, Y, eps in _cases():
        want = spec.neighbour_counts(X, Y, eps)
        for cap in (1, 2, 5, 9, 10, 40, 1 << 30):
            got, st = model_counts(X, Y, eps, cap, tile=tile, rmax=rmax)
            assert np.array_equal(got, np.minimum(want, cap)), (eps, cap, tile, np.flatnonzero(got != np.minimum(want, cap))[:5])
            queued += st["queued"]
            tiles += st["tiles"]
            fallback += st["fallback"]
    assert queued > 0 and fallback < tiles
