"""CPU restatement (numpy/scipy) of the cLoops clustering + scoring hot path.  TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may import this
file; the product (``cloops_b200``) never does.

Each function restates, as order-free data-parallel rules, what the reference's sequential Python
computes under Python-3 dict insertion order.  The restatement is pinned
against the reference itself, executed in the build container through ``oracle/ref_shim.py``:
``oracle/make_golden.py`` writes ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` re-checks
them; ``tests/test_oracle_vs_reference.py`` re-runs a randomized battery whenever ``/root/reference``
is present.

Reference anchors (relative to /root/reference):
  neighbourhood / core definition   cLoops/cDBSCAN.py:186-205, cLoops/cDBSCAN2.py:304-346,364-378
  v1 labels                         cLoops/cDBSCAN.py:128-184
  v2 labels                         cLoops/cDBSCAN2.py:55-192
  block labels                      cLoops/blockDBSCAN.py:69-239
  cluster -> candidate records      cLoops/pipe.py:52-110
  range counts / permuted windows   cLoops/cModel.py:60-161
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components
from scipy.spatial import cKDTree

# --------------------------------------------------------------------------------------------------
# eps-neighbourhood (Manhattan, inclusive) -- cDBSCAN.py:43-51,200-204 ; cDBSCAN2.py:364-378


def neighbour_pairs(X: np.ndarray, Y: np.ndarray, eps: int) -> np.ndarray:
    """All unordered pairs (i<j) of rows with |dX|+|dY| <= eps (inclusive, integers).

    Integer distances make ``d <= eps``  <=>  ``d < eps + 0.5``; the half-unit slack keeps the float
    k-d tree away from the boundary, and the result is re-checked in exact int64 arithmetic."""
    X = np.asarray(X, dtype=np.int64)
    Y = np.asarray(Y, dtype=np.int64)
    if len(X) < 2:
        return np.zeros((0, 2), dtype=np.int64)
    tree = cKDTree(np.stack([X, Y], axis=1).astype(np.float64))
    pairs = tree.query_pairs(r=float(eps) + 0.5, p=1.0, output_type="ndarray").astype(np.int64)
    if len(pairs):
        d = np.abs(X[pairs[:, 0]] - X[pairs[:, 1]]) + np.abs(Y[pairs[:, 0]] - Y[pairs[:, 1]])
        pairs = pairs[d <= eps]
    return pairs


def neighbour_counts(X, Y, eps, pairs=None) -> np.ndarray:
    """n(p) = #{q : d1(p,q) <= eps} INCLUDING p itself (cDBSCAN.py:196 ``result=[pointKey]``;
    cDBSCAN2.py:333 ``n + cell_pt_num`` where the own cell holds p)."""
    n = len(X)
    if pairs is None:
        pairs = neighbour_pairs(X, Y, eps)
    cnt = np.ones(n, dtype=np.int64)
    if len(pairs):
        cnt += np.bincount(pairs[:, 0], minlength=n) + np.bincount(pairs[:, 1], minlength=n)
    return cnt


def _core_components(n, pairs, core):
    """Connected components of the core graph; returns comp id per row (-1 for non-core)."""
    comp = np.full(n, -1, dtype=np.int64)
    idx = np.flatnonzero(core)
    if len(idx) == 0:
        return comp, 0
    remap = np.full(n, -1, dtype=np.int64)
    remap[idx] = np.arange(len(idx))
    if len(pairs):
        cc = pairs[core[pairs[:, 0]] & core[pairs[:, 1]]]
    else:
        cc = pairs
    g = coo_matrix((np.ones(len(cc), dtype=np.int8), (remap[cc[:, 0]], remap[cc[:, 1]])),
                   shape=(len(idx), len(idx)))
    k, lab = connected_components(g, directed=False)
    comp[idx] = lab
    return comp, k


def _border_adjacency(pairs, core, comp):
    """(border row, component) incidences, de-duplicated."""
    if len(pairs) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
    a, b = pairs[:, 0], pairs[:, 1]
    m1 = core[a] & ~core[b]
    m2 = core[b] & ~core[a]
    brow = np.concatenate([b[m1], a[m2]])
    crow = np.concatenate([a[m1], b[m2]])       # the core endpoint
    return brow, crow, comp[crow]


# --------------------------------------------------------------------------------------------------
# v1  (cDBSCAN.py:128-184)  -- SURVEY Appendix A.1


def cdbscan_v1(X, Y, eps, minPts) -> np.ndarray:
    """Labels in row order, -1 = absent from the reference's ``labels`` dict."""
    X = np.asarray(X, np.int64)
    Y = np.asarray(Y, np.int64)
    n = len(X)
    out = np.full(n, -1, dtype=np.int64)
    if n == 0:
        return out
    pairs = neighbour_pairs(X, Y, eps)
    core = neighbour_counts(X, Y, eps, pairs) >= minPts
    comp, k = _core_components(n, pairs, core)
    if k == 0:
        return out
    # seed(K) = core row with the smallest row index; ids ascend with seed row (cDBSCAN.py:134-137)
    seed = np.full(k, n, dtype=np.int64)
    np.minimum.at(seed, comp[core], np.flatnonzero(core))
    order = np.argsort(seed, kind="stable")
    cid = np.empty(k, dtype=np.int64)
    cid[order] = np.arange(k)
    out[core] = cid[comp[core]]
    # border: largest id among seed-adjacent clusters (seed relabels all neighbours, :172-173),
    # else smallest adjacent id (first claim sticks, :179-182)
    brow, crow, bcomp = _border_adjacency(pairs, core, comp)
    if len(brow):
        bid = cid[bcomp]
        is_seed = crow == seed[bcomp]
        best_seed = np.full(n, -1, dtype=np.int64)
        np.maximum.at(best_seed, brow[is_seed], bid[is_seed])
        best_any = np.full(n, k, dtype=np.int64)
        np.minimum.at(best_any, brow, bid)
        touched = np.zeros(n, bool)
        touched[brow] = True
        lab = np.where(best_seed >= 0, best_seed, best_any)
        out[touched] = lab[touched]
    # clusters with < minPts members are deleted, ids keep their gaps (cDBSCAN.py:149-152)
    sizes = np.bincount(out[out >= 0], minlength=k)
    small = sizes < minPts
    out[(out >= 0) & small[np.maximum(out, 0)]] = -1
    return out


# --------------------------------------------------------------------------------------------------
# v2  (cDBSCAN2.py:55-192)  -- SURVEY Appendix A.2


def _first_row_per_group(keys: np.ndarray):
    """group id per row + first row index of each group (dict insertion order semantics)."""
    uniq, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    return inv.reshape(-1), first


def cdbscan_v2(X, Y, eps, minPts, return_info=False):
    X = np.asarray(X, np.int64)
    Y = np.asarray(Y, np.int64)
    n = len(X)
    out = np.full(n, -1, dtype=np.int64)
    info = {"dead": 0}
    if n == 0:
        return (out, info) if return_info else out
    pairs = neighbour_pairs(X, Y, eps)
    core = neighbour_counts(X, Y, eps, pairs) >= minPts
    comp, k = _core_components(n, pairs, core)
    if k == 0:
        return (out, info) if return_info else out
    # rotated floor cells (cDBSCAN2.py:66-70, py2 floor); cellorder = first row of the cell
    u = X - Y
    v = X + Y
    cells = np.stack([u // eps, v // eps], axis=1)
    cell_of, cell_first = _first_row_per_group(cells)
    # rank(K) = min cellorder over cells holding a core point of K (outer loop :117-140)
    rank = np.full(k, n, dtype=np.int64)
    np.minimum.at(rank, comp[core], cell_first[cell_of[core]])
    order = np.argsort(rank, kind="stable")            # attempt order
    pos = np.empty(k, dtype=np.int64)
    pos[order] = np.arange(k)                          # position of component in attempt order
    ncore = np.bincount(comp[core], minlength=k)
    brow, _, bcomp = _border_adjacency(pairs, core, comp)
    # unique (border,row) incidences sorted by (row, attempt position)
    if len(brow):
        inc = np.unique(np.stack([brow, pos[bcomp]], axis=1), axis=0)
        brow_u, bpos_u = inc[:, 0], inc[:, 1]
    else:
        brow_u = bpos_u = np.zeros(0, np.int64)
    # sequential survival in attempt order (:180-185): a cluster takes every still-free adjacent
    # border point; released if core+border < minPts.  Only clusters with < minPts core points can die.
    alive = np.ones(k, bool)
    owner = np.full(n, -1, dtype=np.int64)             # attempt position owning a border row
    # fast path: assign lowest-position adjacency, then repair around small clusters sequentially
    start = np.flatnonzero(np.r_[True, brow_u[1:] != brow_u[:-1]]) if len(brow_u) else np.zeros(0, np.int64)
    if len(brow_u):
        owner[brow_u[start]] = bpos_u[start]
    small_pos = np.sort(pos[np.flatnonzero(ncore < minPts)])
    if len(small_pos):
        # adjacency lists restricted to rows touching a small cluster
        by_pos: dict = {}
        for r, p in zip(brow_u.tolist(), bpos_u.tolist()):
            by_pos.setdefault(p, []).append(r)
        row_adj: dict = {}
        for r, p in zip(brow_u.tolist(), bpos_u.tolist()):
            row_adj.setdefault(r, []).append(p)        # ascending p (np.unique sorted)
        ncore_pos = np.empty(k, dtype=np.int64)
        ncore_pos[pos] = ncore
        for p in small_pos.tolist():
            rows = by_pos.get(p, [])
            mine = [r for r in rows if owner[r] == p]
            if ncore_pos[p] + len(mine) < minPts:
                alive[p] = False
                info["dead"] += 1
                for r in mine:                         # released rows fall to the next alive adjacency
                    nxt = -1
                    for q in row_adj[r]:
                        if q > p and alive[q]:
                            nxt = q
                            break
                    owner[r] = nxt
    # final ids: alive clusters dense in attempt order (:184-185)
    newid = np.cumsum(alive) - 1
    comp_pos = pos[comp[core]]
    out[core] = np.where(alive[comp_pos], newid[comp_pos], -1)
    has = owner >= 0
    out[has] = newid[owner[has]]
    return (out, info) if return_info else out


# --------------------------------------------------------------------------------------------------
# blockDBSCAN (blockDBSCAN.py:69-239)  -- SURVEY Appendix A.3


def blockdbscan(X, Y, eps, minPts) -> np.ndarray:
    X = np.asarray(X, np.int64)
    Y = np.asarray(Y, np.int64)
    n = len(X)
    out = np.full(n, -1, dtype=np.int64)
    if n == 0:
        return out
    cx = (X - X.min()) // eps + 1                      # blockDBSCAN.py:74-82
    cy = (Y - Y.min()) // eps + 1
    cells = np.stack([cx, cy], axis=1)
    cell_of, cell_first = _first_row_per_group(cells)
    nc = len(cell_first)
    cnt = np.bincount(cell_of, minlength=nc)
    sx = np.zeros(nc, np.int64)
    sy = np.zeros(nc, np.int64)
    np.add.at(sx, cell_of, X)
    np.add.at(sy, cell_of, Y)
    cenx = sx // cnt                                   # :132-138 (py2 floor)
    ceny = sy // cnt
    ccx = cx[cell_first]
    ccy = cy[cell_first]
    # 8-adjacent occupied cell pairs
    key = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(ccx, ccy))}
    ea, eb = [], []
    for i, (a, b) in enumerate(zip(ccx.tolist(), ccy.tolist())):
        for da, db in ((1, -1), (1, 0), (1, 1), (0, 1)):
            j = key.get((a + da, b + db))
            if j is not None:
                ea.append(i)
                eb.append(j)
    ea = np.array(ea, dtype=np.int64)
    eb = np.array(eb, dtype=np.int64)
    conn = np.zeros(len(ea), bool)
    if len(ea):
        conn |= (np.abs(cenx[ea] - cenx[eb]) + np.abs(ceny[ea] - ceny[eb])) <= eps   # :232
        # any point pair <= eps between the two cells (:204-213,236)
        pairs = neighbour_pairs(X, Y, eps)
        if len(pairs):
            ca, cb = cell_of[pairs[:, 0]], cell_of[pairs[:, 1]]
            m = ca != cb
            lo = np.minimum(ca[m], cb[m])
            hi = np.maximum(ca[m], cb[m])
            pk = set((lo * nc + hi).tolist())
            elo = np.minimum(ea, eb)
            ehi = np.maximum(ea, eb)
            conn |= np.array([int(a) * nc + int(b) in pk for a, b in zip(elo, ehi)], dtype=bool)
    ea, eb = ea[conn], eb[conn]
    near = cnt.copy()
    np.add.at(near, ea, cnt[eb])
    np.add.at(near, eb, cnt[ea])
    corec = near >= minPts                             # :181,191
    idx = np.flatnonzero(corec)
    if len(idx) == 0:
        return out
    remap = np.full(nc, -1, np.int64)
    remap[idx] = np.arange(len(idx))
    m = corec[ea] & corec[eb]
    g = coo_matrix((np.ones(m.sum(), np.int8), (remap[ea[m]], remap[eb[m]])), shape=(len(idx), len(idx)))
    k, lab = connected_components(g, directed=False)
    rank = np.full(k, n, dtype=np.int64)
    np.minimum.at(rank, lab, cell_first[idx])
    order = np.argsort(rank, kind="stable")
    cid = np.empty(k, np.int64)
    cid[order] = np.arange(k)
    clab = np.full(nc, -1, np.int64)
    clab[idx] = cid[lab]
    # non-core cell: largest id among connected core neighbours (:195-198, last toucher wins)
    best = np.full(nc, -1, np.int64)
    m1 = corec[ea] & ~corec[eb]
    np.maximum.at(best, eb[m1], clab[ea[m1]])
    m2 = corec[eb] & ~corec[ea]
    np.maximum.at(best, ea[m2], clab[eb[m2]])
    clab = np.where(corec, clab, best)
    return clab[cell_of]


# --------------------------------------------------------------------------------------------------
# cluster -> candidate records (pipe.py:76-109)  -- SURVEY Appendix A.4


def cluster_records(X, Y, labels):
    """Returns (inter[K1,5], self[K2,5]) rows ``[minX,maxX,minY,maxY,label]`` in ascending label
    order, plus boolean row masks of the members of inter / self clusters."""
    X = np.asarray(X, np.int64)
    Y = np.asarray(Y, np.int64)
    labels = np.asarray(labels, np.int64)
    inter, selfl = [], []
    in_i = np.zeros(len(X), bool)
    in_s = np.zeros(len(X), bool)
    for lab in np.unique(labels[labels >= 0]).tolist():
        m = labels == lab
        x0, x1, y0, y1 = int(X[m].min()), int(X[m].max()), int(Y[m].min()), int(Y[m].max())
        if x0 == x1 or y0 == y1:                       # pipe.py:83-85
            continue
        if x1 < y0:                                    # pipe.py:97
            inter.append([x0, x1, y0, y1, lab])
            in_i |= m
        else:
            selfl.append([x0, x1, y0, y1, lab])
            in_s |= m
    return (np.array(inter, np.int64).reshape(-1, 5), np.array(selfl, np.int64).reshape(-1, 5), in_i, in_s)


# --------------------------------------------------------------------------------------------------
# permuted-local-background range counts (cModel.py:60-143)  -- SURVEY Appendix A.5


def nearby_windows(iva, ivb, win=5):
    """cModel.py:83-105 with py2 integer division."""
    ca = (iva[0] + iva[1]) // 2
    cb = (ivb[0] + ivb[1]) // 2
    sa = (iva[1] - iva[0]) // 2
    sb = (ivb[1] - ivb[0]) // 2
    step = (sa + sb) // 2
    ivas, ivbs = [], []
    for i in range(-win, win + 1):
        if i == 0:
            continue
        ivas.append([max(0, ca + i * step - sa), max(0, ca + i * step + sa)])
        ivbs.append([max(0, cb + i * step - sb), max(0, cb + i * step + sb)])
    return ivas, ivbs


def _in(W, X, Y):
    return ((X >= W[0]) & (X <= W[1])) | ((Y >= W[0]) & (Y <= W[1]))


def range_counts(X, Y, iva, ivb, win=5) -> np.ndarray:
    """The 123 integers every statistic of getMultiplePsFdr is a function of:
    ``[ra, rb, rab, na_0..na_9, nb_0..nb_9, C_00..C_99]`` (C row-major, i over A windows)."""
    X = np.asarray(X, np.int64)
    Y = np.asarray(Y, np.int64)
    ina = _in(iva, X, Y)
    inb = _in(ivb, X, Y)
    rab = int(np.sum((X >= iva[0]) & (X <= iva[1]) & (Y >= ivb[0]) & (Y <= ivb[1])))
    ivas, ivbs = nearby_windows(iva, ivb, win)
    ma = [_in(w, X, Y) for w in ivas]
    mb = [_in(w, X, Y) for w in ivbs]
    out = [int(ina.sum()), int(inb.sum()), rab]
    out += [int(m.sum()) for m in ma]
    out += [int(m.sum()) for m in mb]
    for a in ma:
        for b in mb:
            out.append(int(np.sum(a & b)))
    return np.array(out, dtype=np.int64)


def stats_from_counts(c: np.ndarray, N: int):
    """cModel.py:113-114,129-161 evaluated on the 123 integers with the same numpy/scipy calls."""
    from scipy.stats import binom, hypergeom, poisson
    ra, rb, rab = int(c[0]), int(c[1]), int(c[2])
    na = c[3:13]
    nb = c[13:23]
    C = c[23:123].reshape(10, 10)
    hyp = max([1e-300, hypergeom.sf(rab - 1.0, N, ra, rb)])
    rabs, nbps = [], []
    for i in range(10):
        nralen = float(na[i])
        for j in range(10):
            nrab = float(C[i, j])
            if nrab > 0:
                rabs.append(nrab)
                nbps.append(nrab / (nralen * int(nb[j])))
            else:
                nbps.append(0.0)
                rabs.append(0.0)
    rabs = np.array(rabs)
    fdr = len(rabs[rabs > rab]) / float(len(rabs))
    mrabs = float(np.mean(rabs))
    if mrabs > 0:
        es = rab / np.mean(rabs[rabs > 0])
    else:
        es = np.inf
    pop = max([1e-300, poisson.sf(rab - 1.0, mrabs)])
    bp = np.mean(nbps) * ra * rb / N
    nbp = max([1e-300, binom.sf(rab - 1.0, N - rab, bp)])
    return ra, rb, rab, es, fdr, hyp, pop, nbp


# --------------------------------------------------------------------------------------------------
# The same counts the way the reference obtains them (cModel.py:31-80,108-143): sorted coordinate
# arrays + searchsorted slices + set algebra on row indices.  Used for the CPU baseline timing; the
# brute-force range_counts() above stays the independent check.


class CoverageIndex:
    def __init__(self, X, Y):
        X = np.asarray(X, np.int64)
        Y = np.asarray(Y, np.int64)
        self.N = len(X)
        self.ox = np.argsort(X, kind="stable")          # cModel.py:41 np.sort(cs) + row lists (:38-40)
        self.kx = X[self.ox]
        self.oy = np.argsort(Y, kind="stable")
        self.ky = Y[self.oy]

    def _rows(self, keys, order, iv):                   # getCounts, cModel.py:60-69
        a = np.searchsorted(keys, iv[0], side="left")
        b = np.searchsorted(keys, iv[1], side="right")
        return order[a:b]

    def either(self, iv):                               # source-union-target, cModel.py:73-78,118-127
        return np.union1d(self._rows(self.kx, self.ox, iv), self._rows(self.ky, self.oy, iv))

    def range_counts(self, iva, ivb, win=5) -> np.ndarray:
        ra = len(self.either(iva))
        rb = len(self.either(ivb))
        rab = len(np.intersect1d(self._rows(self.kx, self.ox, iva), self._rows(self.ky, self.oy, ivb), assume_unique=True))
        ivas, ivbs = nearby_windows(iva, ivb, win)
        sa = [self.either(w) for w in ivas]
        sb = [self.either(w) for w in ivbs]
        out = [ra, rb, rab] + [len(s) for s in sa] + [len(s) for s in sb]
        for a in sa:
            for b in sb:
                out.append(len(np.intersect1d(a, b, assume_unique=True)))
        return np.array(out, dtype=np.int64)


def hot_path_cpu(X, Y, eps, minPts):
    """One pass of the whole hot path on the CPU (clusterer v2 -> candidate records -> range counts of
    every inter-ligation candidate); returns (labels, inter records, counts[K,123])."""
    lab = cdbscan_v2(X, Y, eps, minPts)
    inter, selfl, in_i, in_s = cluster_records(X, Y, lab)
    cov = CoverageIndex(X, Y)
    counts = np.zeros((len(inter), 123), np.int64)
    for k, r in enumerate(inter):
        counts[k] = cov.range_counts([max(0, int(r[0])), int(r[1])], [max(0, int(r[2])), int(r[3])])
    return lab, inter, counts
