"""The multi-round engine on resident chromosomes (pipe.call_loops: what bench.py times for configs 3 / 4) against an
oracle pipeline built from the C oracle and the reference's numpy cut-off estimate: per round the cut filter, cDBSCAN2
labels, candidate records, pooled dis / dss -> estIntSelCutFrag (cLoops/pipe.py:247-281, ests.py:36-61), then combineTwice
and filterClusterByDis.  The device-side statistics (histogram + log2 moments, one reduction per round) must give the same
integer cut-offs, and the candidates that reach scoring must be the same boxes in the same order."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def oracle_rounds(chroms, eps_list, mp_list):
    from cloops_b200 import ests
    from oracle import coracle
    cut, cuts = 0, [0]
    recs = {name: [] for name, _, _ in chroms}
    trace = []
    for ep in eps_list:
        for m in mp_list:
            dis, dss, contributed = [], [], 0
            for name, X, Y in chroms:
                d = Y - X
                act = d >= cut if cut > 0 else np.ones(len(d), bool)
                lab = np.full(len(X), -1, np.int32)
                if act.any():
                    lab[act] = coracle.dbscan(X[act], Y[act], ep, m, coracle.V2)
                bbox, size, kind = coracle.cluster_records(X, Y, lab)
                inter = bbox[kind == 1]
                if len(inter) == 0:                                 # pipe.py:121-122
                    continue
                contributed += 1
                recs[name].append(inter)
                dis.append(d[np.isin(lab, np.flatnonzero(kind == 1))])
                if cut > 0:
                    dss.append(d[~act])
                if (kind == 2).any():
                    dss.append(d[np.isin(lab, np.flatnonzero(kind == 2))])
            dis = np.concatenate(dis) if dis else np.zeros(0)
            dss = np.concatenate(dss) if dss else np.zeros(0)
            if contributed and len(dis) and len(dss):
                cut = ests.estIntSelCutFrag(dis, dss)[0]
                cuts.append(cut)
            trace.append((ep, m, cut, len(dis), len(dss)))
    final = min(c for c in cuts if c > 0)
    out = {}
    for name, rounds in recs.items():
        if not rounds:
            continue
        seen, keep = set(), []
        for r in rounds:                                            # combineTwice, pipe.py:155-174
            rows = [tuple(x) for x in r.tolist()]
            keep += [x for x in rows if x not in seen]
            seen |= set(rows)
        r = np.array(keep, np.int64).reshape(-1, 4)
        out[name] = r[(r[:, 2] + r[:, 3]) // 2 - (r[:, 0] + r[:, 1]) // 2 >= final]     # filterClusterByDis, pipe.py:130-143
    return out, final, trace


@pytest.mark.parametrize("eps_list,mp_list,dens", [([2000, 4000], [12, 6], 0.4), ([5000, 7500], [30, 20], 1.0)])
def test_rounds_match_oracle_pipeline(eps_list, mp_list, dens, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cloops_b200 import pipe, synth
    chroms = []
    for k, (n, L) in enumerate(((300_000, 5_000_000), (180_000, 3_000_000), (90_000, 2_000_000))):
        X, Y = synth.chromosome(int(n * dens), L, seed=900 + k, loop_frac=0.1, sigma=800.0)
        chroms.append(("chr%d" % (k + 1), X.astype(np.int64), Y.astype(np.int64)))
    want, final, trace = oracle_rounds(chroms, eps_list, mp_list)
    pipe._Resident.clear()
    cfs = [pipe._Resident.register(name, X, Y) for name, X, Y in chroms]
    seen = []
    orig = pipe._round
    monkeypatch.setattr(pipe, "_round", lambda fs, e, m, c, w=None: (lambda r: (seen.append((e, m, r[4], r[2], r[3])), r)[1])(orig(fs, e, m, c, w)))
    run = pipe.call_loops(cfs, eps_list, mp_list, hic=1, tail=False)
    pipe._Resident.clear()
    assert [(e, m, c, nd, ns) for e, m, c, nd, ns in seen] == [(e, m, c, nd, ns) for e, m, c, nd, ns in trace]
    assert run["cut"] == final
    assert sorted(k[0] for k in run["dataI"]) == sorted(want)
    for key, v in run["dataI"].items():
        assert np.array_equal(np.asarray(v["records"], np.int64), want[key[0]]), key


@pytest.mark.parametrize("variant", [2, 1])
def test_pass_from_full_index_equals_fresh_build(variant):
    """cloops_pass_run_base (the round's index as a compaction of the chromosome's cut = 0 index) must give exactly what
    cloops_pass_run_stats gives with a freshly built index: labels in index order, coordinates, records, the round
    statistics -- for no cut, ordinary cuts, a cut that removes every row, and negative / mixed coordinates."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cloops_b200 import _lib, device, synth
    X, Y = synth.chromosome(250_000, 4_000_000, seed=77, loop_frac=0.15, sigma=600.0)
    sets = [(X, Y), (X - 1_500_000, Y - 1_500_000)]
    for X, Y in sets:
        dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
        for eps, mp in ((1500, 6), (6000, 25)):
            base = device.Index(dx, dy, eps, 0)
            for cut in (0, 900, 7000, 40_000_000):
                out = []
                for use_base in (False, True):
                    hist = torch.zeros(_lib.ROUND_HIST_BINS + 1, dtype=torch.int32, device="cuda")
                    mom = torch.zeros(_lib.ROUND_MOM, dtype=torch.float64, device="cuda")
                    p = device.Pass(dx, dy, eps, mp, variant, cut, score=False, stats=(hist, mom), base=base if use_base else None)
                    bbox, size, kind = p.records()
                    out.append((p.n_members, p.info["n_clusters"], p.info["n_dead"], bbox, size, kind, p.xs.cpu().numpy(), p.ys.cpu().numpy(),
                                p.labels_sorted.cpu().numpy(), hist.cpu().numpy(), mom.cpu().numpy()))
                    p.close()
                for a, b in zip(*out):
                    assert np.array_equal(a, b), (eps, mp, cut)
            base.close()


def test_workspace_reuse_and_release():
    """Scratch memory comes from a per-stream workspace that is rewound, grown and replaced between calls: passes of
    different sizes in a row (small after large, large after small), on two streams, must equal a run in which every
    call starts from released workspaces -- and cloops_workspace_release must leave the library usable."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cloops_b200 import _lib, device, synth
    L = _lib.lib()
    shapes = [(40_000, 900_000, 5), (600_000, 9_000_000, 6), (3_000, 200_000, 4), (1_200_000, 14_000_000, 8), (600_000, 9_000_000, 6)]
    data = [synth.chromosome(n, span, seed=100 + k, loop_frac=0.2, sigma=500.0) for k, (n, span, _) in enumerate(shapes)]

    def run(k, stream=None):
        n, span, mp = shapes[k]
        X, Y = data[k]
        with torch.cuda.stream(stream):                      # None: the current stream
            dx, dy = device.to_device_i32(X), device.to_device_i32(Y)
            p = device.Pass(dx, dy, 1000, mp, 2, 500 if k % 2 else 0, score=True)
            bbox, size, kind = p.records()
            out = (p.n_members, bbox, size, kind, p.labels_sorted.cpu().numpy(), p.counts.cpu().numpy())
            p.close()
        return out

    fresh = []
    for k in range(len(shapes)):
        _lib.check(L.cloops_workspace_release())
        fresh.append(run(k))
    side = torch.cuda.Stream()
    for order in (range(len(shapes)), reversed(range(len(shapes)))):
        for k in order:
            for st in (None, side):
                got = run(k, st)
                for a, b in zip(got, fresh[k]):
                    assert np.array_equal(a, b), (k, st is side)
    torch.cuda.synchronize()
    _lib.check(L.cloops_workspace_release())
    for a, b in zip(run(1), fresh[1]):
        assert np.array_equal(a, b)
