"""Worker for tests/test_host_logic.py::test_two_rank_gloo (launched under torchrun, gloo, CPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

from cloops_b200 import dist, pipe  # noqa: E402

out = sys.argv[1]
dist.init_from_env("gloo")
assert dist.world() == 2
files = ["chrA-chrA.jd", "chrB-chrB.jd", "chrC-chrC.jd", "chrD-chrD.jd", "chrE-chrE.jd"]
weights = {"chrA-chrA.jd": 50, "chrB-chrB.jd": 40, "chrC-chrC.jd": 30, "chrD-chrD.jd": 30, "chrE-chrE.jd": 10}
seen = []


def fake_single(f, eps, minPts, cut=0):
    seen.append(f)
    c = f.split("-")[0]
    k = weights[f]
    recs = [[c, 10 * i, 10 * i + 5, c, 1000 + 10 * i, 1000 + 10 * i + 5] for i in range(k // 10)]
    return (c, c), f, recs if c != "chrE" else [], [[c, 1, 2, c, 3, 4]], np.array([float(k)] * 3), np.array([float(k + 1)] * 2)


orig_assign = dist.assign
dist.assign = lambda items, w=None, nranks=None: orig_assign(items, [weights.get(i, 1) for i in items] if w is None else w, nranks)
pipe._single = fake_single
dataI, dataS, dis, dss = pipe.runDBSCAN(files, 1000, 5, 0)
# every rank sees the merged result in FILE order (pipe.py:120-127)
assert list(dataI.keys()) == [("chrA", "chrA"), ("chrB", "chrB"), ("chrC", "chrC"), ("chrD", "chrD")], dataI.keys()
assert dis.tolist() == [50.0] * 3 + [40.0] * 3 + [30.0] * 6, dis
assert len(dataS) == 4 and len(dss) == 8
mine = sorted(seen)
got = dist.merge_in_order([0, 1], {dist.rank(): mine})
assert sorted(got[0] + got[1]) == sorted(files) and not set(got[0]) & set(got[1])
assert abs(sum(weights[f] for f in got[0]) - sum(weights[f] for f in got[1])) <= 10      # LPT balance
# scoring fan-out: tables come back in key order on every rank
pipe.getIntSig = lambda f, records, minPts, cut: pd.DataFrame({"ES": [3.0], "FDR": [0.0], "hypergeometric_p-value": [1e-20],
                                                                "poisson_p-value": [1e-9], "binomial_p-value": [1e-9]},
                                                               index=["%s-0" % f])
rc = pipe.runStat(dataI, [5], 0, 1, os.path.join(out, "t"), 0)
assert rc == 0
dist.barrier()
if dist.rank() == 0:
    tab = pd.read_csv(os.path.join(out, "t.loop"), sep="\t", index_col=0)
    assert list(tab.index) == ["%s-0" % f for f in files[:4]], list(tab.index)
    assert list(tab["significant"]) == [1.0] * 4
    open(os.path.join(out, "ok"), "w").write("ok")
assert dist.broadcast_object({"cut": 4601} if dist.rank() == 0 else None) == {"cut": 4601}

# device-side cut-off reduction across ranks (CPU tensors over gloo here, CUDA tensors over NCCL on the box)
import torch  # noqa: E402

from cloops_b200 import ests  # noqa: E402

rng = np.random.default_rng(5)
dist_i = {f: rng.integers(5000, 400000, 300 + 40 * k) for k, f in enumerate(files)}
dist_s = {f: rng.integers(40, 3000, 500 + 70 * k) for k, f in enumerate(files)}


def mom(v):
    x = np.log2(v.astype(np.float64))
    return (len(x), float(x.mean()), float(((x - x.mean()) ** 2).sum()))


def fake_stats(f, eps, minPts, cut=0):
    c = f.split("-")[0]
    recs = [[c, 1, 2, c, 30, 40]] if c != "chrE" else []
    return (c, c), f, recs, [], len(dist_i[f]), len(dist_s[f]), mom(dist_i[f]), mom(dist_s[f]), torch.from_numpy(dist_s[f].astype(np.int32))


pipe._single_stats = fake_stats
dataI, dataS, n_dis, n_dss, cut = pipe._round(files, 1000, 5, 0)
used = files[:4]                                         # chrE has no inter-ligation records: excluded (pipe.py:121-122)
want = ests.estIntSelCutFrag(np.concatenate([dist_i[f] for f in used]), np.concatenate([dist_s[f] for f in used]))[0]
assert cut == want, (cut, want)
assert n_dis == sum(len(dist_i[f]) for f in used) and n_dss == sum(len(dist_s[f]) for f in used)
got = dist.all_gather_concat(torch.arange(3 + dist.rank(), dtype=torch.int32))
assert got.tolist() == [0, 1, 2, 0, 1, 2, 3]

# pipe() itself under two ranks: rank 0 alone decides about the output directory and every rank follows.
# (1) existing directory -> every rank returns, no collective is left hanging; (2) fresh directory -> the run completes.
pipe.parseRawBedpe2 = lambda fs, fout, chroms, cut, log: list(files)
calls = []
orig_round = pipe._round
pipe._round = lambda fs, ep, m, cut: (calls.append((ep, m, cut)), orig_round(fs, ep, m, cut))[1]
exists = os.path.join(out, "exists")
if dist.rank() == 0:
    os.mkdir(exists)
dist.barrier()
assert pipe.pipe(["x.bedpe"], exists, [1000], [5]) is None and calls == []
dist.barrier()
fresh = os.path.join(out, "fresh")
pipe.pipe(["x.bedpe"], fresh, [1000, 2000], [5], tmp=1)
assert [c[:2] for c in calls] == [(1000, 5), (2000, 5)] and calls[1][2] == want, calls
dist.barrier()
if dist.rank() == 0:
    assert os.path.isdir(fresh) and os.path.isfile(fresh + ".loop")
    open(os.path.join(out, "ok_pipe"), "w").write("ok")
