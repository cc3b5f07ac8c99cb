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

# device-side cut-off reduction across ranks: per-chromosome contributions to the round accumulators, summed with ONE
# all-reduce (CPU tensors over gloo here, CUDA tensors over NCCL on the box), then the reference's estimate (ests.py:36-61)
import torch  # noqa: E402

from cloops_b200 import _lib, ests  # noqa: E402

rng = np.random.default_rng(5)
dist_i = {f: rng.integers(5000, 400000, 300 + 40 * k) for k, f in enumerate(files)}
dist_s = {f: rng.integers(40, 3000, 500 + 70 * k) for k, f in enumerate(files)}


class CpuAcc(pipe._RoundAcc):
    """The accumulators as CPU tensors; middle() restated with numpy (the CUDA one is cloops_round_middle)."""

    def __init__(self):
        self.hist = torch.zeros(_lib.ROUND_HIST_BINS + 1, dtype=torch.int32)
        self.mom = torch.zeros(_lib.ROUND_MOM, dtype=torch.float64)

    def middle(self):
        cum = np.cumsum(self.hist.numpy().astype(np.int64))
        k = int(self.mom[3])
        return int(np.searchsorted(cum, (k - 1) // 2, side="right")), int(np.searchsorted(cum, k // 2, side="right")), self.mom.tolist()


acc = CpuAcc()
pipe._RoundAcc.get = classmethod(lambda cls: acc)
clustered = []


def fake_cluster(f, eps, minPts, cut, acc):
    clustered.append(f)
    c = f.split("-")[0]
    if c == "chrE":                                       # no inter-ligation clusters: contributes nothing (pipe.py:121-122)
        return (c, c), np.zeros((0, 4), np.int32), 3
    li, ls = np.log2(dist_i[f].astype(np.float64)), np.log2(dist_s[f].astype(np.float64))
    acc.mom += torch.tensor([len(li), li.sum(), (li * li).sum(), len(ls), ls.sum(), (ls * ls).sum(), len(li), len(ls), 1] + [0] * 7, dtype=torch.float64)
    acc.hist += torch.from_numpy(np.bincount(dist_s[f], minlength=_lib.ROUND_HIST_BINS + 1).astype(np.int32))
    return (c, c), np.array([[1, 2, 9000030 + eps, 9000040 + eps]], np.int32), 1


pipe._cluster_chrom = fake_cluster
w = [weights[f] for f in files]
dataI, n_self, n_dis, n_dss, cut, n_contrib = pipe._round(files, 1000, 5, 0, w)
used = files[:4]
want = ests.estIntSelCutFrag(np.concatenate([dist_i[f] for f in used]), np.concatenate([dist_s[f] for f in used]))[0]
assert cut == want, (cut, want)
assert n_contrib == 4 and n_dis == sum(len(dist_i[f]) for f in used) and n_dss == sum(len(dist_s[f]) for f in used)
assert sorted(k[0] + "-" + k[0] + ".jd" for k in dataI) == sorted(set(clustered) & set(used))      # local chromosomes only
got = dist.all_gather_concat(torch.arange(3 + dist.rank(), dtype=torch.int32))
assert got.tolist() == [0, 1, 2, 0, 1, 2, 3]
assert ests.cut_from_round([10, 100.0, 1001.0, 10, 50.0, 251.0] + [0] * 10, 1 << 20, 5) is None          # median outside the histogram
assert pipe._combine_rounds([np.array([[1, 2, 3, 4], [5, 6, 7, 8]]), np.array([[5, 6, 7, 8], [9, 9, 9, 9], [9, 9, 9, 9]]),
                             np.array([[1, 2, 3, 4], [0, 0, 0, 0]])]).tolist() == [[1, 2, 3, 4], [5, 6, 7, 8], [9, 9, 9, 9], [9, 9, 9, 9], [0, 0, 0, 0]]

# pipe() itself under two ranks: rank 0 alone decides about the output directory and every rank follows.
# (1) existing directory -> every rank returns, no collective is left hanging; (2) fresh directory -> the run completes.
pipe.parseRawBedpe2 = lambda fs, fout, chroms, cut, log: list(files)
pipe._weights = lambda fs: [weights[f] for f in fs]
calls = []
orig_round = pipe._round
pipe._round = lambda fs, ep, m, cut, weights=None: (calls.append((ep, m, cut)), orig_round(fs, ep, m, cut, weights))[1]
exists = os.path.join(out, "exists")
if dist.rank() == 0:
    os.mkdir(exists)
dist.barrier()
assert pipe.pipe(["x.bedpe"], exists, [1000], [5]) is None and calls == []
dist.barrier()
fresh = os.path.join(out, "fresh")
seen_records = {}
pipe.cModel.countCandidates = lambda f, records, minPts, cut: (f, np.asarray(records).tolist())
pipe.cModel.tableFromCounts = lambda c: (seen_records.__setitem__(c[0], c[1]), pd.DataFrame(
    {"ES": [3.0], "FDR": [0.0], "hypergeometric_p-value": [1e-20], "poisson_p-value": [1e-9], "binomial_p-value": [1e-9]}, index=["%s-0" % c[0]]))[1]
pipe.pipe(["x.bedpe"], fresh, [1000, 2000], [5], tmp=1)
assert [c[:2] for c in calls] == [(1000, 5), (2000, 5)] and calls[1][2] == want, calls
for f, recs in seen_records.items():                                # both rounds' records of the rank's own chromosomes, merged
    assert recs == [[1, 2, 9001030, 9001040], [1, 2, 9002030, 9002040]], recs
dist.barrier()
if dist.rank() == 0:
    assert os.path.isdir(fresh) and os.path.isfile(fresh + ".loop")
    tab = pd.read_csv(fresh + ".loop", sep="\t", index_col=0)
    assert list(tab.index) == ["%s-0" % f for f in files[:4]], list(tab.index)      # file order, from all ranks
    open(os.path.join(out, "ok_pipe"), "w").write("ok")
