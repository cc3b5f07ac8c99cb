#!/usr/bin/env python
"""Benchmark of the cLoops hot path on B200 (BASELINE.json: "PETs/sec clustered+scored (bit-exact) at 1/2/4/8 B200 vs ref CPU").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl ours|reference] [--pets P]

Default workload = the north-star configuration, BASELINE.json configs[3]: "synthetic deep Hi-C 200M cis PETs, -m 3
(eps 5000/7500/10000, minPts 50/40/30/20)", 23 hg38-proportional chromosomes (cloops_b200/synth.py).

A step (configs 3, 4) = the whole multi-round pipeline of cLoops/pipe.py:247-284 on resident chromosomes: 12 (config 3: 8)
clustering rounds over every chromosome with the pooled distance cut-off fed forward between rounds, candidate merging
(combineTwice) and filtering, the coverage models and the permuted-background range counts of every candidate.  At N GPUs the
chromosomes are packed onto the ranks by PET count (LPT; strong scaling, ideal 1.97 / 3.93 / 7.66x) and the per-round
cut-off statistics cross the GPUs in one NCCL all-reduce per round, inside the timed region.
  value    = total PETs / max-over-ranks device time of a step, inputs resident in HBM, results (candidates, counts) on the host;
             the scipy / de-duplication tail is NOT in it.
  e2e      = the same from pinned HOST coordinates through to the marked loop table on rank 0: H2D, rounds, counts, D2H, the
             host statistics tail and the gather of the tables -- SURVEY 8d's "wall up to the concatenated loop table".
  roofline = region-query kernel: sum over every launch of a step of 12 B x active PETs / sum of its launch times.
Config 2 (one chromosome, one round; weak replicas at N > 1) and config 5 (8 x 6 eps x minPts sweep on one 50 M-PET
chromosome) are kept as --config 2 / 5.

--impl reference times the REFERENCE'S OWN code (oracle/_ref through oracle/ref_shim.py: runDBSCAN rounds, estIntSelCutFrag,
combineTwice, filterClusterByDis, runStat with joblib over all host cores) on a bounded same-density sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    2: dict(kind="single", pets=10_000_000, eps=[1000], minPts=[5], hic=0,
            name="synthetic ChIA-PET 10M cis PETs, single chromosome, eps=1000 minPts=5 (BASELINE.json configs[1])"),
    3: dict(kind="genome", pets=100_000_000, eps=[2500, 5000, 7500, 10000], minPts=[30, 20], hic=1,
            name="synthetic HiChIP 100M cis PETs, 23 chroms, -m 4 (BASELINE.json configs[2])"),
    4: dict(kind="genome", pets=200_000_000, eps=[5000, 7500, 10000], minPts=[50, 40, 30, 20], hic=1,
            name="synthetic deep Hi-C 200M cis PETs, 23 chroms, -m 3 (BASELINE.json configs[3])"),
    5: dict(kind="sweep", pets=50_000_000, eps=[500, 1000, 2000, 2500, 5000, 7500, 10000, 20000], minPts=[50, 40, 30, 20, 10, 5], hic=0,
            name="eps x minPts grid sweep (8x6) on 50M PETs, single chromosome (BASELINE.json configs[4])"),
}
CHROM_LEN_SINGLE = 249_000_000
METRIC = "PETs/sec clustered+scored"


# ------------------------------------------------------------------------------------------------------------------------
def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def load_traffic(name):
    """DRAM bytes per launch of the region-query kernel from the committed ncu capture of the same workload, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_count_kernel_metrics.json")) as fh:
            return json.load(fh).get(name)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def visible_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def quiet_logs():
    import logging
    from cloops_b200 import pipe
    lg = logging.getLogger("cloops_bench")
    lg.handlers, lg.propagate = [logging.NullHandler()], False
    pipe.logger = lg
    pipe.cModel.QUIET = True
    sys.stderr = open(os.devnull, "w") if os.environ.get("CLOOPS_BENCH_VERBOSE") is None else sys.stderr


def generate(jobs, fn):
    """jobs -> [fn(job)] with a few host threads (numpy releases the GIL in the generators)."""
    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as ex:
        return list(ex.map(fn, jobs))


# ------------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own pipeline code on a bounded, same-density sample


def reference_sample(cfg, args, cores, have=None):
    """[(name, X, Y)]: for the genome configs one window (PETs whose left anchor lies in the first part of the chromosome,
    same density as the full chromosome) of each of the first `cores` chromosomes; for the single-chromosome configs one
    window of that chromosome (the reference cannot use more than one core on one chromosome, cLoops/pipe.py:117)."""
    from cloops_b200 import synth
    if cfg["kind"] == "genome":
        nchr = max(1, min(cores, 23))
        per = int(os.environ.get("CLOOPS_REF_SAMPLE", "30000"))
        counts = synth.genome_counts(args.pets)

        def one(ci):
            if have is not None and ci in have:
                name, X, Y = have[ci]
            else:
                name, X, Y = synth.genome_chrom(args.pets, args.config, ci)
            lim = int(synth.HG38[ci] * min(1.0, per / max(1, counts[ci])))
            m = X < lim
            return name, X[m].astype(np.int64), Y[m].astype(np.int64)

        return generate(list(range(nchr)), one), "per step: the PETs whose left anchor lies in the first ~%d-PET window of each of %d chromosomes (same density as the full chromosomes)" % (per, nchr)
    if have is not None:
        X, Y = have
    else:
        X, Y = synth.chromosome(args.pets, CHROM_LEN_SINGLE, 20240 + args.config * 100, loop_frac=0.08, sigma=500.0)
    frac = float(os.environ.get("CLOOPS_REF_FRAC", "0.03" if cfg["kind"] == "single" else "0.001"))
    m = X < int(CHROM_LEN_SINGLE * frac)
    return [("chr1", X[m].astype(np.int64), Y[m].astype(np.int64))], "per step: the PETs with X < %.1f%% of the chromosome (same density)" % (frac * 100)


class ReferenceRun:
    """The reference's runDBSCAN rounds + cut feedback + runStat (cLoops/pipe.py:247-284) through oracle/ref_shim.py."""

    def __init__(self, cfg, sample, cores):
        import logging
        import tempfile
        import joblib
        from oracle import ref_shim
        self.ns = ref_shim.load()
        self.cfg, self.cores = cfg, cores
        self.ns.pipe.logger = logging.getLogger("ref_bench")
        self.ns.pipe.logger.handlers, self.ns.pipe.logger.propagate = [logging.NullHandler()], False
        self.tmp = tempfile.mkdtemp(prefix="cloops_ref_")
        self.cfs = []
        for name, X, Y in sample:
            f = os.path.join(self.tmp, "%s-%s.jd" % (name, name))
            joblib.dump(np.stack([np.arange(len(X)), X, Y], axis=1).astype(np.int64), f)     # the reference's own format (io.py:192-203)
            self.cfs.append(f)
        self.pets = sum(len(s[1]) for s in sample)
        self.kind = "reference-shim (%s)" % ref_shim.REF_ROOT

    def step(self):
        import contextlib
        import io
        import joblib
        P = self.ns.pipe
        cfg, cpu = self.cfg, self.cores
        out = os.path.join(self.tmp, "out")
        sink = io.StringIO()
        with joblib.parallel_backend("multiprocessing"), contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            dataI, cut, cuts = {}, 0, [0]
            for ep in cfg["eps"]:
                for m in cfg["minPts"]:
                    d2, s2, dis2, dss2 = P.runDBSCAN(self.cfs, ep, m, cut, cpu)
                    if len(d2) == 0:
                        continue
                    if len(dis2) == 0 or len(dss2) == 0:
                        dataI = P.combineTwice(dataI, d2)
                    else:
                        cut2, frags = P.estIntSelCutFrag(np.array(dis2), np.array(dss2))
                        cuts.append(cut2)
                        cut = cut2
                        dataI = P.combineTwice(dataI, d2)
            cuts = [c for c in cuts if c > 0]
            if cuts and dataI:
                dataI = P.filterClusterByDis(dataI, np.min(cuts))
                P.runStat(dataI, cfg["minPts"], 0, cpu, out, cfg["hic"])
        n = 0
        if os.path.isfile(out + ".loop"):
            n = sum(1 for _ in open(out + ".loop")) - 1
            os.remove(out + ".loop")
        return n

    def close(self):
        import shutil
        shutil.rmtree(self.tmp, ignore_errors=True)


class PortRun:
    """Fallback when oracle/_ref is absent: the C restatement (oracle/coracle.c), one round, one core."""

    def __init__(self, cfg, sample, cores):
        from oracle import coracle
        self.co, self.cfg, self.sample, self.cores = coracle, cfg, sample, 1
        self.pets = sum(len(s[1]) for s in sample)
        self.kind = "port (oracle/coracle.c; oracle/_ref not found)"

    def step(self):
        for name, X, Y in self.sample:
            for ep in self.cfg["eps"]:
                for m in self.cfg["minPts"]:
                    self.co.hot_path(X, Y, ep, m)
        return 0

    def close(self):
        pass


class SweepReferenceRun:
    """Config 5 on the CPU: the reference's cDBSCAN2 class on the sample at the four corner (eps, minPts) pairs of the grid
    (BASELINE.md section 3); PETs/s counts one PET per pair, like the GPU arm."""

    def __init__(self, cfg, sample, cores):
        from oracle import ref_shim
        self.ns = ref_shim.load()
        name, X, Y = sample[0]
        self.mat = np.stack([np.arange(len(X)), X, Y], axis=1).astype(np.int64)
        self.pairs = [(cfg["eps"][0], min(cfg["minPts"])), (cfg["eps"][0], max(cfg["minPts"])),
                      (cfg["eps"][-1], min(cfg["minPts"])), (cfg["eps"][-1], max(cfg["minPts"]))]
        self.pets = len(X) * len(self.pairs)
        self.cores = 1
        self.kind = "reference-shim (%s), cDBSCAN2 at the 4 corner pairs %s" % (ref_shim.REF_ROOT, self.pairs)

    def step(self):
        for ep, m in self.pairs:
            self.ns.cDBSCAN2(self.mat, ep, m)
        return 0

    def close(self):
        pass


def make_cpu_run(cfg, sample, cores):
    from oracle import ref_shim
    if not ref_shim.available():
        return PortRun(cfg, sample, cores)
    return SweepReferenceRun(cfg, sample, cores) if cfg["kind"] == "sweep" else ReferenceRun(cfg, sample, cores)


def workload_config(cfg, args, world):
    c = {"workload": cfg["name"], "config": args.config, "pets_total": args.pets * (world if cfg["kind"] == "single" else 1),
         "eps": cfg["eps"], "minPts": cfg["minPts"], "rounds": len(cfg["eps"]) * len(cfg["minPts"]), "clusterer": "cDBSCAN2",
         "l2": "256 MiB buffer written between timed steps; inputs (>= 80 MB per chromosome) exceed L2",
         "streams": "up to %s chromosomes of a rank in flight on separate CUDA streams (CLOOPS_STREAMS); stage and roofline times are taken with one" % os.environ.get("CLOOPS_STREAMS", "6")}
    if cfg["kind"] == "genome":
        c["sharding"] = "23 hg38-proportional chromosomes packed onto %d rank(s) by PET count (LPT); one NCCL all-reduce of the cut-off statistics per round" % world
    elif cfg["kind"] == "single":
        c["sharding"] = "%d chromosome(s) of %d PETs, one per GPU (replicas; a single chromosome does not shard)" % (world, args.pets)
    return c


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    sample, what = reference_sample(cfg, args, cores)
    run = make_cpu_run(cfg, sample, cores if cfg["kind"] == "genome" else 1)
    for _ in range(args.warmup):
        run.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loops = run.step()
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    run.close()
    value = run.pets / dt
    used = min(run.cores, len(sample)) if cfg["kind"] == "genome" else 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "PETs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if cfg["kind"] == "genome" else "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(cfg, args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "PETs/s", "cores": used, "kind": run.kind,
                         "sample": "%s: %d PET-clusterings per step (%s); PETs/s of the sample stands for the full workload (the reference's per-PET "
                                   "cost grows with chromosome size, so this favours the reference)" %
                                   (what, run.pets, "clustering only, 4 corner (eps, minPts) pairs" if cfg["kind"] == "sweep" else "all %d rounds + scoring" % (len(cfg["eps"]) * len(cfg["minPts"])))},
        "e2e": {"value": value, "unit": "PETs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "result": {"loops_in_sample": loops},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument("--pets", type=int, default=0, help="total PETs (default: the configuration's own size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.pets <= 0:
        args.pets = cfg["pets"]
    args.warmup = max(args.warmup, 3)                      # both arms: at least three untimed steps, the same K timed ones
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from cloops_b200 import _lib, device, dist, pipe, synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_from_env("nccl")
        import torch.distributed as td
    L = _lib.lib()
    dev = torch.device("cuda", local)
    quiet_logs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Per-step CUDA events on the launching stream, ranks aligned and L2 flushed before each step -> ms list."""
        out = []
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
        return out

    res, extra = {}, {}
    # ---------------------------------------------------------------------------------------------------- workloads
    if cfg["kind"] == "genome":
        counts = synth.genome_counts(args.pets)
        owner = dist.assign(list(range(23)), weights=counts, nranks=world)
        mine = [ci for ci in range(23) if owner[ci] == rank]
        chroms = generate(mine, lambda ci: synth.genome_chrom(args.pets, args.config, ci))
        host = {name: (torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()) for name, X, Y in chroms}
        cfs = ["mem:%s-%s.jd" % (n, n) for n in synth.CHROMS]
        n_mine = sum(len(c[1]) for c in chroms)
        total = sum(counts)

        def upload():
            for name, X, Y in chroms:
                hx, hy = host[name]
                pipe._Resident.register(name, X, Y, hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True))

        upload()

        def step_dev():
            res["r"] = pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], cfg["hic"], weights=counts, tail=False)

        def step_host():
            upload()
            res["h"] = pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], cfg["hic"], weights=counts, tail=True)

        h2d = 8 * total
        lpt = max(sum(counts[ci] for ci in range(23) if owner[ci] == r) for r in range(world))
        extra["lpt_ideal_speedup"] = round(total / lpt, 3)
        scaling = "strong"
        sample_have = {ci: c for ci, c in zip(mine, chroms)}
    elif cfg["kind"] == "single":
        X, Y = synth.chromosome(args.pets, CHROM_LEN_SINGLE, 20240 + args.config * 100 + rank, loop_frac=0.08, sigma=500.0)
        hx, hy = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()
        name = "chr%d" % (rank + 1)
        cfs = ["mem:%s-%s.jd" % (name, name)]
        total = args.pets * world

        def upload():
            pipe._Resident.register(name, X, Y, hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True))

        upload()
        # every rank runs its own single-chromosome pipeline: no collective (a lone chromosome does not shard)
        saved_world = dist._state["world"]

        def solo(fn):
            def run():
                dist._state["world"] = 1
                try:
                    fn()
                finally:
                    dist._state["world"] = saved_world
            return run

        def step_dev():
            res["r"] = pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], cfg["hic"], tail=False)

        def step_host():
            upload()
            res["h"] = pipe.call_loops(cfs, cfg["eps"], cfg["minPts"], cfg["hic"], tail=True)

        step_dev, step_host = solo(step_dev), solo(step_host)
        h2d = 8 * total
        scaling = "weak"
        sample_have = (X, Y)
    else:                                                    # config 5: eps x minPts sweep with index re-use across minPts
        X, Y = synth.chromosome(args.pets, CHROM_LEN_SINGLE, 20240 + args.config * 100 + rank, loop_frac=0.08, sigma=500.0)
        hx, hy = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()
        dxy = [hx.to(dev), hy.to(dev)]
        total = args.pets * world * len(cfg["eps"]) * len(cfg["minPts"])

        def sweep(dx, dy):
            out = []
            for ep in cfg["eps"]:
                ix = device.Index(dx, dy, ep)
                for m in cfg["minPts"]:
                    _, ls, info = ix.dbscan(m, _lib.V2, want_sorted=True, want_rows=False)
                    if device.Profile.on:
                        device.Profile.add_stages()
                        device.Profile.rq_bytes += 12 * info["n_active"]
                    out.append((ep, m, info["n_clusters"]))
                ix.close()
            torch.cuda.current_stream().synchronize()
            return out

        def step_dev():
            res["r"] = sweep(dxy[0], dxy[1])

        def step_host():
            res["h"] = sweep(hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True))

        h2d = 8 * args.pets * world
        scaling = "weak"
        sample_have = (X, Y)

    # ---------------------------------------------------------------------------------------------------- timing
    L.cloops_set_profiling(0)
    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(visible_index(local))
    sampler.start()
    launches0 = L.cloops_kernel_launches()
    ms = timed(step_dev, args.steps)
    launches = (L.cloops_kernel_launches() - launches0) / max(1, args.steps)
    barrier()
    clocks = sampler.stop()
    # stage times of the library's kernels, live, on the same stream, in separate steps (the event synchronisations of the
    # profiling mode would perturb `value`)
    streams_saved, pipe.STREAMS = pipe.STREAMS, 1          # one chromosome at a time: a kernel's events then bracket that kernel alone
    step_dev()                                             # untimed: this thread's scratch workspace grows to its steady size
    device.Profile.begin()
    flush.fill_(1)
    step_dev()
    stages = device.Profile.end()
    pipe.STREAMS = streams_saved
    rq_bytes, rc_bytes = device.Profile.rq_bytes, device.Profile.rc_bytes
    barrier()
    # end to end from pinned host buffers to the loop table
    step_host()
    device.Profile.d2h_bytes = 0
    n_e2e = max(2, args.steps // 2)
    ms_e2e = timed(step_host, n_e2e)
    d2h = device.Profile.d2h_bytes / n_e2e
    barrier()

    t_dev, t_e2e = float(np.mean(ms)), float(np.mean(ms_e2e))
    if world > 1:
        t = torch.tensor([t_dev, t_e2e, -t_dev], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        t_dev, t_e2e, t_min = float(t[0]), float(t[1]), -float(t[2])
        s = torch.tensor([launches, d2h, rq_bytes, rc_bytes, stages.get("region_query", 0.0), stages.get("range_counts", 0.0) + stages.get("region_pets", 0.0)],
                         dtype=torch.float64, device=dev)
        td.all_reduce(s, op=td.ReduceOp.SUM)
        launches, d2h, rq_bytes, rc_bytes, t_rq, t_rc = (float(v) for v in s)
        td.barrier()
    else:
        t_min = t_dev
        t_rq = stages.get("region_query", 0.0)
        t_rc = stages.get("range_counts", 0.0) + stages.get("region_pets", 0.0)
    if rank != 0:
        dist.shutdown()
        return
    peak, peak_kind = load_peaks()
    achieved = rq_bytes / (t_rq * 1e-3) / 1e9 if t_rq > 0 else 0.0
    achieved_rc = rc_bytes / (t_rc * 1e-3) / 1e9 if t_rc > 0 else 0.0
    line = {
        "metric": METRIC, "value": total / (t_dev * 1e-3), "unit": "PETs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(cfg, args, world),
        "e2e": {"value": total / (t_e2e * 1e-3), "unit": "PETs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e,
                "what": "pinned host coordinates -> marked loop table on rank 0 (H2D, all rounds, range counts, D2H, scipy tail, de-duplication, table gather)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "count_kernel_tiled (region query), all launches of one step", "achieved": achieved, "peak": peak,
                     "peak_source": peak_kind, "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic("config%d" % args.config),
                     "ms": t_rq, "algorithmic_bytes": int(rq_bytes), "frac_of_8TBs_nominal": achieved / 8000.0},
        "roofline_range_count": {"bound": "hbm", "kernel": "range_count_kernel<5> + <0> (cModel.py:118-143), all launches of one step", "achieved": achieved_rc,
                                 "peak": peak, "unit": "GB/s", "frac": achieved_rc / peak, "ms": t_rc, "algorithmic_bytes": int(rc_bytes),
                                 "definition": "8 B x (PETs with X in the hull of a candidate's windows + PETs with Y in it) + 4 B per output integer (SURVEY 8d)"},
        "stages_ms": {k: round(v, 3) for k, v in sorted(stages.items(), key=lambda kv: -kv[1])},
        "rank_balance": {"slowest_ms": t_dev, "fastest_ms": t_min},
        "step_ms": {"value": [round(v, 1) for v in ms], "e2e": [round(v, 1) for v in ms_e2e], "of": "rank 0"},
    }
    host_ms = res.get("h", {}).get("host_ms") if isinstance(res.get("h"), dict) else None
    if host_ms:                                            # where the host thread of the last e2e step spent its wall-clock (after the H2D copy)
        line["e2e"]["host_ms_last_step"] = {k: round(v, 1) for k, v in host_ms.items()}
    line.update(extra)
    if cfg["kind"] != "sweep":
        r, h = res["r"], res["h"]
        line["result"] = {"cut": r["cut"], "candidates_rank0": int(sum(len(v["records"]) for v in r["dataI"].values())),
                          "scored_rank0": int(sum(len(c["keep"]) for c in r["counted"].values() if c is not None)),
                          "loops": None if h["table"] is None else int(len(h["table"])),
                          "significant": None if h["table"] is None else int(h["table"]["significant"].sum())}
    else:
        line["result"] = {"sweep": res["r"][:4]}
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        sample, what = reference_sample(cfg, args, cores, have=sample_have)
        run = make_cpu_run(cfg, sample, cores if cfg["kind"] == "genome" else 1)
        t0 = time.perf_counter()
        run.step()
        dt = time.perf_counter() - t0
        run.close()
        line["cpu_baseline"] = {"value": run.pets / dt, "unit": "PETs/s", "cores": min(run.cores, len(sample)), "kind": run.kind,
                                "sample": "%s: %d PET-clusterings (%s), %.1f s" % (what, run.pets, "clustering only, 4 corner (eps, minPts) pairs" if cfg["kind"] == "sweep" else "all rounds + scoring", dt)}
    print(json.dumps(line), flush=True)
    dist.shutdown()


if __name__ == "__main__":
    main()
