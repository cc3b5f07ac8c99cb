#!/usr/bin/env python
"""Benchmark of the cLoops hot path on B200 (BASELINE.json: "PETs/sec clustered+scored").

A step = one pass of the hot path over one chromosome: cDBSCAN2 clustering (eps-neighbourhood region
query, core-graph components, border ownership, survival, numbering) -> per-cluster candidate records
-> coverage model -> permuted-background range counts (123 integers) of every inter-ligation
candidate.  Workload at N GPUs: N synthetic "ChIA-PET 10M cis PETs, single chromosome, eps=1000,
minPts=5" chromosomes (BASELINE.json configs[1]), one per rank (weak scaling; chromosomes are
independent, no data-path collective).  The scipy p-value tail on the host is not part of the step on
either arm.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pets P]

Prints ONE JSON line (rank 0).  `value` = PETs of all ranks / max-over-ranks device time with inputs
resident in HBM; `e2e` = same pass from pinned host buffers with H2D and D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EPS, MINPTS = 1000, 5
CHROM_LEN = 249_000_000


def load_traffic(n_pets):
    """DRAM bytes per launch of the region-query kernel from the committed ncu capture (same workload)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_count_kernel_metrics.json")) as fh:
            m = json.load(fh)
        return float(m["traffic_bytes"]) if int(m["n_pets"]) == int(n_pets) else None
    except Exception:
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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
            time.sleep(0.004)

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


def cpu_sample(X, Y, frac):
    """Same-density sample: the PETs whose left anchor lies in the first `frac` of the chromosome."""
    m = X < int(CHROM_LEN * frac)
    return X[m].astype(np.int64), Y[m].astype(np.int64)


def run_reference(args):
    """Reference arm: the CPU restatement of the reference's algorithm (oracle port; the reference
    itself is Python 2 and cannot run on this image) on a bounded same-density sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cloops_b200 import synth
    from oracle import spec
    n_chrom = max(1, args.gpus)
    frac = 0.03
    samples = []
    for c in range(n_chrom):
        X, Y = synth.config2(args.pets, seed=20240 + 200 + c)
        samples.append(cpu_sample(X, Y, frac))
    cores = min(n_chrom, os.cpu_count() or 1)
    pets = sum(len(s[0]) for s in samples)

    def one_step():
        if cores == 1:
            for xs, ys in samples:
                spec.hot_path_cpu(xs, ys, EPS, MINPTS)
        else:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(cores) as pool:
                pool.starmap(spec.hot_path_cpu, [(xs, ys, EPS, MINPTS) for xs, ys in samples])

    for _ in range(min(args.warmup, 1)):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = (time.perf_counter() - t0) / args.steps
    value = pets / dt
    sample = "per step: PETs with X < %.0f%% of each %d-PET chromosome (same density), %d PETs total" % (frac * 100, args.pets, pets)
    line = {
        "impl": "reference", "metric": "PETs/sec clustered+scored", "value": value, "unit": "PETs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, n_chrom),
        "cpu_baseline": {"value": value, "unit": "PETs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "PETs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_chrom):
    return {"workload": "synthetic ChIA-PET %d cis PETs, single chromosome, eps=%d minPts=%d (BASELINE.json configs[1]); "
                        "%d chromosome(s), one per GPU" % (args.pets, EPS, MINPTS, n_chrom),
            "clusterer": "cDBSCAN2", "scoring": "range counts (123 ints) of every inter-ligation candidate; scipy tail excluded",
            "outputs": "candidate records, per-PET inter/self membership (index order), range counts (the returns of pipe.py:52-110 + cModel.py:118-143)",
            "l2": "256 MiB buffer written between timed steps", "pets_per_gpu": args.pets}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pets", type=int, default=10_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from cloops_b200 import _lib, dist, hotpath, synth
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

    X, Y = synth.config2(args.pets, seed=20240 + 200 + rank)
    hx, hy = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()
    dx, dy = hx.to(dev), hy.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n = args.pets

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Per-step CUDA events on the launching stream, L2 flushed between steps; returns ms list."""
        out = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
        return out

    res = {}

    def step_dev():
        res["r"] = hotpath.run_device(dx, dy, EPS, MINPTS)

    host = hotpath.HostStep(n)

    def step_host():
        res["h"] = host(hx, hy, EPS, MINPTS)

    L.cloops_set_profiling(0)
    for _ in range(args.warmup):
        step_dev()
    step_host()
    # ---- device-resident timing (value) + live stage timing of the region-query kernel
    L.cloops_set_profiling(1)
    rq_ms, stage_tot = [], {}

    barrier()
    sampler = ClockSampler(visible_index(local))
    sampler.start()
    launches0 = L.cloops_kernel_launches()
    L.cloops_set_profiling(0)
    ms = timed(step_dev, args.steps)
    launches = L.cloops_kernel_launches() - launches0
    barrier()
    clocks = sampler.stop()
    # region-query kernel, live, profiling events on the same stream (separate passes so that the
    # event syncs do not perturb `value`)
    L.cloops_set_profiling(1)
    from cloops_b200 import device
    for _ in range(max(3, args.steps // 2)):
        flush.fill_(1)
        device.dbscan_device(dx, dy, EPS, MINPTS, _lib.V2)
        st = _lib.stage_times()
        rq_ms.append(st.get("region_query", 0.0))
        for k, v in st.items():
            stage_tot[k] = stage_tot.get(k, 0.0) + v
    L.cloops_set_profiling(0)
    stage_avg = {k: round(v / len(rq_ms), 4) for k, v in stage_tot.items()}
    # ---- end-to-end from pinned host buffers
    barrier()
    ms_e2e = timed(step_host, max(3, args.steps // 2))
    barrier()

    t_dev = float(np.mean(ms))
    t_e2e = float(np.mean(ms_e2e))
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
    if world > 1:
        td.barrier()
    if rank != 0:
        dist.shutdown()
        return
    r = res["r"]
    n_act = r.info["n_active"]
    peak, peak_kind = load_peaks()
    t_rq = float(np.mean(rq_ms))
    achieved = 12.0 * n_act / (t_rq * 1e-3) / 1e9 if t_rq > 0 else 0.0
    total = n * world
    line = {
        "metric": "PETs/sec clustered+scored", "value": total / (t_dev * 1e-3), "unit": "PETs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": total / (t_e2e * 1e-3), "unit": "PETs/s", "h2d_bytes_per_step": host.h2d_bytes,
                "d2h_bytes_per_step": int(host.d2h_bytes), "ms_per_step": t_e2e},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "count_kernel_tiled (region query)", "achieved": achieved, "peak": peak,
                     "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6650 GB/s",
                     "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic(n_act), "ms": t_rq,
                     "algorithmic_bytes": 12 * n_act, "frac_of_8TBs_nominal": achieved / 8000.0},
        "stages_ms": stage_avg,
        "result": {"clusters": r.info["n_clusters"], "core": r.info["n_core"], "dead": r.info["n_dead"],
                   "labelled": r.info["n_labelled"], "inter_candidates": int(r.cand.shape[0])},
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import spec
        xs, ys = cpu_sample(X, Y, 0.12)
        t0 = time.perf_counter()
        spec.hot_path_cpu(xs, ys, EPS, MINPTS)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": len(xs) / dt, "unit": "PETs/s", "cores": 1, "kind": "port",
                                "sample": "PETs with X < 12%% of the chromosome (same density): %d PETs, %.1f s" % (len(xs), dt)}
    print(json.dumps(line), flush=True)
    dist.shutdown()


if __name__ == "__main__":
    main()
