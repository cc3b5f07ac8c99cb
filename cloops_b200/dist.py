"""Multi-GPU plumbing: one process per GPU (torchrun), chromosomes sharded across ranks.

The reference parallelises with ``joblib.Parallel`` over chromosome files (cLoops/pipe.py:117,184);
chromosomes are independent inside a clustering round and inside scoring, so the data path needs no
collective.  Two exchanges remain, both host-object gathers over ``torch.distributed`` (NCCL on GPUs,
gloo on CPU-only hosts): the per-round candidate records + distance lists that feed the pooled
cut-off estimate (pipe.py:120-127,259) and the final per-chromosome loop tables (pipe.py:187-191).
"""
from __future__ import annotations

import os

_state = {"rank": 0, "world": 1, "init": False}


def init_from_env(backend: str | None = None) -> None:
    """Join the torchrun job described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*; no-op for one process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or _state["init"]:
        return
    import torch
    import torch.distributed as td
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kwargs = {}
    if torch.cuda.is_available():
        torch.cuda.set_device(local % torch.cuda.device_count())
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local % torch.cuda.device_count())
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if not td.is_initialized():
        td.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    _state.update(rank=rank, world=world, init=True)


def rank() -> int:
    return _state["rank"]


def world() -> int:
    return _state["world"]


def assign(items, weights=None, nranks: int | None = None):
    """Longest-processing-time packing of ``items`` onto ranks.  Default weight of a path is its file
    size (a .jd holds 24 B per PET), so every rank derives the same deterministic assignment."""
    nranks = world() if nranks is None else nranks
    items = list(items)
    if weights is None:
        weights = [os.path.getsize(i) if isinstance(i, str) and os.path.exists(i) else 1 for i in items]
    order = sorted(range(len(items)), key=lambda k: (-weights[k], k))
    load = [0] * nranks
    owner = [0] * len(items)
    for k in order:
        r = min(range(nranks), key=lambda q: (load[q], q))
        owner[k] = r
        load[r] += weights[k]
    return owner


def my_share(items, weights=None):
    items = list(items)
    if world() == 1:
        return items
    owner = assign(items, weights)
    return [it for it, o in zip(items, owner) if o == rank()]


def merge_in_order(items, part: dict):
    """All ranks contribute ``{item: result}``; every rank gets ``[result for item in items]``."""
    if world() == 1:
        return [part[i] for i in items]
    import torch.distributed as td
    gathered = [None] * world()
    td.all_gather_object(gathered, part)
    merged = {}
    for g in gathered:
        merged.update(g)
    return [merged[i] for i in items]


def all_gather_concat(t):
    """Concatenation over ranks (rank order) of 1-D tensors of different lengths; the result lives on
    the device of ``t``.  Tensor collective (NCCL for CUDA tensors, gloo for CPU tensors)."""
    if world() == 1:
        return t
    import torch
    import torch.distributed as td
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world())]
    td.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros(cap, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    bufs = [torch.empty_like(pad) for _ in range(world())]
    td.all_gather(bufs, pad)
    return torch.cat([b[:k] for b, k in zip(bufs, sizes)])


def all_reduce_sum(t):
    """In-place sum over ranks of a tensor (NCCL for CUDA tensors, gloo for CPU tensors)."""
    if world() > 1:
        import torch.distributed as td
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def all_gather_objects(obj):
    """-> [object of rank 0, ..., object of rank world-1] on every rank."""
    if world() == 1:
        return [obj]
    import torch.distributed as td
    out = [None] * world()
    td.all_gather_object(out, obj)
    return out


def broadcast_object(obj, src: int = 0):
    if world() == 1:
        return obj
    import torch.distributed as td
    box = [obj]
    td.broadcast_object_list(box, src=src)
    return box[0]


def barrier() -> None:
    if world() > 1:
        import torch.distributed as td
        td.barrier()


def shutdown() -> None:
    """Leave the process group (quietens NCCL's exit warning)."""
    if _state["init"]:
        import torch.distributed as td
        if td.is_initialized():
            td.destroy_process_group()
        _state.update(rank=0, world=1, init=False)
