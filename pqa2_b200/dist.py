"""One-process-per-GPU mode (torchrun): frame shards per rank, host gather of the per-frame rows.

The path has no exchange step: rank r scores the contiguous chunk
``shard_ranges(n, world)[r]`` with a one-frame lead-in that only feeds the motion state, and the
per-frame feature rows (a few hundred bytes each) are gathered on rank 0, which applies libvmaf's
motion2 rule across the shard boundaries, the SVR and the pooling.  ``torch.distributed`` is used
for that gather and for the max-over-ranks timing only -- NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import threading

from . import engine, report


def rank_range(n_frames: int, rank: int, world: int, first: int = 0):
    """[start, end) of this rank's chunk (may be empty when world > n_frames)."""
    r = engine.shard_ranges(n_frames, world)
    if rank >= len(r):
        return (first + n_frames, first + n_frames)
    return (first + r[rank][0], first + r[rank][1])


def _default_shard_fn(src, model, opt, device, start, end, mask):
    rows = engine.Rows(src.nb_frames)
    errors, holders = [], []
    engine._run_shard(src, model, opt, device, start, end, mask, rows, None, threading.Event(), errors, holders)
    for kind, e in errors:
        if kind == "error":
            raise e
    return ("block", start, rows.arr[start:end].copy())        # one structured array per shard (~1.3 KB/frame)


def max_over_ranks(value: float, group=None, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def analyze_distributed(src, model, opt: engine.EngineOptions | None = None, device: int = 0, group=None,
                        shard_fn=None, svr_device=None):
    """Call on every rank.  Returns the libvmaf log dict on rank 0 and None elsewhere.

    ``shard_fn(src, model, opt, device, start, end, mask) -> {frame_index: row}`` computes one shard
    (default: the CUDA extractors on ``device``); the CPU tests inject a stub."""
    import torch.distributed as dist
    opt = opt or engine.EngineOptions()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = src.nb_frames
    start, end = rank_range(n, rank, world)
    mask = engine.feature_mask(model, opt)
    mine = (shard_fn or _default_shard_fn)(src, model, opt, device, start, end, mask) if end > start else {}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    if rank != 0:
        return None
    if all(isinstance(p, tuple) or not p for p in gathered):
        rows = engine.Rows(n)
        for part in gathered:
            if part:
                _, s0, arr = part
                rows.put(list(range(s0, s0 + len(arr))), arr)
        missing = [i for i in range(n) if not rows.present[i]]
    else:                                               # dict rows (stub shard functions in the CPU tests)
        rows = [None] * n
        for part in gathered:
            for i, r in part.items():
                rows[i] = r
        missing = [i for i, r in enumerate(rows) if r is None]
    if missing:
        raise RuntimeError(f"frames missing after the gather: {missing[:8]}...")
    frames = engine.build_frames(rows, model, opt, svr_device)
    return {"version": report.VERSION, "frames": frames, "pooled_metrics": report.pooled_metrics(frames),
            "aggregate_metrics": {}, "rows": rows, "model": model.name, "n_frames": n, "world_size": world}
