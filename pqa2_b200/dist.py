"""One-process-per-GPU mode (torchrun): frame shards per rank, host gather of the per-frame rows.

The path has no exchange step: rank r scores the contiguous chunk
``shard_ranges(n, world)[r]`` with a one-frame lead-in that only feeds the motion state, and the
per-frame feature rows (a few hundred bytes each) are gathered on rank 0, which applies libvmaf's
motion2 rule across the shard boundaries, the SVR and the pooling.  ``torch.distributed`` is used
for that gather and for the max-over-ranks timing only -- NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import threading

from . import engine, report


def rank_range(n_frames: int, rank: int, world: int, first: int = 0, weights=None):
    """[start, end) of this rank's chunk (may be empty when world > n_frames)."""
    r = engine.shard_ranges(n_frames, world, weights)
    if rank >= len(r):
        return (first + n_frames, first + n_frames)
    return (first + r[rank][0], first + r[rank][1])


def _default_shard_fn(src, model, opt, device, start, end, mask, session=None):
    rows = engine.Rows(src.nb_frames)
    errors, holders = [], []
    engine._run_shard(src, model, opt, device, start, end, mask, rows, None, threading.Event(), errors, holders,
                      session)
    for kind, e in errors:
        if kind == "error":
            raise e
    return ("block", start, rows.arr[start:end].copy())        # one structured array per shard (~1.3 KB/frame)


def _world(group=None):
    import sys
    dist = sys.modules.get("torch.distributed")       # never import torch here: a process group can only exist if the
    if dist is not None and dist.is_available() and dist.is_initialized():      # caller already did (and the import costs seconds)
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def max_over_ranks(value: float, group=None, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def calibrate(src, model, opt: engine.EngineOptions | None = None, device: int = 0, group=None, frames: int = 0,
              session: "engine.Engine | None" = None) -> list:
    """Per-rank frames/s of a short trial, all ranks at once: the weights for ``analyze_distributed(weights=...)``.

    Every rank scores the same first ``frames`` frames of ``src`` (default: four launch groups) on its own GPU while all
    the others do the same, so the number reflects what the rank gets when the host's PCIe links are shared -- on the
    8 x B200 box four GPUs hang off one host bridge and see ~30 GB/s each under load, against 50 GB/s alone
    (tools/h2d_concurrent.py).  Call on every rank; returns the same list everywhere."""
    import time
    opt = opt or engine.EngineOptions()
    rank, world = _world(group)
    if world == 1:
        return [1.0]
    import torch
    import torch.distributed as dist
    n = min(src.nb_frames, frames or 64)
    mask = engine.feature_mask(model, opt)
    _default_shard_fn(src, model, opt, device, 0, min(n, 8), mask, session)          # contexts, first launches
    dist.barrier(group)
    t0 = time.perf_counter()
    _default_shard_fn(src, model, opt, device, 0, n, mask, session)
    fps = n / (time.perf_counter() - t0)
    nccl = dist.get_backend(group) == "nccl"
    t = torch.tensor([fps], dtype=torch.float64, device=f"cuda:{device}" if nccl else "cpu")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [float(x.item()) for x in out]


def analyze_distributed(src, model, opt: engine.EngineOptions | None = None, device: int = 0, group=None,
                        shard_fn=None, svr_device=None, session: "engine.Engine | None" = None, weights=None):
    """Call on every rank.  Returns the libvmaf log dict on rank 0 and None elsewhere.

    ``shard_fn(src, model, opt, device, start, end, mask) -> {frame_index: row}`` computes one shard
    (default: the CUDA extractors on ``device``); the CPU tests inject a stub.  ``session`` keeps this rank's
    CUDA context alive between clips.  ``weights`` (from ``calibrate``; identical on every rank) sizes the chunks by each
    rank's sustained rate.  Without an initialised process group the call is the world-size-1 case."""
    opt = opt or engine.EngineOptions()
    rank, world = _world(group)
    n = src.nb_frames
    start, end = rank_range(n, rank, world, weights=weights)
    mask = engine.feature_mask(model, opt)
    failure = None
    try:
        if end <= start:
            mine = {}
        elif shard_fn is not None:
            mine = shard_fn(src, model, opt, device, start, end, mask)
        else:
            mine = _default_shard_fn(src, model, opt, device, start, end, mask, session)
    except Exception as e:                # noqa: BLE001
        # every rank must still reach the gather (a rank that raised before it would leave the others waiting for ever):
        # the failure travels as this rank's part, rank 0 reports it, this rank re-raises its own exception afterwards
        failure, mine = e, ("failed", rank, f"{type(e).__name__}: {e}")
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0, group=group)
    else:
        gathered = [mine]
    if failure is not None:
        raise failure
    if rank != 0:
        return None
    bad = [p for p in gathered if isinstance(p, tuple) and p and p[0] == "failed"]
    if bad:
        raise RuntimeError("; ".join(f"rank {r}: {msg}" for _, r, msg in bad))
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
    pooled_out: dict = {}
    frames = engine.build_frames(rows, model, opt, svr_device, pooled_out)
    return {"version": report.VERSION, "frames": frames,
            "pooled_metrics": pooled_out.get("pooled") or report.pooled_metrics(frames),
            "aggregate_metrics": {}, "rows": rows, "model": model.name, "n_frames": n, "world_size": world}


def analyze_batch_distributed(clips: list, model, opt: engine.EngineOptions | None = None, device: int = 0, group=None,
                              session: "engine.Engine | None" = None, summarize=None, concurrency: int = 3):
    """Many clips over the ranks (BASELINE.json configs[4]; the one-process-per-GPU twin of engine.analyze_batch):
    whole clips are the unit, clip k runs on rank ``k % world``, and one small summary per clip -- by default its
    pooled report -- is gathered on rank 0, in input order.  No lead-in frames, no collective on the data path.

    ``concurrency`` clips are in flight on this rank's GPU at once, each through its own session (own CUDA context,
    streams and pinned ring): a clip's pipeline fill and its drain + read-back + scoring leave the GPU partly idle for a
    few ms, which for clips of a few hundred frames is 10-15 % of the clip; the other clips fill those gaps (one B200,
    300-frame 1080p clips: 9.3k / 11.3k / 11.5k / 11.6k fps with 1 / 2 / 3 / 4 in flight).  ``session``
    may be one Engine or a list of them (one per worker) that outlive the call; missing ones are created and closed here.
    Returns the list on rank 0, None elsewhere; a failed clip yields ``{"error": str}``."""
    from dataclasses import replace
    opt = replace(opt or engine.EngineOptions(), devices=(device,))
    rank, world = _world(group)
    summarize = summarize or (lambda log: {k: v for k, v in log.items() if k not in ("frames", "rows")})
    todo = list(range(rank, len(clips), world))
    mine: dict = {}
    lock = threading.Lock()

    def worker(sess, own):
        try:
            while True:
                with lock:
                    if not todo:
                        return
                    k = todo.pop(0)
                try:
                    out = summarize(sess.analyze(clips[k], model, opt))
                except Exception as e:            # noqa: BLE001  (the reference's per-clip error convention)
                    out = {"error": str(e)}
                with lock:
                    mine[k] = out
        finally:
            if own:
                sess.close()

    given = list(session) if isinstance(session, (list, tuple)) else ([session] if session is not None else [])
    n_workers = max(1, min(int(concurrency), max(len(todo), 1)))
    sessions = [(given[w], False) if w < len(given) else (engine.Engine(), True) for w in range(n_workers)]
    threads = [threading.Thread(target=worker, args=sw, daemon=True) for sw in sessions[1:]]
    for t in threads:
        t.start()
    worker(*sessions[0])
    for t in threads:
        t.join()
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0, group=group)
    else:
        gathered = [mine]
    if rank != 0:
        return None
    out = [None] * len(clips)
    for part in gathered:
        for k, v in part.items():
            out[k] = v
    return out
