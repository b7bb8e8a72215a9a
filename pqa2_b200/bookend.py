"""White "bookend" frame detection on the GPU -- the per-pixel scan of the alignment stage that feeds
the VMAF path (reference ``app/bookend_alignment.py:755-1134``,
``app/reference_analyzer.py:112-152``).

The reference decodes with cv2 and computes, per sampled frame, ``np.mean(gray)``, ``np.std(gray)`` and
``np.sum(gray > threshold)`` on the CPU.  Here the luma planes go through ``bv_luma_stats_device`` (one
streaming pass: exact integer sums of y, y^2 and three threshold counts per frame), and the decision
logic below restates the reference's thresholds and frame test on those statistics.  Raw planar input
carries Y directly, so ``gray`` is the luma plane (the reference's BGR->gray of a decoded frame is the
same quantity up to the decoder's range conversion).

Not rebuilt: the ffmpeg trim / re-encode that follows detection (``:530-602``; row f4)."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .extractor import DeviceBuffer


@dataclass
class FrameStat:
    mean: float
    std: float
    ratios: tuple          # share of pixels above each of the 3 thresholds


def luma_stats_host(frames, bpc: int, thresholds) -> list:
    """numpy statement of the same statistics (used by the tests as the checker)."""
    out = []
    for y in frames:
        v = y.astype(np.float64)
        out.append(FrameStat(float(v.mean()), float(v.std()), tuple(float((y > t).mean()) for t in thresholds)))
    return out


def luma_stats(frames, bpc: int = 8, thresholds=(230, 207, 184), device: int = 0, chunk: int = 64) -> list:
    """Per-frame mean / population std / share of pixels > thresholds[k] of luma planes, on the GPU.

    ``frames``: iterable of C-contiguous 2-D uint8 (bpc 8) or uint16 arrays of one size."""
    lib = L.load()
    frames = list(frames)
    if not frames:
        return []
    h, w = frames[0].shape
    bps = 1 if bpc == 8 else 2
    plane = w * h * bps
    thr = (C.c_uint * 3)(*[int(math.floor(t)) for t in thresholds])
    out: list = []
    buf = DeviceBuffer(plane * min(chunk, len(frames)), device)
    try:
        for c0 in range(0, len(frames), chunk):
            part = frames[c0:c0 + chunk]
            for k, y in enumerate(part):
                if y.shape != (h, w):
                    raise ValueError("all frames must have the same size")
                buf.upload(k * plane, np.ascontiguousarray(y))
            raw = (C.c_uint64 * (5 * len(part)))()
            rc = lib.bv_luma_stats_device(device, buf.ptr, w * bps, plane, len(part), w, h, bpc, thr, raw)
            if rc != 0:
                raise RuntimeError(f"bv_luma_stats_device failed ({rc}); the engine has no CPU fallback")
            n = float(w * h)
            for k in range(len(part)):
                s1, s2, a, b, c = (int(raw[5 * k + j]) for j in range(5))
                mean = s1 / n
                var = max(s2 / n - mean * mean, 0.0)
                out.append(FrameStat(mean, math.sqrt(var), (a / n, b / n, c / n)))
    finally:
        buf.free()
    return out


def brightness_thresholds(means, stds, adaptive: bool = True, white_threshold: float = 230.0):
    """The three whiteness thresholds and the std-dev threshold (reference :806-852, :870)."""
    avg_b, std_b, max_b = float(np.mean(means)), float(np.std(means)), float(np.max(means))
    avg_sd = float(np.mean(stds))
    if adaptive:
        dyn = max(avg_b + 2.0 * std_b, max_b * 0.85, 180)
        if max_b > 240:
            dyn = max(dyn, 220)
        elif max_b < 200:
            dyn = max(avg_b + 1.5 * std_b, 160)
        thr = [dyn, dyn * 0.9, max(avg_b + 20, 160)]
    else:
        thr = [white_threshold, white_threshold * 0.9, white_threshold * 0.8]
    return thr, min(45.0, avg_sd * 1.8)


def is_white_frame(mean: float, std: float, white_ratio: float, threshold: float, std_dev_threshold: float) -> bool:
    """The detailed-pass frame test (reference :1003-1020); white_ratio = share of pixels > threshold."""
    if std < std_dev_threshold * 1.2:
        return mean > threshold * 0.95
    if mean > threshold:
        return True
    if mean > threshold * 0.9:
        return white_ratio > 0.7
    return False


def begins_with_bookend(stats_200: list, max_frames: int = 30) -> bool:
    """reference_analyzer._check_for_bookends (:112-152): any of the first 30 frames with > 85 % of its
    pixels above 200.  stats_200: FrameStat list computed with 200 as the first threshold."""
    return any(s.ratios[0] > 0.85 for s in stats_200[:max_frames])


def detect_white_bookends(frames, fps: float, bpc: int = 8, adaptive: bool = True, white_threshold: float = 230.0,
                          device: int = 0) -> list:
    """Runs of white frames [{start_frame, end_frame, start_time, end_time, frame_count}] over a clip.

    Two GPU passes over the luma planes: one for mean/std (threshold selection, reference :774-852), one
    with the selected thresholds for the white-pixel shares; every frame is tested (the reference samples
    and then re-reads regions of interest because its scan is CPU-bound, :873-948)."""
    frames = list(frames)
    sc = 1 << (bpc - 8)
    first = luma_stats(frames, bpc, (200 * sc, 200 * sc, 200 * sc), device)
    means = [s.mean / sc for s in first]
    stds = [s.std / sc for s in first]
    thr, sd_thr = brightness_thresholds(means, stds, adaptive, white_threshold)
    second = luma_stats(frames, bpc, tuple(t * sc for t in thr), device)
    min_white = max(3, int(0.1 * fps)) if fps > 25 else 3
    for k, threshold in enumerate(thr):                       # strictest threshold first, then the fallbacks
        runs, cur = [], None
        for i, s in enumerate(second):
            white = is_white_frame(means[i], stds[i], s.ratios[k], threshold, sd_thr)
            if white and cur is None:
                cur = i
            if (not white or i == len(second) - 1) and cur is not None:
                end = i if white else i - 1
                if end - cur + 1 >= min_white:
                    runs.append({"start_frame": cur, "end_frame": end, "start_time": cur / fps, "end_time": end / fps,
                                 "frame_count": end - cur + 1, "threshold": threshold})
                cur = None
        if len(runs) >= 2:                                    # a loop needs an opening and a closing bookend
            return runs
    return runs
