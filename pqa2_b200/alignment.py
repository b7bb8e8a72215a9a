"""Bookend alignment without the lossy re-encode.

The reference aligns a captured clip to its reference in three steps (``app/bookend_alignment.py``):
detect the white bookend runs (``:755-1134`` -- here ``pqa2_b200.bookend`` on the GPU), derive the content
window between two runs (``:333-389``), and then *re-encode* both clips with libx264 at CRF 23 so that the
captured one starts after the first bookend, runs at the reference frame rate and has exactly the
reference's frame count (``:530-602``).  Those two encodes perturb the very pixels VMAF is about to
measure.  This engine reads frames by index, so the same alignment is expressed as a **frame map**
(output frame k <- reference frame k, captured frame j_k) and scored directly; ``write_aligned_y4m``
materialises the aligned pair losslessly when files are wanted (``<base>_<stamp>_aligned.y4m`` instead of
the reference's ``_aligned.mp4``, ``:525-527``).

Timing rules restated from the reference's ffmpeg command lines:

* content window: ``[first.end_time + 1.5/cap_fps, last.start_time - 1.5/cap_fps]`` (``:336-338``); when the
  window is more than 1.5x the reference duration the capture holds several loops: with > 2 bookends take
  the consecutive pair whose gap is closest to the reference duration, else keep one reference duration
  from the start (``:353-389``);
* captured trim: ``-itsoffset frame_offset/cap_fps -i cap -ss content_start + 0.2 -r ref_fps -frames:v N``
  (``:548-583``, default ``frame_offset`` 6): input frame j carries pts ``j/cap_fps + offset``; frames with
  pts below ``content_start + 0.2`` are dropped; the rest is re-timed to the reference rate (nearest input
  frame per output tick, duplicates / drops as ffmpeg's constant-frame-rate sync does) and cut at the
  reference's frame count N;
* reference: the whole clip, every frame (``:530-537``).
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass, field

import numpy as np

from .engine import FrameSource

DEFAULT_FRAME_OFFSET = 6          # options_manager "bookend.frame_offset" default (reference :548)
START_PAD_S = 0.2                 # "ensure we start after any white frames" (reference :563-565)
BUFFER_FRAMES = 1.5               # reference :336


class AlignmentError(ValueError):
    pass


@dataclass
class ContentWindow:
    start_time: float
    end_time: float
    duration: float
    first_bookend: dict
    last_bookend: dict
    loop_index: int = 0            # which consecutive bookend pair was chosen (multi-loop captures)
    multi_loop: bool = False


def select_content(bookends: list, cap_fps: float, ref_duration: float) -> ContentWindow:
    """The content window between two white bookend runs (reference :318-389).

    ``bookends``: runs as ``bookend.detect_white_bookends`` returns them (start_time / end_time in seconds)."""
    if len(bookends) < 2:
        raise AlignmentError("Failed to detect at least two white bookends in the captured video")
    first, last = bookends[0], bookends[-1]
    buf = BUFFER_FRAMES / (cap_fps or 30.0)
    start, end = first["end_time"] + buf, last["start_time"] - buf
    if end <= start:
        raise AlignmentError("Invalid content timing between bookends")
    win = ContentWindow(start, end, end - start, first, last)
    if ref_duration > 0 and win.duration > ref_duration * 1.5:
        win.multi_loop = True
        if len(bookends) > 2:
            best, best_diff = 0, float("inf")
            for i in range(len(bookends) - 1):
                ls = bookends[i]["end_time"] + buf
                le = bookends[i + 1]["start_time"] - buf
                diff = abs((le - ls) - ref_duration)
                if diff < best_diff:
                    best, best_diff = i, diff
            a, b = bookends[best], bookends[best + 1]
            win.start_time = a["end_time"] + buf
            win.end_time = b["start_time"] - buf
            win.duration = win.end_time - win.start_time
            win.first_bookend, win.last_bookend, win.loop_index = a, b, best
        else:
            win.duration = ref_duration
            win.end_time = win.start_time + ref_duration
    return win


def captured_frame_map(content_start_time: float, n_out: int, ref_fps: float, cap_fps: float, cap_frames: int,
                       frame_offset: int = DEFAULT_FRAME_OFFSET, start_pad: float = START_PAD_S) -> np.ndarray:
    """Index of the captured frame shown at each of the ``n_out`` output ticks (reference :548-583).

    Shorter than ``n_out`` when the capture ends first (ffmpeg would stop there too; the reference then logs
    a frame-count mismatch and libvmaf scores the common prefix)."""
    if ref_fps <= 0 or cap_fps <= 0:
        raise AlignmentError("frame rates must be positive")
    offset_time = frame_offset / cap_fps
    t0 = content_start_time + start_pad - offset_time         # stream time of the first kept frame
    j0 = max(0, int(math.ceil(t0 * cap_fps - 1e-9)))
    k = np.arange(n_out, dtype=np.float64)
    j = j0 + np.floor(k * (cap_fps / ref_fps) + 0.5).astype(np.int64)
    return j[j < cap_frames]


@dataclass
class AlignmentPlan:
    ref_frames: np.ndarray                       # reference frame index per output frame (identity)
    cap_frames: np.ndarray                       # captured frame index per output frame
    window: ContentWindow
    ref_fps: float
    cap_fps: float
    frame_offset: int = DEFAULT_FRAME_OFFSET
    warnings: list = field(default_factory=list)

    @property
    def n_frames(self) -> int:
        return int(len(self.cap_frames))


def plan_alignment(ref_info, cap_info, bookends: list, frame_offset: int = DEFAULT_FRAME_OFFSET) -> AlignmentPlan:
    """Frame map for a (reference, captured) pair given the captured clip's bookend runs."""
    ref_fps, cap_fps = ref_info.fps or 30.0, cap_info.fps or 30.0
    ref_duration = ref_info.nb_frames / ref_fps
    win = select_content(bookends, cap_fps, ref_duration)
    cap = captured_frame_map(win.start_time, ref_info.nb_frames, ref_fps, cap_fps, cap_info.nb_frames, frame_offset)
    plan = AlignmentPlan(np.arange(len(cap), dtype=np.int64), cap, win, ref_fps, cap_fps, frame_offset)
    if len(cap) != ref_info.nb_frames:
        plan.warnings.append(f"Frame count mismatch: reference={ref_info.nb_frames}, captured={len(cap)}")
    if len(cap) == 0:
        raise AlignmentError("captured clip ends before the content window starts")
    return plan


class AlignedSource(FrameSource):
    """Frame source over two raw clips and a frame map: what the reference's ``*_aligned.mp4`` pair holds,
    minus the two CRF-23 encodes.  Output frame k = (reference frame ref_frames[k], captured frame
    cap_frames[k])."""

    def __init__(self, ref_info, cap_info, ref_frames, cap_frames):
        if (ref_info.width, ref_info.height, ref_info.bpc) != (cap_info.width, cap_info.height, cap_info.bpc):
            raise AlignmentError(
                f"geometry differs: reference {ref_info.width}x{ref_info.height}@{ref_info.bpc}b, "
                f"captured {cap_info.width}x{cap_info.height}@{cap_info.bpc}b")
        if len(ref_frames) != len(cap_frames):
            raise AlignmentError("frame maps differ in length")
        self._ri, self._ci = ref_info, cap_info
        self.ref_frames = np.asarray(ref_frames, dtype=np.int64)
        self.cap_frames = np.asarray(cap_frames, dtype=np.int64)
        self.width, self.height, self.bpc = ref_info.width, ref_info.height, ref_info.bpc
        self.chroma = ref_info.chroma if ref_info.chroma == cap_info.chroma else 400
        self.nb_frames = int(len(self.ref_frames))
        self.fps = ref_info.fps or 30.0
        self._r = self._c = None

    @classmethod
    def from_plan(cls, ref_info, cap_info, plan: AlignmentPlan) -> "AlignedSource":
        return cls(ref_info, cap_info, plan.ref_frames, plan.cap_frames)

    def open(self):
        from .yuvio import ClipReader
        s = AlignedSource(self._ri, self._ci, self.ref_frames, self.cap_frames)
        s._r, s._c = ClipReader(self._ri), ClipReader(self._ci)
        return s

    def close(self):
        for r in (self._r, self._c):
            if r:
                r.close()

    def read_into(self, i, ref_planes, dis_planes, luma_only):
        self._r.read_into(int(self.ref_frames[i]), ref_planes, luma_only)
        self._c.read_into(int(self.cap_frames[i]), dis_planes, luma_only)


def aligned_names(reference_path: str, captured_path: str, output_dir: str | None = None,
                  stamp: str | None = None, ext: str = ".y4m"):
    """Output paths as the reference names them (``:519-527``): ``<base>_<stamp>_aligned<ext>``."""
    out = output_dir or os.path.dirname(os.path.abspath(captured_path))
    if stamp is None:
        d = os.path.basename(out)
        parts = d.split("_")
        stamp = parts[-1] if "_" in d and len(parts) >= 2 and parts[-1].isdigit() else time.strftime("%Y%m%d_%H%M%S")
    rb = os.path.splitext(os.path.basename(reference_path))[0]
    cb = os.path.splitext(os.path.basename(captured_path))[0].replace("_motion_comp", "")
    return (os.path.join(out, f"{rb}_{stamp}_aligned{ext}"), os.path.join(out, f"{cb}_{stamp}_aligned{ext}"))


def write_aligned_y4m(src: AlignedSource, ref_out: str, cap_out: str) -> tuple:
    """Materialise the aligned pair losslessly (both files get ``src.nb_frames`` frames at the reference rate)."""
    from .yuvio import write_y4m
    h = src.open()
    try:
        shapes_n = 1 if src.chroma in (0, 400) else 3
        dtype = np.uint8 if src.bpc == 8 else np.uint16
        from .engine import _plane_shapes
        shapes = _plane_shapes(src)[:shapes_n]
        fr = max(1, int(round(src.fps * 1000)))

        def gen(which):
            rp = [np.empty(s, dtype) for s in shapes]
            dp = [np.empty(s, dtype) for s in shapes]
            for i in range(src.nb_frames):
                h.read_into(i, rp, dp, False)
                yield rp if which == 0 else dp

        chroma = 400 if shapes_n == 1 else src.chroma
        write_y4m(ref_out, gen(0), src.width, src.height, src.bpc, (fr, 1000), chroma)
        write_y4m(cap_out, gen(1), src.width, src.height, src.bpc, (fr, 1000), chroma)
    finally:
        h.close()
    return ref_out, cap_out


def align_by_bookends(reference_path: str, captured_path: str, device: int = 0, adaptive: bool = True,
                      white_threshold: float = 230.0, frame_offset: int = DEFAULT_FRAME_OFFSET, probe_kw=None) -> dict:
    """``BookendAligner.align_bookend_videos`` (reference :277-470) on raw clips: GPU bookend scan of the
    captured clip, content window, frame map.  Returns the reference's result keys plus ``plan`` / ``source``
    (an ``AlignedSource`` ready for ``engine.analyze``); never re-encodes."""
    from . import bookend, yuvio
    probe_kw = probe_kw or {}
    ri, ci = yuvio.probe(reference_path, **probe_kw), yuvio.probe(captured_path, **probe_kw)
    rd = yuvio.ClipReader(ci)
    try:
        lumas = []
        planes = rd.alloc_planes(pinned=False)
        for i in range(ci.nb_frames):
            rd.read_into(i, planes, luma_only=True)
            lumas.append(planes[0].copy())
    finally:
        rd.close()
    runs = bookend.detect_white_bookends(lumas, ci.fps or 30.0, ci.bpc, adaptive, white_threshold, device)
    plan = plan_alignment(ri, ci, runs, frame_offset)
    src = AlignedSource.from_plan(ri, ci, plan)
    return {"alignment_method": "bookend", "offset_frames": 0, "offset_seconds": 0, "confidence": 0.95,
            "aligned_reference": reference_path, "aligned_captured": captured_path,
            "bookend_info": {"first_bookend": plan.window.first_bookend, "last_bookend": plan.window.last_bookend,
                             "content_duration": plan.window.duration, "motion_compensated": False},
            "bookends": runs, "plan": plan, "source": src}


def align_bookend_videos(reference_path: str, captured_path: str, output_dir: str | None = None, device: int = 0,
                         frame_offset: int = DEFAULT_FRAME_OFFSET, **kw) -> dict | None:
    """File-level twin of ``BookendAligner.align_bookend_videos`` (reference :277-470): returns the same result
    dict with ``aligned_reference`` / ``aligned_captured`` pointing at lossless Y4M files that
    ``VMAFAnalyzer.analyze_videos`` takes as they are; ``None`` when no content window can be found (the
    reference's error convention, :300-345)."""
    try:
        res = align_by_bookends(reference_path, captured_path, device=device, frame_offset=frame_offset, **kw)
    except AlignmentError:
        return None
    a, b = aligned_names(reference_path, captured_path, output_dir)
    os.makedirs(os.path.dirname(a) or ".", exist_ok=True)
    write_aligned_y4m(res["source"], a, b)
    res["aligned_reference"], res["aligned_captured"] = a, b
    return res
