"""Bookend (white-frame) detection: decision logic on CPU, GPU statistics vs numpy (row f1 of SURVEY.md §8;
reference app/bookend_alignment.py:755-1134, app/reference_analyzer.py:112-152)."""
import numpy as np
import pytest

from pqa2_b200 import bookend as B
from pqa2_b200 import synth


def test_frame_test_matches_reference_branches():
    # low std: mean > 0.95 * threshold decides (reference :1006-1009)
    assert B.is_white_frame(220, 5, 0.0, 230, 30) is True
    assert B.is_white_frame(215, 5, 0.0, 230, 30) is False
    # high std: mean > threshold, or mean > 0.9 threshold with > 70 % white pixels (:1010-1020)
    assert B.is_white_frame(231, 60, 0.0, 230, 30) is True
    assert B.is_white_frame(210, 60, 0.75, 230, 30) is True
    assert B.is_white_frame(210, 60, 0.65, 230, 30) is False
    assert B.is_white_frame(200, 60, 0.99, 230, 30) is False


def test_thresholds():
    thr, sd = B.brightness_thresholds([100, 110, 250, 105], [20, 22, 3, 21], adaptive=True)
    assert thr[0] >= 220 and abs(thr[1] - 0.9 * thr[0]) < 1e-9 and sd == min(45.0, np.mean([20, 22, 3, 21]) * 1.8)
    thr, _ = B.brightness_thresholds([100], [10], adaptive=False, white_threshold=230)
    assert thr == [230, 207.0, 184.0]


def _clip(w, h, n, white_at, bpc=8):
    sc = 1 << (bpc - 8)
    frames = []
    for f in range(n):
        if f in white_at:
            y = np.full((h, w), 240 * sc, np.uint8 if bpc == 8 else np.uint16)
            y[::7, ::5] -= 3 * sc
        else:
            y = synth.ref_luma(3, f, w, h, bpc)
        frames.append(y)
    return frames


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(320, 180, 8), (333, 251, 8), (416, 240, 10)])
def test_gpu_luma_stats_match_numpy(w, h, bpc):
    frames = _clip(w, h, 9, {0, 1, 2}, bpc)
    sc = 1 << (bpc - 8)
    thr = (200 * sc, 120 * sc, 90 * sc)
    got = B.luma_stats(frames, bpc, thr, chunk=4)
    want = B.luma_stats_host(frames, bpc, thr)
    for g, t in zip(got, want):
        assert abs(g.mean - t.mean) < 1e-9 * max(1, t.mean)
        assert abs(g.std - t.std) < 1e-6 * max(1, t.std)
        assert g.ratios == pytest.approx(t.ratios, abs=1e-12)


@pytest.mark.gpu
def test_gpu_detects_bookends():
    w, h, n, fps = 320, 180, 60, 30.0
    white = set(range(5, 10)) | set(range(40, 46))
    runs = B.detect_white_bookends(_clip(w, h, n, white), fps)
    assert [(r["start_frame"], r["end_frame"]) for r in runs] == [(5, 9), (40, 45)]
    stats = B.luma_stats(_clip(w, h, 12, {0, 1}), 8, (200, 200, 200))
    assert B.begins_with_bookend(stats) is True
    assert B.begins_with_bookend(B.luma_stats(_clip(w, h, 12, set()), 8, (200, 200, 200))) is False
