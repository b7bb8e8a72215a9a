"""Bookend alignment as a frame map (row f4 of SURVEY.md §8; reference app/bookend_alignment.py:318-389 for the
content window, :530-602 for the trim / re-time command lines restated in pqa2_b200.alignment)."""
import os

import numpy as np
import pytest

from pqa2_b200 import alignment as A
from pqa2_b200 import synth, yuvio


def _run(s, e, fps=30.0):
    return {"start_frame": s, "end_frame": e, "start_time": s / fps, "end_time": e / fps, "frame_count": e - s + 1}


def test_content_window_single_loop():
    # first.end + 1.5 frames .. last.start - 1.5 frames (reference :336-338)
    win = A.select_content([_run(0, 9), _run(100, 109)], 30.0, ref_duration=3.0)
    assert win.start_time == pytest.approx(9 / 30 + 1.5 / 30)
    assert win.end_time == pytest.approx(100 / 30 - 1.5 / 30)
    assert win.duration == pytest.approx(win.end_time - win.start_time)
    assert not win.multi_loop and win.loop_index == 0


def test_content_window_needs_two_bookends_and_valid_timing():
    with pytest.raises(A.AlignmentError):
        A.select_content([_run(0, 9)], 30.0, 3.0)
    with pytest.raises(A.AlignmentError):                     # runs 2 frames apart: end <= start after the buffers
        A.select_content([_run(0, 9), _run(11, 20)], 30.0, 3.0)


def test_content_window_multi_loop_picks_closest_pair():
    # three loops of ~3 s, ~2 s, ~3.1 s: reference duration 2 s selects the middle pair (reference :357-385)
    b = [_run(0, 9), _run(100, 109), _run(170, 179), _run(273, 282)]
    win = A.select_content(b, 30.0, ref_duration=2.0)
    assert win.multi_loop and win.loop_index == 1
    assert win.first_bookend is b[1] and win.last_bookend is b[2]
    assert win.start_time == pytest.approx(109 / 30 + 0.05) and win.end_time == pytest.approx(170 / 30 - 0.05)
    # only two bookends but far too much content: keep one reference duration from the start (:386-389)
    win = A.select_content([_run(0, 9), _run(400, 409)], 30.0, ref_duration=3.0)
    assert win.multi_loop and win.duration == pytest.approx(3.0)


def test_captured_frame_map_same_rate():
    # -itsoffset 6/30 -ss (start + 0.2): first kept frame has j/30 + 0.2 >= start + 0.2  ->  j >= start * 30
    m = A.captured_frame_map(10.5 / 30, 5, 30.0, 30.0, 1000)
    assert m.tolist() == [11, 12, 13, 14, 15]
    m = A.captured_frame_map(10.5 / 30, 5, 30.0, 30.0, 1000, frame_offset=0)          # no offset: 0.2 s = 6 frames later
    assert m.tolist() == [17, 18, 19, 20, 21]
    assert A.captured_frame_map(10.5 / 30, 8, 30.0, 30.0, 14).tolist() == [11, 12, 13]   # capture ends first


def test_captured_frame_map_retimes_to_reference_rate():
    m = A.captured_frame_map(1.0, 4, 30.0, 60.0, 10_000)      # 60 fps capture, 30 fps reference: every 2nd frame
    j0 = int(np.ceil((1.0 + 0.2 - 6 / 60) * 60 - 1e-9))
    assert m.tolist() == [j0, j0 + 2, j0 + 4, j0 + 6]
    m = A.captured_frame_map(0.0, 5, 60.0, 30.0, 10_000, frame_offset=6)   # 30 fps capture shown at 60: duplicates
    assert m.tolist() == [0, 1, 1, 2, 2]


def _write_pair(tmp_path, w=96, h=64, n_ref=12, lead=9, tail=8):
    ref = [list(synth.frame_pair(5, f, w, h, 8)[0]) for f in range(n_ref)]
    cap = []
    white = [np.full((h, w), 240, np.uint8), np.full((h // 2, w // 2), 128, np.uint8), np.full((h // 2, w // 2), 128, np.uint8)]
    for _ in range(lead):
        cap.append(white)
    for f in range(n_ref):
        cap.append(list(synth.frame_pair(5, f, w, h, 8)[1]))
    for _ in range(tail):
        cap.append(white)
    rp, cp = str(tmp_path / "ref.y4m"), str(tmp_path / "cap.y4m")
    yuvio.write_y4m(rp, ref, w, h, 8, (30, 1))
    yuvio.write_y4m(cp, cap, w, h, 8, (30, 1))
    return rp, cp, ref, cap


def test_aligned_source_reads_mapped_frames_and_writes_lossless_pair(tmp_path):
    rp, cp, ref, cap = _write_pair(tmp_path)
    ri, ci = yuvio.probe(rp), yuvio.probe(cp)
    lead, n = 9, 12
    runs = [_run(0, lead - 1), _run(lead + n, lead + n + 7)]
    plan = A.plan_alignment(ri, ci, runs)
    # content starts at (8 + 1.5)/30 s; offset 6 frames and +0.2 s cancel at 30 fps -> first kept frame = ceil(9.5) = 10,
    # i.e. the reference's recipe skips the first content frame of this capture; the capture then runs out 1 frame early
    assert plan.cap_frames[0] == 10
    src = A.AlignedSource.from_plan(ri, ci, plan)
    assert src.nb_frames == plan.n_frames and (src.width, src.height, src.bpc, src.chroma) == (96, 64, 8, 420)
    h = src.open()
    try:
        rpl = [np.empty(s, np.uint8) for s in ri.plane_shapes()]
        dpl = [np.empty(s, np.uint8) for s in ri.plane_shapes()]
        for k in (0, 3, src.nb_frames - 1):
            h.read_into(k, rpl, dpl, False)
            for p in range(3):
                assert np.array_equal(rpl[p], ref[k][p])
                assert np.array_equal(dpl[p], cap[int(plan.cap_frames[k])][p])
    finally:
        h.close()
    a, b = A.aligned_names(rp, cp, str(tmp_path), stamp="20240101_000000")
    assert os.path.basename(a) == "ref_20240101_000000_aligned.y4m" and os.path.basename(b) == "cap_20240101_000000_aligned.y4m"
    A.write_aligned_y4m(src, a, b)
    ai, bi = yuvio.probe(a), yuvio.probe(b)
    assert ai.nb_frames == bi.nb_frames == src.nb_frames and ai.fps == pytest.approx(30.0)
    rd = yuvio.ClipReader(bi)
    try:
        pl = rd.alloc_planes(pinned=False)
        rd.read_into(2, pl)
        assert np.array_equal(pl[0], cap[int(plan.cap_frames[2])][0])          # bit-identical: no re-encode
    finally:
        rd.close()


def test_aligned_source_rejects_mismatched_geometry(tmp_path):
    a, b = str(tmp_path / "a.y4m"), str(tmp_path / "b.y4m")
    yuvio.write_y4m(a, [list(synth.frame_pair(1, 0, 64, 48, 8)[0])], 64, 48, 8)
    yuvio.write_y4m(b, [list(synth.frame_pair(1, 0, 96, 64, 8)[0])], 96, 64, 8)
    with pytest.raises(A.AlignmentError):
        A.AlignedSource(yuvio.probe(a), yuvio.probe(b), [0], [0])


def test_aligned_names_take_stamp_from_directory(tmp_path):
    d = tmp_path / "Test_20250102_123456"
    d.mkdir()
    a, b = A.aligned_names("/x/ref.mp4", str(d / "cap_motion_comp.mp4"))
    assert a == str(d / "ref_123456_aligned.y4m") and b == str(d / "cap_123456_aligned.y4m")   # reference :506-522


@pytest.mark.gpu
def test_align_by_bookends_scores_the_content_without_reencode(tmp_path):
    """Captured = white lead-in + distorted content + white tail.  The GPU bookend scan finds both runs, the frame
    map pairs reference frame k with the right captured frame, and the aligned score is bit-identical to scoring
    the same frame pairs directly."""
    from pqa2_b200 import engine, model as M
    rp, cp, ref, cap = _write_pair(tmp_path, w=320, h=180, n_ref=14, lead=9, tail=8)
    res = A.align_by_bookends(rp, cp)
    assert [(r["start_frame"], r["end_frame"]) for r in res["bookends"]] == [(0, 8), (23, 30)]
    plan, src = res["plan"], res["source"]
    assert plan.cap_frames.tolist() == list(range(10, 10 + plan.n_frames))
    mdl = M.resolve_model("vmaf_v0.6.1")
    got = engine.analyze(src, mdl, engine.EngineOptions(psnr=True))

    class Direct(engine.FrameSource):
        width, height, bpc, chroma, nb_frames, fps = 320, 180, 8, 420, plan.n_frames, 30.0

        def read_into(self, i, r, d, luma_only):
            for p in range(1 if luma_only else 3):
                r[p][...] = ref[i][p]
                d[p][...] = cap[int(plan.cap_frames[i])][p]

    want = engine.analyze(Direct(), mdl, engine.EngineOptions(psnr=True))
    assert [f["metrics"] for f in got["frames"]] == [f["metrics"] for f in want["frames"]]
    assert got["pooled_metrics"]["vmaf"]["mean"] == want["pooled_metrics"]["vmaf"]["mean"]
    # the pairing is off by one content frame (the reference's +0.2 s pad), so scores are finite but below identity
    assert 0.0 <= got["pooled_metrics"]["vmaf"]["mean"] <= 100.0


@pytest.mark.gpu
def test_sweep_writes_the_artifacts_the_history_tab_indexes(tmp_path):
    """configs[4] in small: clips dealt over device sessions, one test directory per clip with the libvmaf log, the
    per-test CSV and the metadata file, plus the combined CSV (rows e / f3)."""
    import csv
    import json
    from pqa2_b200 import engine, model as M, sweep
    clips = [engine.SynthSource(320, 180, 8, 4, seed=60 + k, chroma=0) for k in range(3)]
    out = sweep.run_sweep(clips, M.resolve_model("vmaf_v0.6.1"), str(tmp_path), names=["a", "b", "c"], devices=[0, 0],
                          stamp="20250101_000000")
    assert len(out["results"]) == 3 and all("error" not in r for r in out["results"])
    for r, c in zip(out["results"], clips):
        log = json.load(open(r["json_path"]))
        assert r["json_path"].endswith("_vmaf.json") and len(log["frames"]) == 4                 # results_tab.py:3105
        solo = engine.analyze(c, M.resolve_model("vmaf_v0.6.1"), engine.EngineOptions(psnr=True, ssim=True))
        assert log["pooled_metrics"]["vmaf"]["mean"] == pytest.approx(solo["pooled_metrics"]["vmaf"]["mean"], abs=5e-7)
        meta = [f for f in os.listdir(r["test_dir"]) if f.endswith("_metadata.json")]
        assert len(meta) == 1 and json.load(open(os.path.join(r["test_dir"], meta[0])))["video_details"]["frame_count"] == 4
    rows = list(csv.reader(open(out["combined_csv"])))
    assert len(rows) == 4 and [x[0] for x in rows[1:]] == ["a", "b", "c"]
    assert float(rows[1][2]) == pytest.approx(out["results"][0]["vmaf_score"], abs=5e-5)
