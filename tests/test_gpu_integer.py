"""GPU parity (bit-exact) of the integer extractors against the CPU oracle, through the C ABI.

Every comparison is on the raw integer accumulators the kernels produce AND on the derived doubles
(libvmaf's `integer_*` metrics).  The oracle (oracle/vmaf_oracle.c) restates libvmaf's
integer_motion.c / integer_vif.c / integer_adm.c / integer_psnr.c."""
import os

import numpy as np
import pytest

import oracle
from pqa2_b200 import _lib as L
from pqa2_b200 import synth
from pqa2_b200.extractor import FeatureExtractor

pytestmark = pytest.mark.gpu


def _oracle_rows(frames, w, h, bpc, vif_egl=100.0, adm_egl=100.0):
    rows = []
    prev_blur = None
    for (rp, dp) in frames:
        blur = oracle.motion_blur(rp[0], bpc)
        sad = 0 if prev_blur is None else oracle.motion_sad(blur, prev_blur)
        prev_blur = blur
        v = oracle.vif(rp[0], dp[0], bpc, vif_egl)
        a = oracle.adm(rp[0], dp[0], bpc, adm_egl)
        rows.append(dict(sad=sad, motion=oracle.motion_score(sad, w, h), vif=v, adm=a,
                         sse=[oracle.sse(rp[k], dp[k], bpc) for k in range(len(rp))]))
    return rows


def _check(feat, row, w, h, bpc, planes=3):
    raw = np.array(feat.raw[:], dtype=np.int64)
    assert raw[L.RAW_SAD] == row["sad"]
    assert feat.motion == row["motion"]
    acc = raw[L.RAW_VIF:L.RAW_VIF + 28].reshape(4, 7)
    np.testing.assert_array_equal(acc, row["vif"]["acc"])
    for s in range(4):
        assert feat.vif_num[s] == row["vif"]["num"][s]
        assert feat.vif_den[s] == row["vif"]["den"][s]
        assert feat.vif_scale[s] == row["vif"]["score"][s]
    cm = raw[L.RAW_ADM_CM:L.RAW_ADM_CM + 12].reshape(4, 3)
    dn = raw[L.RAW_ADM_DEN:L.RAW_ADM_DEN + 12].reshape(4, 3)
    np.testing.assert_array_equal(cm, row["adm"]["cm"])
    np.testing.assert_array_equal(dn.astype(np.uint64), row["adm"]["den"])
    for s in range(4):
        assert feat.adm_num[s] == row["adm"]["num_scale"][s]
        assert feat.adm_den[s] == row["adm"]["den_scale"][s]
    assert feat.adm2 == row["adm"]["adm2"]
    for k in range(planes):
        assert raw[L.RAW_SSE + k] == row["sse"][k]
    cw, ch = (w + 1) // 2, (h + 1) // 2
    assert feat.psnr_y == oracle.psnr_from_sse(row["sse"][0], bpc, w, h)
    if planes == 3:
        assert feat.psnr_cb == oracle.psnr_from_sse(row["sse"][1], bpc, cw, ch)
        assert feat.psnr_cr == oracle.psnr_from_sse(row["sse"][2], bpc, cw, ch)


FEATS = L.FEAT_VMAF_INT | L.FEAT_PSNR_Y | L.FEAT_PSNR_UV


@pytest.mark.parametrize("w,h,bpc,nframes,batch", [
    (176, 144, 8, 3, 0),
    (322, 242, 8, 2, 0),        # odd-ish dims: VIF floors, ADM ceils
    (333, 251, 8, 2, 0),        # odd dims, unaligned rows
    (640, 360, 8, 6, 4),        # group boundary inside the clip (motion state carried over)
    (416, 240, 10, 3, 2),
    (1920, 1080, 8, 1, 0),
    (960, 540, 10, 1, 0),
    (48, 36, 8, 2, 0),          # smaller than one tile
])
def test_integer_features_bit_exact(w, h, bpc, nframes, batch):
    frames = [synth.frame_pair(3, f, w, h, bpc) for f in range(nframes)]
    rows = _oracle_rows(frames, w, h, bpc)
    with FeatureExtractor(w, h, bpc, 420, FEATS, batch_frames=batch) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        fx.flush()
        out = fx.fetch()
        assert fx.kernel_launches > 0
    for f in range(nframes):
        assert out[f].frame_index == f
        _check(out[f], rows[f], w, h, bpc)


def test_neg_gain_limits():
    """vmaf_*neg models: both enhancement gain limits = 1.0 (reference models/vmaf_v0.6.1neg.json:34-51)."""
    w, h = 480, 270
    frames = [synth.frame_pair(5, f, w, h, 8) for f in range(2)]
    rows = _oracle_rows(frames, w, h, 8, vif_egl=1.0, adm_egl=1.0)
    rows_default = _oracle_rows(frames, w, h, 8)
    assert not np.array_equal(rows[0]["vif"]["acc"], rows_default[0]["vif"]["acc"])   # the patch sharpens
    assert not np.array_equal(rows[0]["adm"]["cm"], rows_default[0]["adm"]["cm"])
    with FeatureExtractor(w, h, 8, 420, FEATS, vif_enhn_gain_limit=1.0, adm_enhn_gain_limit=1.0) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(2):
        _check(out[f], rows[f], w, h, 8)


def test_identical_pair_and_static_clip():
    """SURVEY.md §8c pins: identical ref/dis => vif ~ 1, adm2 ~ 1, psnr_y = 60; static clip => motion = 0."""
    w, h = 352, 288
    rp, _ = synth.frame_pair(1, 0, w, h, 8)
    with FeatureExtractor(w, h, 8, 420, FEATS) as fx:
        for f in range(3):
            fx.submit(f, rp, rp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(3):
        assert out[f].motion == 0.0
        assert out[f].psnr_y == 60.0
        for s in range(4):
            assert abs(out[f].vif_scale[s] - 1.0) < 1e-4
        assert abs(out[f].adm2 - 1.0) < 1e-4


def test_lead_in_and_subsample_flags():
    """Frame shards: a lead-in frame only feeds the motion state; n_subsample skips VIF/ADM."""
    w, h = 320, 180
    frames = [synth.frame_pair(9, f, w, h, 8) for f in range(5)]
    rows = _oracle_rows(frames, w, h, 8)
    # shard [2, 5) with frame 1 as lead-in; frame 3 skipped spatially
    with FeatureExtractor(w, h, 8, 420, FEATS, batch_frames=3) as fx:
        fx.submit(1, *frames[1], L.FRAME_LEAD_IN | L.FRAME_FIRST)
        fx.submit(2, *frames[2], 0)
        fx.submit(3, *frames[3], L.FRAME_SKIP_SPATIAL)
        fx.submit(4, *frames[4], 0)
        out = fx.fetch()
    assert out[0].valid_mask == 0
    _check(out[1], rows[2], w, h, 8)
    assert out[2].valid_mask == L.FEAT_MOTION and out[2].motion == rows[3]["motion"]
    _check(out[3], rows[4], w, h, 8)


def test_device_resident_submission_matches_host_submission():
    from pqa2_b200.extractor import DeviceBuffer
    w, h = 640, 360
    frames = [synth.frame_pair(2, f, w, h, 8, chroma=False) for f in range(3)]
    with FeatureExtractor(w, h, 8, 0, L.FEAT_VMAF_INT | L.FEAT_PSNR_Y) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        host = fx.fetch()
    buf = DeviceBuffer(w * h * 2 * 3)
    for f, (rp, dp) in enumerate(frames):
        buf.upload((2 * f) * w * h, rp[0])
        buf.upload((2 * f + 1) * w * h, dp[0])
    with FeatureExtractor(w, h, 8, 0, L.FEAT_VMAF_INT | L.FEAT_PSNR_Y) as fx:
        for f in range(3):
            fx.submit_device(f, [(buf.ptr + (2 * f) * w * h, w)], [(buf.ptr + (2 * f + 1) * w * h, w)],
                             L.FRAME_FIRST if f == 0 else 0)
        dev = fx.fetch()
    for f in range(3):
        assert list(host[f].raw) == list(dev[f].raw)
    buf.free()


@pytest.mark.parametrize("w,h,bpc", [(176, 144, 8), (333, 251, 8), (416, 240, 10)])
def test_ffmpeg_ssim_filter(w, h, bpc):
    """FFmpeg `ssim` filter (reference app/vmaf_analyzer.py:1057-1064): x264-style integer SSIM per plane.
    Per-window values are bit-identical to the oracle; only the order of the final float sum differs."""
    frames = [synth.frame_pair(6, f, w, h, bpc) for f in range(2)]
    with FeatureExtractor(w, h, bpc, 420, L.FEAT_FFSSIM | L.FEAT_PSNR_Y) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f, (rp, dp) in enumerate(frames):
        assert out[f].valid_mask & L.FEAT_FFSSIM
        for k in range(3):
            assert abs(out[f].ffssim[k] - oracle.ffssim_plane(rp[k], dp[k], bpc)) < 2e-7


def test_analyzer_end_to_end_on_y4m(tmp_path):
    """The drop-in VMAFAnalyzer (reference app/vmaf_analyzer.py:242-616) on a small Y4M pair: same files,
    same results-dict keys, per-frame metrics equal to the oracle's."""
    import json
    from pqa2_b200 import yuvio
    from pqa2_b200.vmaf_analyzer import VMAFAnalyzer
    w, h, n = 320, 180, 5
    frames = [synth.frame_pair(12, f, w, h, 8) for f in range(n)]
    rpath, dpath = tmp_path / "ref_320x180.y4m", tmp_path / "dis_320x180.y4m"
    yuvio.write_y4m(str(rpath), [fr[0] for fr in frames], w, h)
    yuvio.write_y4m(str(dpath), [fr[1] for fr in frames], w, h)
    a = VMAFAnalyzer()
    a.set_output_directory(str(tmp_path))
    a.set_test_name("Clip")
    a.set_advanced_options(pool_method="harmonic_mean")       # reference :383-386: adds psnr=1, ssim=1
    prog, done, errs = [], [], []
    a.analysis_progress.connect(prog.append)
    a.analysis_complete.connect(done.append)
    a.error_occurred.connect(errs.append)
    res = a.analyze_videos(str(rpath), str(dpath), "vmaf_v0.6.1")
    assert not errs and res is not None and done and done[0] is res
    assert prog[-1] == 100 and max(prog[:-1]) <= 95
    for k in ("vmaf_score", "psnr_score", "ssim_score", "json_path", "psnr_log", "ssim_log", "reference_video",
              "distorted_video", "raw_results", "model", "width", "height"):
        assert k in res
    assert res["json_path"].endswith("_vmaf.json") and res["width"] == w and res["reference_video"] == rpath.name
    log = json.load(open(res["json_path"]))
    assert len(log["frames"]) == n and set(log) >= {"version", "fps", "frames", "pooled_metrics", "aggregate_metrics"}
    rows = _oracle_rows(frames, w, h, 8)
    motion = [r["motion"] for r in rows]
    motion2 = [min(motion[i], motion[i + 1]) if i + 1 < n else motion[i] for i in range(n)]
    for i, fr in enumerate(log["frames"]):
        m = fr["metrics"]
        assert fr["frameNum"] == i
        assert abs(m["integer_adm2"] - rows[i]["adm"]["adm2"]) < 5e-7            # %.6f in the log
        assert abs(m["integer_motion2"] - motion2[i]) < 5e-7
        for s in range(4):
            assert abs(m[f"integer_vif_scale{s}"] - rows[i]["vif"]["score"][s]) < 5e-7
        assert "psnr_y" in m and "float_ssim" in m and 0 <= m["vmaf"] <= 100
    assert abs(res["vmaf_score"] - log["pooled_metrics"]["vmaf"]["mean"]) < 1e-6
    assert os.path.exists(res["psnr_log"]) and os.path.exists(res["ssim_log"])
    first = open(res["psnr_log"]).readline()
    assert first.startswith("n:1 mse_avg:") and "psnr_y:" in first
    assert open(res["ssim_log"]).readline().startswith("n:1 Y:")


def test_model_variants_end_to_end():
    """configs[3]: the NEG model (gain limits 1.0 from the model file) and the bootstrap model (21 SVRs ->
    bagging / stddev / CI) through engine.analyze, checked against the oracle's features + the host SVR."""
    from pqa2_b200 import engine, model as M
    w, h, n = 480, 270, 4
    frames = [synth.frame_pair(31, f, w, h, 8) for f in range(n)]
    for name in ("vmaf_v0.6.1neg", "vmaf_b_v0.6.3", "vmaf_4k_v0.6.1"):
        model = M.resolve_model(name)
        res = engine.analyze(engine.SynthSource(w, h, 8, n, seed=31), model, engine.EngineOptions(svr_on_device=True))
        rows = _oracle_rows(frames, w, h, 8, vif_egl=model.vif_enhn_gain_limit, adm_egl=model.adm_enhn_gain_limit)
        motion = [r["motion"] for r in rows]
        motion2 = engine.motion2_from_motion(motion)
        feats = np.array([[r["adm"]["adm2"], motion2[i]] + [r["vif"]["score"][s] for s in range(4)] for i, r in enumerate(rows)])
        want = model.main.predict(feats, False, False, device=None)
        got = np.array([fr["metrics"]["vmaf"] for fr in res["frames"]])
        assert np.max(np.abs(got - want)) < 1e-9, name          # integer features are bit-exact; SVR in double both ways
        m0 = res["frames"][0]["metrics"]
        if name.endswith("neg"):
            assert "integer_adm2_egl_1" in m0 and "integer_vif_scale0_egl_1" in m0
        if "_b_" in name:
            for k in ("vmaf_bagging", "vmaf_stddev", "vmaf_ci_p95_lo", "vmaf_ci_p95_hi"):
                assert k in m0
            assert m0["vmaf_ci_p95_lo"] <= m0["vmaf_bagging"] <= m0["vmaf_ci_p95_hi"]


def test_4k_10bit_single_frame_parity():
    """configs[2] shape: 3840x2160 yuv420p10le, integer extractors, bit-exact vs the oracle (one frame)."""
    w, h = 3840, 2160
    frames = [synth.frame_pair(13, f, w, h, 10, chroma=False) for f in range(2)]
    rows = _oracle_rows(frames, w, h, 10)
    with FeatureExtractor(w, h, 10, 0, L.FEAT_VMAF_INT | L.FEAT_PSNR_Y) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(2):
        _check(out[f], rows[f], w, h, 10, planes=1)


def test_shard_invariance_and_batch_of_clips():
    """results identical however the frames are sharded (here: 1 shard vs 3 shards with lead-in
    frames on one GPU), and configs[4]: a batch of clips through per-device sessions == clip-by-clip."""
    from pqa2_b200 import engine, model as M
    w, h, n = 320, 180, 11
    model = M.resolve_model("vmaf_v0.6.1")
    src = engine.SynthSource(w, h, 8, n, seed=17, chroma=0)
    one = engine.analyze(src, model, engine.EngineOptions(devices=(0,)))
    three = engine.analyze(src, model, engine.EngineOptions(devices=(0, 0, 0)))
    assert [fr["metrics"] for fr in one["frames"]] == [fr["metrics"] for fr in three["frames"]]
    from pqa2_b200 import report
    assert one["pooled_metrics"] == report.pooled_metrics(one["frames"])          # column pooling == per-frame walk
    fmodel = M.resolve_model("vmaf_float_v0.6.1")
    fone = engine.analyze(src, fmodel, engine.EngineOptions(devices=(0,)))
    fthree = engine.analyze(src, fmodel, engine.EngineOptions(devices=(0, 0, 0)))
    assert [fr["metrics"] for fr in fone["frames"]] == [fr["metrics"] for fr in fthree["frames"]]
    clips = [engine.SynthSource(w, h, 8, 5, seed=40 + k, chroma=0) for k in range(5)]
    batch = engine.analyze_batch(clips, model, engine.EngineOptions(), devices=[0, 0])
    for k, c in enumerate(clips):
        solo = engine.analyze(c, model, engine.EngineOptions())
        assert "error" not in batch[k]
        assert [fr["metrics"] for fr in batch[k]["frames"]] == [fr["metrics"] for fr in solo["frames"]]
        assert batch[k]["pooled_metrics"]["vmaf"] == solo["pooled_metrics"]["vmaf"]


def yuvio_probe(path):
    from pqa2_b200 import yuvio
    return yuvio.probe(path)


def test_analyzer_on_mp4_inputs(tmp_path):
    """The reference's real inputs are H.264/MPEG-4 MP4s (app/bookend_alignment.py:526-536): decode through the cv2
    wheel's libavcodec (all planes; luma only if it cannot be driven), score on the GPU, same files and dict as for raw clips."""
    cv2 = pytest.importorskip("cv2")
    from pqa2_b200.vmaf_analyzer import VMAFAnalyzer
    w, h, n = 320, 176, 8
    paths = []
    for name, q in (("ref", None), ("dis", 6)):
        p = str(tmp_path / f"{name}.mp4")
        wr = cv2.VideoWriter(p, cv2.VideoWriter_fourcc(*"mp4v"), 25, (w, h))
        if not wr.isOpened():
            pytest.skip("cv2 cannot encode mp4v here")
        for f in range(n):
            y = synth.ref_luma(2, f, w, h)
            if q:
                y = synth.distort(y, 2, f, 8, q)
            wr.write(cv2.cvtColor(y, cv2.COLOR_GRAY2BGR))
        wr.release()
        paths.append(p)
    a = VMAFAnalyzer()
    a.set_output_directory(str(tmp_path))
    errs = []
    a.error_occurred.connect(errs.append)
    res = a.analyze_videos(paths[0], paths[1])
    assert not errs and res is not None and 0 < res["vmaf_score"] < 100 and len(res["raw_results"]["frames"]) == n
    from pqa2_b200 import avdec
    if avdec.available():
        # all three planes reach the GPU, so the reference's FFmpeg psnr / ssim passes (app/vmaf_analyzer.py:996-1092) run
        # on compressed inputs too: one stats line per frame, Y / U / V columns
        assert res["psnr_log"] and res["ssim_log"] and a.get_video_metadata(paths[0])["pix_fmt"] == "yuv420p"
        pl = open(res["psnr_log"]).read().strip().splitlines()
        sl = open(res["ssim_log"]).read().strip().splitlines()
        assert len(pl) == n and len(sl) == n and "mse_u:" in pl[0] and "psnr_v:" in pl[0] and " U:" in sl[0] and " V:" in sl[0]
        assert yuvio_probe(paths[0]).decoder == "av"
    same = a.analyze_videos(paths[0], paths[0])
    assert same["vmaf_score"] >= 97.4                        # identical pair: adm2 = vif = 1 -> 97.43 at motion 0, more with motion


def test_strided_host_planes_and_12bit():
    """Planes that are views into larger arrays (row stride != width) go through the 2-D copy path; 12-bit
    samples use the same u16 kernels with their own shifts."""
    w, h = 200, 120
    frames = [synth.frame_pair(23, f, w, h, 8, chroma=False) for f in range(2)]
    rows = _oracle_rows(frames, w, h, 8)
    big_r = [np.zeros((h, w + 56), np.uint8) for _ in frames]
    big_d = [np.zeros((h, w + 24), np.uint8) for _ in frames]
    with FeatureExtractor(w, h, 8, 0, L.FEAT_VMAF_INT | L.FEAT_PSNR_Y) as fx:
        for f, (rp, dp) in enumerate(frames):
            big_r[f][:, 8:8 + w] = rp[0]
            big_d[f][:, :w] = dp[0]
            fx.submit(f, [big_r[f][:, 8:8 + w]], [big_d[f][:, :w]], L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(2):
        _check(out[f], rows[f], w, h, 8, planes=1)
    frames = [synth.frame_pair(29, f, 208, 120, 12, chroma=False) for f in range(2)]
    rows = _oracle_rows(frames, 208, 120, 12)
    with FeatureExtractor(208, 120, 12, 0, L.FEAT_VMAF_INT | L.FEAT_PSNR_Y) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(2):
        _check(out[f], rows[f], 208, 120, 12, planes=1)


def test_engine_subsample_and_cancel():
    """n_subsample (reference :379): VIF/ADM on every N-th frame, motion on all; terminate -> None."""
    import threading
    from pqa2_b200 import engine, model as M
    w, h, n = 320, 180, 9
    model = M.resolve_model("vmaf_v0.6.1")
    src = engine.SynthSource(w, h, 8, n, seed=19, chroma=0)
    full = engine.analyze(src, model, engine.EngineOptions())
    sub = engine.analyze(src, model, engine.EngineOptions(n_subsample=3))
    assert [fr["frameNum"] for fr in sub["frames"]] == [0, 3, 6]
    byn = {fr["frameNum"]: fr["metrics"] for fr in full["frames"]}
    for fr in sub["frames"]:
        assert fr["metrics"] == byn[fr["frameNum"]]          # same per-frame values, motion2 from all frames
    ev = threading.Event()
    ev.set()
    assert engine.analyze(src, model, engine.EngineOptions(), cancel=ev) is None
    # the context stays usable after a cancel inside a session
    with engine.Engine() as sess:
        assert sess.analyze(src, model, engine.EngineOptions(), cancel=ev) is None
        again = sess.analyze(src, model, engine.EngineOptions())
        assert [fr["metrics"] for fr in again["frames"]] == [fr["metrics"] for fr in full["frames"]]


@pytest.mark.parametrize("w,h,bpc", [(1920, 1080, 8), (3840, 2160, 10)])
def test_full_size_properties(w, h, bpc):
    """Size-independent properties at the BASELINE.json picture sizes (the oracle needs minutes per 4K clip, so full-size
    clips are checked through invariants): identical pair => vif ~ 1, adm2 ~ 1, psnr = 6 * bpc + 12; static clip =>
    motion = 0; SSE symmetric under swapping ref and dis while motion follows the reference only; integer accumulators
    identical however the frames are grouped (groups of 32/16 vs groups of 3) and sharded (1 vs 2 shards)."""
    from pqa2_b200 import engine, model as M
    feats = L.FEAT_VMAF_INT | L.FEAT_PSNR_Y
    frames = [synth.frame_pair(9, f, w, h, bpc, chroma=False) for f in range(5)]
    with FeatureExtractor(w, h, bpc, 0, feats) as fx:
        for f in range(3):
            fx.submit(f, frames[0][0], frames[0][0], L.FRAME_FIRST if f == 0 else 0)
        same = fx.fetch()
    for f in range(3):
        assert same[f].motion == 0.0 and same[f].psnr_y == 6.0 * bpc + 12.0
        assert abs(same[f].adm2 - 1.0) < 1e-4 and all(abs(same[f].vif_scale[s] - 1.0) < 1e-4 for s in range(4))

    def run(pairs, batch):
        with FeatureExtractor(w, h, bpc, 0, feats, batch_frames=batch) as fx:
            for f, (a, b) in enumerate(pairs):
                fx.submit(f, a, b, L.FRAME_FIRST if f == 0 else 0)
            return [np.array(r.raw[:], dtype=np.int64) for r in fx.fetch()]

    fwd = run(frames, 0)
    assert all(np.array_equal(x, y) for x, y in zip(fwd, run(frames, 3)))             # grouping-invariant
    swp = run([(d, r) for r, d in frames], 0)
    for f in range(5):
        assert fwd[f][L.RAW_SSE] == swp[f][L.RAW_SSE] and fwd[f][L.RAW_SSE] > 0       # SSE symmetric
    assert [int(x[L.RAW_SAD]) for x in fwd[1:]] != [int(x[L.RAW_SAD]) for x in swp[1:]]   # motion reads the reference
    src = engine.SynthSource(w, h, bpc, 5, seed=9, chroma=0)
    mdl = M.resolve_model("vmaf_4k_v0.6.1" if w > 1920 else "vmaf_v0.6.1")
    one = engine.analyze(src, mdl, engine.EngineOptions(devices=(0,)))
    two = engine.analyze(src, mdl, engine.EngineOptions(devices=(0, 0)))
    assert [fr["metrics"] for fr in one["frames"]] == [fr["metrics"] for fr in two["frames"]]
    assert all(0.0 <= fr["metrics"]["vmaf"] <= 100.0 for fr in one["frames"])


def test_repeated_runs_are_bit_identical():
    """Race detector: the persistent kernels hand shared-memory slots from tile to tile; any ordering hole shows up as
    run-to-run differences.  Same 1080p clip four times through one context (integer accumulators and float sums)."""
    w, h = 1920, 1080
    frames = [synth.frame_pair(21, f, w, h, 8, chroma=False) for f in range(6)]
    feats = L.FEAT_VMAF_INT | L.FEAT_VMAF_FLOAT | L.FEAT_PSNR_Y | L.FEAT_FLOAT_SSIM | L.FEAT_FLOAT_MS_SSIM
    with FeatureExtractor(w, h, 8, 0, feats) as fx:
        runs = []
        for rep in range(4):
            fx.reset()
            for f, (a, b) in enumerate(frames):
                fx.submit(f, a, b, L.FRAME_FIRST if f == 0 else 0)
            out = fx.fetch()
            runs.append([(tuple(r.raw[:]), tuple(r.f_vif_scale[:]), r.f_adm2, r.f_motion, r.float_ssim, r.float_ms_ssim)
                         for r in out])
    assert runs[0] == runs[1] == runs[2] == runs[3]


def test_two_gpus_in_one_process_match_one_gpu():
    """VMAFAnalyzer's default is every GPU of the box from ONE process (one context + host thread per device).  Function
    attributes (the > 48 KB dynamic shared memory opt-in) and constant uploads are per device: a second device must get
    its own.  Needs 2 GPUs; skipped on single-GPU boxes."""
    from pqa2_b200 import engine, model as M
    if L.load().bv_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    src = engine.SynthSource(640, 360, 8, 12, seed=5, chroma=0)
    for name in ("vmaf_v0.6.1", "vmaf_float_v0.6.1"):
        mdl = M.resolve_model(name)
        opt1 = engine.EngineOptions(psnr=True, ssim=True, ms_ssim=True, devices=(0,))
        opt2 = engine.EngineOptions(psnr=True, ssim=True, ms_ssim=True, devices=(0, 1))
        opt3 = engine.EngineOptions(psnr=True, ssim=True, ms_ssim=True, devices=(1,))
        one, two, other = (engine.analyze(src, mdl, o) for o in (opt1, opt2, opt3))
        assert [fr["metrics"] for fr in one["frames"]] == [fr["metrics"] for fr in two["frames"]]
        assert [fr["metrics"] for fr in one["frames"]] == [fr["metrics"] for fr in other["frames"]]
