"""GPU parity at the BASELINE.json picture sizes that round 1 only reached through invariants:

* configs[3] -- 2160p 10-bit with the NEG model's gain limits (bit-exact accumulators) and with the bootstrap model end to end;
* configs[1] at high bit depth -- the float extractors at 1080p 10-bit and on one 2160p 10-bit frame pair;
* configs[4] over several devices in one process (needs >= 2 GPUs; also run by bench.py's N >= 2 arm).

Oracle: oracle/ (CPU restatement, parity unpinned); one or two frames per case keep the scalar oracle to seconds."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # the sibling test modules' helpers

import oracle
from pqa2_b200 import _lib as L
from pqa2_b200 import engine, model as M, synth
from pqa2_b200.extractor import FeatureExtractor
from test_gpu_float import _check as check_float, _oracle_rows as float_rows
from test_gpu_integer import _check as check_int, _oracle_rows as int_rows

pytestmark = pytest.mark.gpu


def test_4k_10bit_neg_gain_limits_bit_exact():
    """vmaf_v0.6.1neg on configs[3]'s shape: vif_enhn_gain_limit = adm_enhn_gain_limit = 1.0 (models/vmaf_v0.6.1neg.json:34-51)."""
    w, h, bpc = 3840, 2160, 10
    model = M.resolve_model("vmaf_v0.6.1neg")
    assert model.vif_enhn_gain_limit == 1.0 and model.adm_enhn_gain_limit == 1.0
    frames = [synth.frame_pair(14, 0, w, h, bpc, chroma=False)]
    rows = int_rows(frames, w, h, bpc, vif_egl=1.0, adm_egl=1.0)
    plain = oracle.vif(frames[0][0][0], frames[0][1][0], bpc, 100.0)
    assert not np.array_equal(plain["acc"], rows[0]["vif"]["acc"])              # the limit bites on this content
    with FeatureExtractor(w, h, bpc, 0, L.FEAT_VMAF_INT | L.FEAT_PSNR_Y, vif_enhn_gain_limit=1.0,
                          adm_enhn_gain_limit=1.0) as fx:
        fx.submit(0, frames[0][0], frames[0][1], L.FRAME_FIRST)
        out = fx.fetch()
    check_int(out[0], rows[0], w, h, bpc, planes=1)


@pytest.mark.parametrize("name", ["vmaf_v0.6.1neg", "vmaf_b_v0.6.3"])
def test_4k_10bit_model_variants_end_to_end(name):
    """configs[3] through engine.analyze: per-frame score == host SVR on the oracle's features; bootstrap adds its four
    extra metrics, NEG renames its features with the _egl_1 suffix."""
    w, h, bpc, n = 3840, 2160, 10, 2
    model = M.resolve_model(name)
    res = engine.analyze(engine.SynthSource(w, h, bpc, n, seed=15, chroma=0), model, engine.EngineOptions())
    frames = [synth.frame_pair(15, f, w, h, bpc, chroma=False) for f in range(n)]
    rows = int_rows(frames, w, h, bpc, vif_egl=model.vif_enhn_gain_limit, adm_egl=model.adm_enhn_gain_limit)
    motion2 = engine.motion2_from_motion([r["motion"] for r in rows])
    feats = np.array([[r["adm"]["adm2"], motion2[i]] + [r["vif"]["score"][s] for s in range(4)] for i, r in enumerate(rows)])
    want = model.main.predict(feats, False, False, device=None)
    got = np.array([fr["metrics"]["vmaf"] for fr in res["frames"]])
    assert np.max(np.abs(got - want)) < 1e-9
    m0 = res["frames"][1]["metrics"]
    sfx = "_egl_1" if name.endswith("neg") else ""
    assert m0[f"integer_adm2{sfx}"] == rows[1]["adm"]["adm2"] and m0[f"integer_vif_scale2{sfx}"] == rows[1]["vif"]["score"][2]
    if "_b_" in name:
        boots = np.array([b.predict(feats, False, True, device=None) for b in model.bootstrap])      # [20, n], unclipped
        assert abs(m0["vmaf_bagging"] - min(max(boots[:, 1].mean(), 0.0), 100.0)) < 1e-9
        assert abs(m0["vmaf_stddev"] - boots[:, 1].std()) < 1e-9
        assert m0["vmaf_ci_p95_lo"] <= m0["vmaf_bagging"] <= m0["vmaf_ci_p95_hi"]


@pytest.mark.parametrize("w,h,bpc,n", [(1920, 1080, 10, 2), (3840, 2160, 10, 1)])
def test_float_extractors_at_high_bit_depth_full_size(w, h, bpc, n):
    frames = [synth.frame_pair(16, f, w, h, bpc, chroma=False) for f in range(n)]
    rows = float_rows(frames, bpc, ssim=True, ms_ssim=True)
    mask = L.FEAT_VMAF_FLOAT | L.FEAT_PSNR_Y | L.FEAT_FLOAT_SSIM | L.FEAT_FLOAT_MS_SSIM
    with FeatureExtractor(w, h, bpc, 0, mask) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(n):
        check_float(out[f], rows[f], ssim=True, ms_ssim=True)


def test_golden_fixtures_at_baseline_shapes():
    """tests/golden/fullsize_oracle.json: the raw integer accumulators of one 1080p 8-bit and one 2160p 10-bit pair
    (oracle outputs, written by tests/tools/make_golden.py) straight against the kernels."""
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "fullsize_oracle.json")
    g = json.load(open(path))
    for c in g["cases"]:
        w, h, bpc = c["w"], c["h"], c["bpc"]
        rp, dp = synth.frame_pair(c["seed"], 0, w, h, bpc, chroma=False)
        with FeatureExtractor(w, h, bpc, 0, L.FEAT_VMAF_INT | L.FEAT_PSNR_Y, vif_enhn_gain_limit=c["egl"],
                              adm_enhn_gain_limit=c["egl"]) as fx:
            fx.submit(0, rp, dp, L.FRAME_FIRST)
            out = fx.fetch()
        raw = np.array(out[0].raw[:], dtype=np.int64)
        assert raw[L.RAW_VIF:L.RAW_VIF + 28].reshape(4, 7).tolist() == c["vif_acc"]
        assert raw[L.RAW_ADM_CM:L.RAW_ADM_CM + 12].reshape(4, 3).tolist() == c["adm_cm"]
        assert raw[L.RAW_ADM_DEN:L.RAW_ADM_DEN + 12].reshape(4, 3).tolist() == c["adm_den"]
        assert raw[L.RAW_SSE] == c["sse_y"] and out[0].adm2 == c["adm2"]


def test_batch_of_clips_over_two_devices_shares_one_model():
    """engine.analyze_batch hands ONE VmafModel to a worker thread per GPU; the device-side SVR keeps a
    mirror per device under a lock.  21 bootstrap predicts per clip widen the window for the old race."""
    if L.load().bv_device_count() < 2:
        pytest.skip("needs two GPUs (bench.py --gpus N >= 2 runs the same check on the scaling box)")
    w, h = 320, 180
    model = M.resolve_model("vmaf_b_v0.6.3")
    clips = [engine.SynthSource(w, h, 8, 6, seed=60 + k, chroma=0) for k in range(12)]
    batch = engine.analyze_batch(clips, model, engine.EngineOptions(svr_on_device=True), devices=[0, 1])
    for k, c in enumerate(clips):
        solo = engine.analyze(c, model, engine.EngineOptions(svr_on_device=False, devices=(0,)))
        assert "error" not in batch[k]
        got = [fr["metrics"]["vmaf"] for fr in batch[k]["frames"]]
        want = [fr["metrics"]["vmaf"] for fr in solo["frames"]]
        assert np.max(np.abs(np.array(got) - np.array(want))) < 1e-9


def test_file_ingest_paths_agree(tmp_path):
    """Raw clip files reach the GPU either through reader threads + a pinned ring (default) or, opt-in, as a CUDA-registered
    mapping of the page cache (no CPU copy): same frames, same scores, all planes (psnr / ssim stats features on)."""
    import os
    import shutil
    from pqa2_b200 import yuvio
    w, h, n = 640, 360, 40
    base = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path)
    d = os.path.join(base, f"b200vmaf_test_{os.getpid()}")
    os.makedirs(d, exist_ok=True)
    try:
        frames = [synth.frame_pair(9, f, w, h, 8) for f in range(n)]
        rp, dp = os.path.join(d, "r.y4m"), os.path.join(d, "d.y4m")
        yuvio.write_y4m(rp, (f[0] for f in frames), w, h)
        yuvio.write_y4m(dp, (f[1] for f in frames), w, h)
        ri, di = yuvio.probe(rp), yuvio.probe(dp)
        model = M.resolve_model("vmaf_v0.6.1")
        opt = engine.EngineOptions(psnr=True, ffmpeg_psnr=True, ffmpeg_ssim=True, batch_frames=8, reader_threads=3,
                                   devices=(0, 0))
        ring_src = engine.FileSource(ri, di)
        assert not ring_src.zero_copy
        ring = engine.analyze(ring_src, model, opt)
        direct = engine.analyze(engine.SynthSource(w, h, 8, n, seed=9), model, opt)
        assert [f["metrics"] for f in ring["frames"]] == [f["metrics"] for f in direct["frames"]]
        assert np.array_equal(ring["rows"].arr["raw"], direct["rows"].arr["raw"])
        msrc = engine.FileSource(ri, di, mapped=True)
        if not msrc.zero_copy:
            pytest.skip("cudaHostRegister refuses file mappings on this filesystem")
        mapped = engine.analyze(msrc, model, opt)
        msrc.release()
        assert [f["metrics"] for f in mapped["frames"]] == [f["metrics"] for f in ring["frames"]]
        assert np.array_equal(mapped["rows"].arr["ffssim"], ring["rows"].arr["ffssim"])
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_dynamic_chunks_match_one_shard():
    """engine.analyze deals a long clip to its devices in chunks (here two contexts on one GPU, chunks of 24 frames so that
    chunk borders fall inside launch groups): same bits as one shard, integer and float model."""
    w, h, n = 320, 180, 200
    src = engine.SynthSource(w, h, 8, n, seed=23, chroma=0)
    for name in ("vmaf_v0.6.1", "vmaf_float_v0.6.1"):
        model = M.resolve_model(name)
        one = engine.analyze(src, model, engine.EngineOptions(devices=(0,)))
        dyn = engine.analyze(src, model, engine.EngineOptions(devices=(0, 0), dynamic_chunk=24))
        assert [fr["metrics"] for fr in one["frames"]] == [fr["metrics"] for fr in dyn["frames"]]
        assert np.array_equal(one["rows"].arr["raw"], dyn["rows"].arr["raw"])


def test_three_contexts_per_gpu_match_one_context():
    """What engine.analyze does by default for a clip of >= 1536 frames: three contexts side by side on the GPU, chunks
    from a shared counter, log entries built while frames are in flight -- against one context and a one-pass build: every
    frame's metrics, the raw accumulators and the pooled report, integer and float model."""
    w, h, n = 192, 108, 1600

    class Cycled(engine.SynthSource):                       # 40 distinct frames, cycled (synthesis would dominate the test)
        def read_into(self, i, ref_planes, dis_planes, luma_only):
            super().read_into(i % 40, ref_planes, dis_planes, luma_only)

    src = Cycled(w, h, 8, n, seed=31, chroma=0)
    for name in ("vmaf_v0.6.1", "vmaf_float_v0.6.1"):
        model = M.resolve_model(name)
        with engine.Engine() as sess:
            one = sess.analyze(src, model, engine.EngineOptions(devices=(0,), contexts_per_device=1, psnr=True))
            auto = sess.analyze(src, model, engine.EngineOptions(devices=(0,), psnr=True))
            assert len({k[1] for k in sess._fx}) == 3         # three shard contexts were brought up for the second call
        assert [fr["frameNum"] for fr in auto["frames"]] == list(range(n))
        assert [fr["metrics"] for fr in one["frames"]] == [fr["metrics"] for fr in auto["frames"]]
        assert np.array_equal(one["rows"].arr["raw"], auto["rows"].arr["raw"])
        assert one["pooled_metrics"] == auto["pooled_metrics"]
