"""GPU parity of the float extractors (vmaf_float_* models, float_ssim, float_ms_ssim) against the CPU
oracle (oracle/vmaf_float_oracle.c), through the C ABI.

Tolerance mode (BASELINE.json north_star): per-frame VMAF within 1e-4, pooled mean within 1e-5.
The feature-level tolerances below are tighter than what that needs: the kernels evaluate the
oracle's fp32 arithmetic in the same tap order (FMA contraction and the order of the final sums are
the only differences)."""
import numpy as np
import pytest

import oracle
from pqa2_b200 import _lib as L
from pqa2_b200 import engine, model as M, synth
from pqa2_b200.extractor import FeatureExtractor

pytestmark = pytest.mark.gpu

FEATS = L.FEAT_VMAF_FLOAT | L.FEAT_PSNR_Y
REL_SUM = 1e-5        # raw num/den sums (fp32 maps; FMA contraction + summation order differ from the oracle)
ABS_FEAT = 1e-5       # vif_scale*, adm2, adm_scale*, motion


def _oracle_rows(frames, bpc, vif_egl=100.0, adm_egl=100.0, ssim=False, ms_ssim=False):
    rows, prev = [], None
    for (rp, dp) in frames:
        rows.append(oracle.float_features(rp[0], dp[0], bpc, prev_ref=prev, vif_egl=vif_egl, adm_egl=adm_egl,
                                          psnr=True, ssim=ssim, ms_ssim=ms_ssim))
        prev = rp[0]
    return rows


def _check(feat, row, ssim=False, ms_ssim=False):
    for s in range(4):
        np.testing.assert_allclose(feat.f_vif_num[s], row["vif"]["num"][s], rtol=REL_SUM)
        np.testing.assert_allclose(feat.f_vif_den[s], row["vif"]["den"][s], rtol=REL_SUM)
        assert abs(feat.f_vif_scale[s] - row[f"vif_scale{s}"]) < ABS_FEAT
        np.testing.assert_allclose(feat.f_adm_num[s], row["adm"]["num_scale"][s], rtol=REL_SUM)
        np.testing.assert_allclose(feat.f_adm_den[s], row["adm"]["den_scale"][s], rtol=REL_SUM)
        assert abs(feat.f_adm_scale[s] - row[f"adm_scale{s}"]) < ABS_FEAT
    assert abs(feat.f_adm2 - row["adm2"]) < ABS_FEAT
    assert abs(feat.f_motion - row["motion"]) < ABS_FEAT
    assert feat.psnr_y == row["psnr_y"]
    if ssim:
        assert abs(feat.float_ssim - row["float_ssim"]) < 1e-6
    if ms_ssim:
        assert abs(feat.float_ms_ssim - row["float_ms_ssim"]) < 1e-6


@pytest.mark.parametrize("w,h,bpc,nframes,batch", [
    (176, 144, 8, 3, 0),
    (322, 242, 8, 2, 0),
    (333, 251, 8, 2, 0),
    (640, 360, 8, 5, 2),        # group boundary inside the clip
    (416, 240, 10, 3, 2),
    (960, 540, 8, 1, 0),
    (48, 36, 8, 2, 0),
])
def test_float_vmaf_features(w, h, bpc, nframes, batch):
    frames = [synth.frame_pair(3, f, w, h, bpc, chroma=False) for f in range(nframes)]
    rows = _oracle_rows(frames, bpc)
    with FeatureExtractor(w, h, bpc, 0, FEATS, batch_frames=batch) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
        assert fx.kernel_launches > 0
    for f in range(nframes):
        assert out[f].valid_mask & L.FEAT_VMAF_FLOAT == L.FEAT_VMAF_FLOAT
        _check(out[f], rows[f])


@pytest.mark.parametrize("w,h,bpc", [(352, 288, 8), (640, 360, 8), (1280, 720, 8), (416, 240, 10), (335, 253, 8)])
def test_float_ssim_and_ms_ssim(w, h, bpc):
    frames = [synth.frame_pair(4, f, w, h, bpc, chroma=False) for f in range(2)]
    rows = _oracle_rows(frames, bpc, ssim=True, ms_ssim=True)
    with FeatureExtractor(w, h, bpc, 0, FEATS | L.FEAT_FLOAT_SSIM | L.FEAT_FLOAT_MS_SSIM) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(2):
        _check(out[f], rows[f], ssim=True, ms_ssim=True)


def test_float_1080p_headline_shape():
    """configs[1] shape: 1080p 8-bit, float model + psnr + float_ssim + float_ms_ssim."""
    w, h = 1920, 1080
    frames = [synth.frame_pair(11, f, w, h, 8, chroma=False) for f in range(2)]
    rows = _oracle_rows(frames, 8, ssim=True, ms_ssim=True)
    with FeatureExtractor(w, h, 8, 0, FEATS | L.FEAT_FLOAT_SSIM | L.FEAT_FLOAT_MS_SSIM) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(2):
        _check(out[f], rows[f], ssim=True, ms_ssim=True)


def test_float_neg_gain_limits():
    w, h = 480, 270
    frames = [synth.frame_pair(5, f, w, h, 8, chroma=False) for f in range(2)]
    rows = _oracle_rows(frames, 8, vif_egl=1.0, adm_egl=1.0)
    rows_default = _oracle_rows(frames, 8)
    assert abs(rows[0]["vif_scale0"] - rows_default[0]["vif_scale0"]) > 1e-4
    with FeatureExtractor(w, h, 8, 0, FEATS, vif_enhn_gain_limit=1.0, adm_enhn_gain_limit=1.0) as fx:
        for f, (rp, dp) in enumerate(frames):
            fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(2):
        _check(out[f], rows[f])


def test_float_identical_pair_and_static_clip():
    w, h = 352, 288
    rp, _ = synth.frame_pair(1, 0, w, h, 8, chroma=False)
    with FeatureExtractor(w, h, 8, 0, FEATS | L.FEAT_FLOAT_SSIM | L.FEAT_FLOAT_MS_SSIM) as fx:
        for f in range(3):
            fx.submit(f, rp, rp, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    for f in range(3):
        assert out[f].f_motion == 0.0
        assert out[f].f_adm2 == 1.0
        for s in range(4):
            assert abs(out[f].f_vif_scale[s] - 1.0) < 1e-5
        assert abs(out[f].float_ssim - 1.0) < 1e-9
        assert abs(out[f].float_ms_ssim - 1.0) < 1e-9


def test_float_model_vmaf_within_north_star_tolerance():
    """End to end: engine.analyze with vmaf_float_v0.6.1 vs the oracle's features through the same SVR:
    per-frame |dVMAF| < 1e-4, pooled mean < 1e-5."""
    w, h, n = 640, 360, 6
    model = M.resolve_model("vmaf_float_v0.6.1")
    res = engine.analyze(engine.SynthSource(w, h, 8, n, seed=21, chroma=0), model,
                         engine.EngineOptions(psnr=True, ssim=True, ms_ssim=True))
    frames = [synth.frame_pair(21, f, w, h, 8, chroma=False) for f in range(n)]
    rows = _oracle_rows(frames, 8, ssim=True, ms_ssim=True)
    motion = [r["motion"] for r in rows]
    motion[0] = 0.0
    motion2 = engine.motion2_from_motion(motion)
    feats = np.array([[r["adm2"], motion2[i], r["vif_scale0"], r["vif_scale1"], r["vif_scale2"], r["vif_scale3"]]
                      for i, r in enumerate(rows)])
    want = model.main.predict(feats, False, False, device=None)
    got = np.array([fr["metrics"]["vmaf"] for fr in res["frames"]])
    assert np.max(np.abs(got - want)) < 1e-4
    assert abs(got.mean() - want.mean()) < 1e-5
    for i, fr in enumerate(res["frames"]):
        assert abs(fr["metrics"]["float_ssim"] - rows[i]["float_ssim"]) < 1e-6
        assert abs(fr["metrics"]["float_ms_ssim"] - rows[i]["float_ms_ssim"]) < 1e-6
        assert "vif_scale0" in fr["metrics"] and "adm2" in fr["metrics"] and "motion2" in fr["metrics"]


def test_gpu_matches_golden_fixtures():
    """The committed regression vectors (tests/golden, oracle outputs) straight against the kernels."""
    import glob
    import json
    import os
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "oracle_*.json"))):
        g = json.load(open(path))
        c = g["case"]
        w, h, bpc = c["w"], c["h"], c["bpc"]
        ms = "float_ms_ssim" in g["frames"][0]["float"]
        mask = L.FEAT_VMAF_INT | L.FEAT_VMAF_FLOAT | L.FEAT_PSNR_Y | L.FEAT_PSNR_UV | L.FEAT_FFSSIM | L.FEAT_FLOAT_SSIM
        mask |= L.FEAT_FLOAT_MS_SSIM if ms else 0
        with FeatureExtractor(w, h, bpc, 420, mask) as fx:
            for f in range(c["n"]):
                rp, dp = synth.frame_pair(c["seed"], f, w, h, bpc)
                fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
            out = fx.fetch()
        for f, row in enumerate(g["frames"]):
            raw = np.array(out[f].raw[:], dtype=np.int64)
            assert raw[L.RAW_SAD] == row["sad"]
            assert raw[L.RAW_VIF:L.RAW_VIF + 28].reshape(4, 7).tolist() == row["vif_acc"]
            assert raw[L.RAW_ADM_CM:L.RAW_ADM_CM + 12].reshape(4, 3).tolist() == row["adm_cm"]
            assert raw[L.RAW_ADM_DEN:L.RAW_ADM_DEN + 12].reshape(4, 3).tolist() == row["adm_den"]
            assert out[f].adm2 == row["adm2"]
            assert raw[L.RAW_SSE:L.RAW_SSE + 3].tolist() == row["sse"]
            for k in range(3):
                assert abs(out[f].ffssim[k] - row["ffssim"][k]) < 2e-7
            fl = row["float"]
            assert abs(out[f].f_adm2 - fl["adm2"]) < ABS_FEAT and abs(out[f].f_motion - fl["motion"]) < ABS_FEAT
            for s in range(4):
                assert abs(out[f].f_vif_scale[s] - fl[f"vif_scale{s}"]) < ABS_FEAT
            assert abs(out[f].float_ssim - fl["float_ssim"]) < 1e-6
            if ms:
                assert abs(out[f].float_ms_ssim - fl["float_ms_ssim"]) < 1e-6
