"""CPU tests of the oracle itself (no GPU): regression vectors under tests/golden/ (made by
tests/tools/make_golden.py from the oracle -- the reference ships no golden vectors), and the pins
SURVEY.md §8c lists: filter tables, identical-pair and static-clip invariants, monotonicity."""
import glob
import json
import os

import numpy as np
import pytest

import oracle
from pqa2_b200 import synth

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "oracle_*.json")))


def _crc(a):
    return int(np.bitwise_xor.reduce(a.astype(np.uint32).ravel() * np.uint32(2654435761)))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_golden(path):
    g = json.load(open(path))
    c = g["case"]
    w, h, bpc = c["w"], c["h"], c["bpc"]
    prev_blur, prev_ref = None, None
    for f, row in enumerate(g["frames"]):
        rp, dp = synth.frame_pair(c["seed"], f, w, h, bpc)
        assert [_crc(rp[0]), _crc(dp[0])] == row["luma_crc"], "synthetic generator changed"
        blur = oracle.motion_blur(rp[0], bpc)
        sad = 0 if prev_blur is None else oracle.motion_sad(blur, prev_blur)
        prev_blur = blur
        assert sad == row["sad"]
        v, a = oracle.vif(rp[0], dp[0], bpc), oracle.adm(rp[0], dp[0], bpc)
        assert v["acc"].tolist() == row["vif_acc"]
        assert a["cm"].tolist() == row["adm_cm"]
        assert [[int(x) for x in r] for r in a["den"]] == row["adm_den"]
        assert a["adm2"] == row["adm2"]
        assert [oracle.sse(rp[k], dp[k], bpc) for k in range(3)] == row["sse"]
        for k in range(3):
            assert oracle.ffssim_plane(rp[k], dp[k], bpc) == row["ffssim"][k]
        fl = oracle.float_features(rp[0], dp[0], bpc, prev_ref=prev_ref, psnr=True, ssim=True,
                                   ms_ssim="float_ms_ssim" in row["float"])
        prev_ref = rp[0]
        for k, val in row["float"].items():
            assert fl[k] == val, k


def test_filter_tables():
    for s, n in enumerate((17, 9, 5, 3)):
        assert int(np.ctypeslib.as_array(oracle.lib().orc_vif_filter(s), (17,))[:n].sum()) == 65536
    t = np.ctypeslib.as_array(oracle.lib().orc_vif_log2_table(), (65536,))
    assert t[32768] == 30720 and t[65535] == 32768
    L = oracle._flib()
    for s, n in enumerate((17, 9, 5, 3)):
        f = np.ctypeslib.as_array(L.orc_f_vif_filter(s), (n,))
        assert abs(float(f.sum()) - 1.0) < 1e-6 and np.allclose(f, f[::-1])
    assert abs(L.orc_f_log2_approx(8.0) - 3.0) < 1e-6 and abs(L.orc_f_log2_approx(1.5) - np.log2(1.5)) < 1e-4


@pytest.mark.parametrize("bpc", [8, 10])
def test_identical_pair_and_static_clip(bpc):
    w, h = 192, 128
    rp, _ = synth.frame_pair(2, 0, w, h, bpc)
    v, a = oracle.vif(rp[0], rp[0], bpc), oracle.adm(rp[0], rp[0], bpc)
    assert all(abs(s - 1.0) < 1e-4 for s in v["score"]) and abs(a["adm2"] - 1.0) < 1e-4
    assert oracle.psnr_from_sse(oracle.sse(rp[0], rp[0], bpc), bpc, w, h) == 6.0 * bpc + 12.0
    b = oracle.motion_blur(rp[0], bpc)
    assert oracle.motion_sad(b, b) == 0 and oracle.motion_score(0, w, h) == 0.0
    fl = oracle.float_features(rp[0], rp[0], bpc, prev_ref=rp[0], ssim=True)
    assert fl["adm2"] == 1.0 and fl["motion"] == 0.0 and abs(fl["float_ssim"] - 1.0) < 1e-12
    assert all(abs(fl[f"vif_scale{s}"] - 1.0) < 1e-5 for s in range(4))
    assert oracle.ffssim_plane(rp[0], rp[0], bpc) == pytest.approx(1.0, abs=1e-6)


def test_monotone_in_distortion_strength():
    w, h = 256, 144
    prev = None
    for strength in (1, 2, 3):
        rp, dp = synth.frame_pair(4, 1, w, h, 8, chroma=False, strength=strength)
        v, a = oracle.vif(rp[0], dp[0], 8), oracle.adm(rp[0], dp[0], 8)
        fl = oracle.float_features(rp[0], dp[0], 8)
        cur = (v["score"][0], a["adm2"], fl["vif_scale0"], fl["adm2"])
        if prev is not None:
            assert all(c < p for c, p in zip(cur, prev)), (strength, cur, prev)
        prev = cur


def test_integer_and_float_extractors_agree_roughly():
    """Two independent restatements (fixed-point and fp32) of the same features."""
    w, h = 320, 192
    rp, dp = synth.frame_pair(6, 0, w, h, 8, chroma=False)
    v, a, fl = oracle.vif(rp[0], dp[0], 8), oracle.adm(rp[0], dp[0], 8), oracle.float_features(rp[0], dp[0], 8)
    for s in range(4):
        assert abs(v["score"][s] - fl[f"vif_scale{s}"]) < 5e-3
    assert abs(a["adm2"] - fl["adm2"]) < 5e-3


def test_oracle_psnr_and_ssim_against_independent_implementations():
    """Two of the oracle's metrics have textbook definitions that other libraries in this image implement independently:
    PSNR (cv2.PSNR) and single-scale SSIM with the 11x11 sigma-1.5 Gaussian window over the valid region (Wang et al.
    2004, restated here with scipy in float64).  libvmaf's float_ssim is iqa's implementation of exactly that formula
    for pictures with min(w, h) < 384 (no decimation); iqa's three-term l*c*s with C3 = C2/2 equals the two-term form."""
    cv2 = pytest.importorskip("cv2")
    from scipy.ndimage import correlate1d
    w, h = 320, 180
    for seed in (4, 11):
        (ref, _, _), (dis, _, _) = synth.frame_pair(seed, 1, w, h, 8)
        sse = oracle.sse(ref, dis, 8)
        assert oracle.psnr_from_sse(sse, 8, w, h) == pytest.approx(cv2.PSNR(ref, dis), abs=1e-9)
        k = np.exp(-0.5 * (np.arange(-5, 6) / 1.5) ** 2)
        k /= k.sum()

        def g(a):
            return correlate1d(correlate1d(a, k, axis=0, mode="constant"), k, axis=1, mode="constant")[5:-5, 5:-5]

        x, y = ref.astype(np.float64), dis.astype(np.float64)
        mx, my = g(x), g(y)
        sxx, syy, sxy = g(x * x) - mx * mx, g(y * y) - my * my, g(x * y) - mx * my
        c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
        ssim = float(np.mean((2 * mx * my + c1) * (2 * sxy + c2) / ((mx * mx + my * my + c1) * (sxx + syy + c2))))
        got = oracle.float_features(ref, dis, 8, ssim=True)["float_ssim"]
        # iqa works in float32 (sums of x^2 up to 65025 before the variance subtraction), so agreement with the float64
        # textbook value is ~2e-4, two orders below what a wrong window, constant or region would cause
        assert got == pytest.approx(ssim, abs=5e-4)


def test_synthetic_frames_exercise_every_branch():
    """SURVEY.md §8d asks the synthetic clips to reach the VIF non-log branch (flat regions), the log branch, a gain above
    1 (where the NEG limits bite) and the ADM angle flag both ways; a generator that stopped doing so would quietly
    weaken every parity test built on it."""
    w, h = 352, 288
    (ref, _, _), (dis, _, _) = synth.frame_pair(3, 2, w, h, 8)
    v = oracle.vif(ref, dis, 8)
    acc = np.asarray(v["acc"])
    assert acc[0, 3] > 0 and acc[0, 6] > 0, "scale 0 must see both the non-log and the log branch"
    assert acc[0, 6] + acc[0, 3] == w * h                                  # every pixel takes exactly one of them
    vneg = oracle.vif(ref, dis, 8, 1.0)
    assert np.asarray(vneg["acc"])[0, 0] < acc[0, 0], "a gain limit of 1 must lower the numerator (sharpened patch)"
    a, aneg = oracle.adm(ref, dis, 8), oracle.adm(ref, dis, 8, 1.0)
    assert aneg["adm2"] < a["adm2"], "the ADM gain limit only acts where the angle flag is set"
    same = oracle.adm(ref, ref, 8, 1.0)
    assert abs(same["adm2"] - 1.0) < 1e-4


def test_oracle_matches_fullsize_golden_1080p():
    """tests/golden/fullsize_oracle.json (1080p 8-bit case; the 2160p cases are checked against the kernels by the -m gpu
    suite, the scalar oracle needs several seconds for each of them)."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fullsize_oracle.json")))
    c = next(c for c in g["cases"] if c["w"] == 1920)
    rp, dp = synth.frame_pair(c["seed"], 0, c["w"], c["h"], c["bpc"], chroma=False)
    v, a = oracle.vif(rp[0], dp[0], c["bpc"], c["egl"]), oracle.adm(rp[0], dp[0], c["bpc"], c["egl"])
    assert v["acc"].tolist() == c["vif_acc"] and a["cm"].tolist() == c["adm_cm"] and a["adm2"] == c["adm2"]
    assert [[int(x) for x in r] for r in a["den"]] == c["adm_den"] and oracle.sse(rp[0], dp[0], c["bpc"]) == c["sse_y"]


@pytest.mark.parametrize("seed,w,h,bpc", [(3, 352, 288, 8), (8, 416, 240, 10), (5, 640, 360, 8)])
def test_integer_and_float_restatements_agree(seed, w, h, bpc):
    """Two differently built restatements of the same features -- fixed point with hard-coded tables, shifts and CSF integers
    (integer_vif / integer_adm / integer_motion) against float formulas (vif / adm / motion) -- agree to a few 1e-4, as
    libvmaf's own integer and float extractors do.  A wrong shift, table or CSF constant on either side (the M / L items of
    SURVEY.md Appendix A, e.g. the scale-0 integers 36453 / 49417 or the ADM DWT shifts of scales 1-3) would show here as a
    difference orders of magnitude larger; it does not replace a libvmaf log, it bounds what such a log could still move."""
    rp, dp = synth.frame_pair(seed, 1, w, h, bpc, chroma=False)
    pp, _ = synth.frame_pair(seed, 0, w, h, bpc, chroma=False)
    i = oracle.integer_features(rp[0], dp[0], bpc)
    f = oracle.float_features(rp[0], dp[0], bpc, prev_ref=pp[0])
    assert abs(i["integer_adm2"] - f["adm2"]) < 5e-4
    for s in range(4):
        assert abs(i[f"integer_vif_scale{s}"] - f[f"vif_scale{s}"]) < 5e-4
        assert abs(i[f"integer_adm_scale{s}"] - f[f"adm_scale{s}"]) < 3e-3
    im = oracle.motion_score(oracle.motion_sad(oracle.motion_blur(rp[0], bpc), oracle.motion_blur(pp[0], bpc)), w, h)
    assert im > 0.5 and abs(im - f["motion"]) < 1e-4
    # through the (shared) SVR the two paths land within a few hundredths of a VMAF point
    from pqa2_b200 import model as M
    mi, mf = M.resolve_model("vmaf_v0.6.1"), M.resolve_model("vmaf_float_v0.6.1")
    xi = np.array([[i["integer_adm2"], 0.0] + [i[f"integer_vif_scale{s}"] for s in range(4)]])
    xf = np.array([[f["adm2"], 0.0] + [f[f"vif_scale{s}"] for s in range(4)]])
    assert abs(mi.main.predict(xi)[0] - mf.main.predict(xf)[0]) < 0.05
