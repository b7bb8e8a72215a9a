"""The C oracle against the independent NumPy mirror (oracle/mirror_np.py): integer VIF accumulators and integer motion
bit for bit, float VIF sums to float rounding -- on seeded synthetic frames that reach every branch of the statistic
(log / non-log, sigma12 <= 0, gain above and below the NEG limit), 8- and 10-bit, odd sizes.

Neither side is libvmaf (parity unpinned); two separately written restatements that agree accumulator
for accumulator rule out transcription slips in either, not a shared misreading."""
import numpy as np
import pytest

import oracle
from oracle import mirror_np as MN
from pqa2_b200 import synth

CASES = [(3, 176, 144, 8, {}), (8, 208, 120, 10, {}), (5, 161, 97, 8, {}), (9, 240, 136, 12, {}),
         (4, 192, 108, 8, {"strength": 3, "q": 8})]


@pytest.mark.parametrize("seed,w,h,bpc,kw", CASES)
@pytest.mark.parametrize("egl", [100.0, 1.0])
def test_integer_vif_accumulators_agree(seed, w, h, bpc, kw, egl):
    rp, dp = synth.frame_pair(seed, 1, w, h, bpc, chroma=False, **kw)
    got = oracle.vif(rp[0], dp[0], bpc, egl)
    want = MN.vif_int(rp[0], dp[0], bpc, egl)
    assert got["acc"].tolist() == want.tolist()
    np.testing.assert_array_equal(got["score"], MN.vif_scores(want))
    assert all(want[s, 6] > 0 for s in range(3))             # the log branch is reached at every scale that matters


def test_integer_vif_flat_content_takes_the_non_log_branch():
    """Half the picture is flat (sigma1^2 < 2): the non-log accumulators carry it, and both restatements agree there."""
    rng = np.random.default_rng(1)
    rp, dp = synth.frame_pair(3, 0, 208, 120, 8, chroma=False)
    ref, dis = rp[0].copy(), dp[0].copy()
    ref[:, :104] = 90
    dis[:, :104] = 90 + (rng.integers(0, 2, (120, 104))).astype(np.uint8)
    got, want = oracle.vif(ref, dis, 8)["acc"], MN.vif_int(ref, dis, 8)
    assert got.tolist() == want.tolist()
    assert all(want[s, 3] > 0 and want[s, 6] > 0 for s in range(3)) and want[0, 2] > 0


def test_vif_neg_limit_changes_the_numerator_only():
    rp, dp = synth.frame_pair(5, 0, 240, 136, 8, chroma=False)
    a, b = MN.vif_int(rp[0], dp[0], 8, 100.0), MN.vif_int(rp[0], dp[0], 8, 1.0)
    assert (a[:, [1, 2, 3, 4, 6]] == b[:, [1, 2, 3, 4, 6]]).all() and (a[:, 0] != b[:, 0]).any()


@pytest.mark.parametrize("seed,w,h,bpc,kw", CASES[:4])
def test_integer_motion_agrees(seed, w, h, bpc, kw):
    a = synth.ref_luma(seed, 0, w, h, bpc)
    b = synth.ref_luma(seed, 1, w, h, bpc)
    ba, bb = oracle.motion_blur(a, bpc), oracle.motion_blur(b, bpc)
    np.testing.assert_array_equal(ba, MN.motion_blur_int(a, bpc))
    np.testing.assert_array_equal(bb, MN.motion_blur_int(b, bpc))
    assert oracle.motion_sad(ba, bb) == MN.motion_sad_int(ba, bb) > 0


@pytest.mark.parametrize("seed,w,h,bpc", [(3, 176, 144, 8), (8, 208, 120, 10)])
def test_float_vif_and_the_row_accumulator_question(seed, w, h, bpc):
    """VERDICT r1 weak #1 asked whether oracle/vmaf_float_oracle.c sums the VIF num / den in double where libvmaf's
    vif_statistic_s keeps a float sum per row.  It does what libvmaf does (vmaf_float_oracle.c:104-126: float `rn`, `rd`
    per row, rows added into double); the mirror's per-row-float variant must therefore reproduce it to double rounding,
    and the all-double variant shows what the accumulator type is worth: < 2e-6 relative on the sums and on the scores,
    two orders below what the 1e-4 VMAF tolerance needs at the feature level."""
    rp, dp = synth.frame_pair(seed, 1, w, h, bpc, chroma=False)
    rf, df = oracle.picture_copy(rp[0], bpc, -128.0), oracle.picture_copy(dp[0], bpc, -128.0)
    np.testing.assert_array_equal(rf, (rp[0].astype(np.float32) / np.float32(1 << (bpc - 8))) - np.float32(128))
    c = oracle.f_vif(rf, df)
    lit = lambda s: np.ctypeslib.as_array(oracle._flib().orc_f_vif_filter(s), ([17, 9, 5, 3][s],)).copy()
    n32, d32 = MN.vif_float(rf, df, row_acc=np.float32, taps_of=lit)
    np.testing.assert_allclose(c["num"], n32, rtol=1e-12)
    np.testing.assert_allclose(c["den"], d32, rtol=1e-12)
    n64, d64 = MN.vif_float(rf, df, row_acc=np.float64, taps_of=lit)
    np.testing.assert_allclose(n32, n64, rtol=2e-6)
    np.testing.assert_allclose(d32, d64, rtol=2e-6)
    assert np.max(np.abs(n32 / d32 - n64 / d64)) < 2e-6


def test_float_vif_kernel_tables_are_the_normalised_gaussians():
    """libvmaf ships the float kernels as literals (vif_options.h); the oracle carries them as literals too.  Evaluating
    exp(-k^2 / (2 (N/5)^2)) / sum in double and rounding to float reproduces them to one unit in the last place."""
    for s, n in enumerate([17, 9, 5, 3]):
        want = np.ctypeslib.as_array(oracle._flib().orc_f_vif_filter(s), (n,))
        np.testing.assert_allclose(MN._gauss(n), want, rtol=0, atol=1e-7)
        assert abs(float(want.astype(np.float64).sum()) - 1.0) < 1e-6


@pytest.mark.parametrize("seed,w,h,bpc", [(3, 176, 144, 8), (5, 161, 97, 8), (8, 208, 120, 10), (9, 242, 137, 12)])
def test_integer_adm_scale0_wavelet_bands_agree(seed, w, h, bpc):
    """The db2 decomposition every later ADM stage builds on: a, v, h, d of the reference picture, value for value."""
    rp, dp = synth.frame_pair(seed, 0, w, h, bpc, chroma=False)
    got = oracle.adm(rp[0], dp[0], bpc, want_bands=True)["ref_bands_s0"]
    want = MN.adm_dwt_scale0(rp[0], bpc)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)
    assert np.abs(want[1:]).max() > 0                     # the detail bands are not trivially zero


@pytest.mark.parametrize("seed,w,h,bpc,egl,vd,dh", [
    (3, 176, 144, 8, 100.0, 3.0, 1080), (5, 161, 97, 8, 1.0, 3.0, 1080), (8, 208, 120, 10, 100.0, 3.0, 1080),
    (9, 242, 137, 12, 1.2, 3.0, 1080), (4, 352, 288, 8, 100.0, 2.5, 720), (2, 64, 48, 8, 100.0, 3.0, 1080)])
def test_integer_adm_accumulators_agree(seed, w, h, bpc, egl, vd, dh):
    """Integer ADM end to end -- four wavelet levels with their per-scale shifts, the Q30 reciprocal table with the
    15-bit reduction of large divisors, the 1-degree angle test, the enhancement-gain limit, fixed-point CSF (hard-coded
    integers at the default viewing condition, derived ones otherwise), the 3x3 threshold with folded band edges, per-row
    rounding of both cube sums -- as whole-band array code against the C oracle's per-pixel loops: all 24 raw accumulators
    of a frame, bit for bit."""
    rp, dp = synth.frame_pair(seed, 1, w, h, bpc, chroma=False)
    c = oracle.adm(rp[0], dp[0], bpc, egl, vd, dh)
    cm, dn = MN.adm_int(rp[0], dp[0], bpc, egl, vd, dh)
    assert [[int(v) for v in r] for r in c["cm"]] == cm
    assert [[int(v) for v in r] for r in c["den"]] == dn
    assert all(v > 0 for r in cm for v in r) and all(v > 0 for r in dn for v in r)


def test_integer_adm_gain_limit_moves_the_numerator_only():
    """The limit acts where the angle test passes: a tight limit changes the numerator accumulators and leaves the
    denominator (reference only) alone, in the mirror as in the oracle."""
    rp, dp = synth.frame_pair(5, 1, 161, 97, 8, chroma=False)
    loose, tight = MN.adm_int(rp[0], dp[0], 8, 100.0), MN.adm_int(rp[0], dp[0], 8, 1.0)
    assert loose[1] == tight[1] and loose[0] != tight[0]


@pytest.mark.parametrize("seed,w,h,bpc,egl", [(3, 176, 144, 8, 100.0), (5, 322, 242, 8, 1.0), (8, 208, 120, 10, 100.0)])
def test_float_adm_agrees(seed, w, h, bpc, egl):
    """float ADM (DWT, decoupling with the 1-degree angle test and the gain limit, CSF, 3x3 contrast-masking threshold, cube
    sums) as whole-plane fp32 array code against the C oracle's per-pixel loops."""
    rp, dp = synth.frame_pair(seed, 1, w, h, bpc, chroma=False)
    rf, df = oracle.picture_copy(rp[0], bpc, -128.0), oracle.picture_copy(dp[0], bpc, -128.0)
    c = oracle.f_adm(rf, df, egl)
    nums, dens, adm2 = MN.adm_float(rf, df, egl, rf_of=oracle.adm_rfactor)
    np.testing.assert_allclose(c["num_scale"], nums, rtol=2e-6)
    np.testing.assert_allclose(c["den_scale"], dens, rtol=2e-6)
    assert abs(c["adm2"] - adm2) < 2e-6 and 0.3 < adm2 < 1.1


def test_csf_factors_follow_the_watson_model():
    """rfactor = 1 / Q(scale, theta): the mirror's evaluation, the oracle's, and the values SURVEY.md Appendix A.4 lists."""
    listed = [(0.0173815, 0.0058907), (0.0319848, 0.0142991), (0.0433727, 0.0243969), (0.0456734, 0.0313127)]
    for s in range(4):
        a, b = MN.rfactor(s), oracle.adm_rfactor(s)
        np.testing.assert_array_equal(a, b)               # `k * temp * temp` is float arithmetic in adm_tools.h
        assert abs(a[0] - listed[s][0]) < 5e-7 and a[0] == a[1] and abs(a[2] - listed[s][1]) < 5e-7
    # the integer path's hard-coded scale-0 factors (include/libvmaf_spec.h, tagged L) against the model
    assert abs(36453 / 2.0 ** 21 - MN.rfactor(0)[0]) < 1e-6 and abs(49417 / 2.0 ** 23 - MN.rfactor(0)[2]) < 1e-6


@pytest.mark.parametrize("seed,w,h,bpc", [(4, 352, 288, 8), (8, 416, 240, 10), (4, 335, 253, 8)])
def test_ms_ssim_agrees_with_a_float64_restatement(seed, w, h, bpc):
    """float_ms_ssim: 5 scales, valid 11-tap Gaussian moments, 9-tap low-pass decimation with symmetric borders, per-scale
    means of l, c, s, exponents -- the fp32 oracle against the same pipeline in float64.  The fp32 variance subtraction costs
    a few 1e-6 on the c and s means of the finest scale and less than 1e-5 on the score."""
    rp, dp = synth.frame_pair(seed, 0, w, h, bpc, chroma=False)
    r0, d0 = oracle.picture_copy(rp[0], bpc, 0.0), oracle.picture_copy(dp[0], bpc, 0.0)
    got, lcs = oracle.f_ms_ssim(r0, d0)
    want, wl = MN.ms_ssim(r0, d0)
    assert lcs.shape == wl.shape == (5, 3)
    np.testing.assert_allclose(lcs, wl, atol=2e-5)
    assert abs(got - want) < 1e-5 and 0.5 < want < 1.0


@pytest.mark.parametrize("seed,w,h,bpc", [(3, 176, 144, 8), (5, 161, 97, 8), (8, 208, 120, 10), (9, 242, 137, 12)])
def test_ffmpeg_ssim_and_sse_agree(seed, w, h, bpc):
    """FFmpeg's ssim filter (integer 4x4 block sums, 8x8 windows, float window score, float row sums) and the squared error
    both psnr flavours start from, on every plane of a 4:2:0 pair; 8-bit constants 416 / 235963 fall out of the formula."""
    rp, dp = synth.frame_pair(seed, 1, w, h, bpc)
    for k in range(3):
        assert oracle.ffssim_plane(rp[k], dp[k], bpc) == MN.ffssim_plane(rp[k], dp[k], bpc)
        assert int(oracle.sse(rp[k], dp[k], bpc)) == MN.sse_plane(rp[k], dp[k])
    assert 0.2 < MN.ffssim_plane(rp[0], dp[0], bpc) < 1.0
    assert (int(.01 * .01 * 255 * 255 * 64 + .5), int(.03 * .03 * 255 * 255 * 64 * 63 + .5)) == (416, 235963)


@pytest.mark.parametrize("seed,w,h,bpc", [(3, 176, 144, 8), (5, 161, 97, 8), (8, 208, 120, 10)])
def test_float_motion_agrees(seed, w, h, bpc):
    """float motion: blurred planes value for value (fp32, taps left to right), SAD with a float sum per row."""
    a, _ = synth.frame_pair(seed, 0, w, h, bpc, chroma=False)
    b, _ = synth.frame_pair(seed, 1, w, h, bpc, chroma=False)
    fa, fb = oracle.picture_copy(a[0], bpc, -128.0), oracle.picture_copy(b[0], bpc, -128.0)
    ba, bb = oracle.f_motion_blur(fa), oracle.f_motion_blur(fb)
    ma, mb = MN.motion_blur_float(fa), MN.motion_blur_float(fb)
    np.testing.assert_array_equal(ba, ma)
    np.testing.assert_array_equal(bb, mb)
    assert oracle.f_motion_sad(ba, bb) == MN.motion_sad_float(ma, mb) > 0.5


@pytest.mark.parametrize("seed,w,h,bpc", [(3, 176, 144, 8), (8, 208, 120, 10), (4, 640, 540, 8), (6, 1000, 780, 8)])
def test_float_ssim_agrees_with_a_float64_restatement(seed, w, h, bpc):
    """float_ssim incl. the box decimation of large pictures (factor 2 at 640x540, 3 at 1000x780, odd factor = centred
    window): the fp32 oracle against float64 array code; the fp32 variance subtraction is worth ~2e-6."""
    rp, dp = synth.frame_pair(seed, 1, w, h, bpc, chroma=False)
    r0, c0 = oracle.picture_copy(rp[0], bpc, 0.0), oracle.picture_copy(dp[0], bpc, 0.0)
    assert abs(oracle.f_ssim(r0, c0) - MN.ssim_float(r0, c0)) < 5e-6
