"""include/libvmaf_spec.h against the closed forms its tables come from.  Every literal there is a restatement from memory
of libvmaf / iqa / FFmpeg sources; where a table has a published mathematical definition (Gaussian windows, the db2
wavelet, the CDF 9/7 low-pass, Q-format roundings, constants derived from formulas) the definition pins it without any
reference binary: a mistyped digit in any of ~150 literals fails here."""
import math
import os
import re

import numpy as np

SPEC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "libvmaf_spec.h")


def _macros():
    txt = open(SPEC).read().replace("\\\n", " ")
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"^#define\s+(SPEC_\w+)\s+(.+)$", txt, flags=re.M):
        body = m.group(2).strip()
        nums = re.findall(r"[-+]?(?:\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+|\d+(?:[eE][-+]?\d+)?)", body.replace("ll", ""))
        out[m.group(1)] = [float(x) for x in nums]
    return out


M = _macros()


def _gauss(n):
    k = np.arange(n) - n // 2
    g = np.exp(-(k.astype(np.float64) ** 2) / (2.0 * (n / 5.0) ** 2))
    return g / g.sum()


def test_vif_and_motion_windows_are_gaussians_with_sigma_n_over_5():
    for n in (17, 9, 5, 3):
        q16, f32 = np.array(M[f"SPEC_VIF_Q16_{n}"]), np.array(M[f"SPEC_VIF_F32_{n}"])
        assert len(q16) == n and len(f32) == n
        g = _gauss(n)
        assert q16.sum() == 65536 and np.abs(q16 - g * 65536).max() < 1.0          # rounded, then nudged to sum to 2^16
        # libvmaf ships these as float literals that were normalised in single precision: up to 12 ulp (9e-8) off the
        # double-precision Gaussian at the centre tap of the 17-tap window, so the formula pins them to 1e-7 only
        assert np.abs(f32 - g).max() < 1e-7 and abs(f32.astype(np.float32).astype(np.float64).sum() - 1) < 1e-6
        assert np.array_equal(q16, q16[::-1]) and np.array_equal(f32, f32[::-1])
    assert M["SPEC_MOTION_Q16_5"] == M["SPEC_VIF_Q16_5"]
    assert np.abs(np.array(M["SPEC_MOTION_F32_5"]) - _gauss(5)).max() < 1e-7


def test_db2_wavelet_taps():
    r3, d = math.sqrt(3.0), 4.0 * math.sqrt(2.0)
    lo = np.array([(1 + r3) / d, (3 + r3) / d, (3 - r3) / d, (1 - r3) / d])
    hi = np.array([lo[3], -lo[2], lo[1], -lo[0]])                                  # quadrature mirror of the low-pass
    assert np.abs(np.array(M["SPEC_DWT_LO_F32"]) - lo).max() < 1e-12
    assert np.abs(np.array(M["SPEC_DWT_HI_F32"]) - hi).max() < 1e-12
    assert M["SPEC_DWT_LO_Q15"] == [float(round(v * 32768)) for v in lo]
    assert M["SPEC_DWT_HI_Q15"] == [float(round(v * 32768)) for v in hi]
    assert M["SPEC_DWT_LO_SUM_Q15"] == [sum(M["SPEC_DWT_LO_Q15"])]
    assert abs(lo.sum() - math.sqrt(2.0)) < 1e-12 and abs(hi.sum()) < 1e-12 and abs((lo * lo).sum() - 1) < 1e-12


def test_ssim_windows():
    k = np.arange(11) - 5
    g = np.exp(-(k ** 2) / (2 * 1.5 ** 2))
    g /= g.sum()
    assert np.abs(np.array(M["SPEC_SSIM_GAUSS11"]) - g).max() < 1e-6               # iqa prints 6 decimals
    # iqa's 9-tap decimation filter is the CDF 9/7 analysis low-pass, normalised to unit DC gain
    cdf = np.array([0.026748757411, -0.016864118443, -0.078223266529, 0.266864118443, 0.602949018236])
    cdf = np.concatenate([cdf, cdf[-2::-1]])
    lpf = np.array(M["SPEC_MS_SSIM_LPF9"])
    assert np.abs(lpf - cdf / cdf.sum()).max() < 5e-5 and abs(lpf.sum() - 1) < 5e-6 and np.array_equal(lpf, lpf[::-1])
    assert abs(sum(M["SPEC_MS_SSIM_EXPONENTS"]) - 1.0001) < 1e-12                  # Wang et al. 2003: 0.0448 ... 0.1333
    assert M["SPEC_SSIM_K1"] == [0.01] and M["SPEC_SSIM_K2"] == [0.03]


def test_derived_integers():
    assert M["SPEC_FFSSIM_C1"] == [float(int(.01 * .01 * 255 * 255 * 64 + .5))]
    assert M["SPEC_FFSSIM_C2"] == [float(int(.03 * .03 * 255 * 255 * 64 * 63 + .5))]
    # applied as (c * |csf_a| + 2048) >> 12, i.e. c / 4096 = 32 / 30 and 32 / 15 (the two planes differ by 5 fraction bits)
    assert M["SPEC_ADM_ONE_BY_30_Q16"] == [float(round(2 ** 17 / 30))] == [4369.0]
    assert M["SPEC_ADM_ONE_BY_15_Q16"] == [float(round(2 ** 17 / 15))] == [8738.0]
    assert M["SPEC_ADM_ONE_BY_30_Q32"] == [float(round(2 ** 32 / 30))] and M["SPEC_ADM_ONE_BY_15_Q32"] == [float(round(2 ** 32 / 15))]
    # the hard-coded scale-0 CSF integers (Q21, Q21, Q23) against Watson's model at 3.0 x 1080: 1 / Q(lambda = 0, theta).
    # They are NOT the rounded formula values (36453 vs 36451.6, 49417 vs 49414.9): libvmaf carries literals, tagged L in
    # the spec file; the model pins them to 1e-6 of the factor, i.e. 4e-5 relative
    r = 3.0 * 1080 * math.pi / 180.0
    a, k, f0 = M["SPEC_DWT79_A"][0], M["SPEC_DWT79_K"][0], M["SPEC_DWT79_F0"][0]
    g, amp = M["SPEC_DWT79_G"], M["SPEC_DWT79_AMP"]
    assert len(g) == 4 and len(amp) == 16
    for theta, lit, q in ((1, 36453, 21), (2, 49417, 23)):
        t = math.log10(2.0 * f0 * g[theta] / r)
        rf = 1.0 / (2.0 * a * 10.0 ** (k * t * t) / amp[theta])
        assert abs(lit / 2.0 ** q - rf) < 1e-6
    assert M["SPEC_ADM_S0_RF"] == [36453.0, 36453.0, 49417.0]
    assert M["SPEC_ADM_S0_RF_ROUND"] == [float(1 << (s - 1)) for s in (15, 15, 17)]
    assert M["SPEC_ADM_S0_RF_SHIFT"] == [15.0, 15.0, 17.0]


def test_log2_polynomial_approximates_log2_of_the_mantissa():
    c = M["SPEC_LOG2_POLY"]
    assert len(c) == 9 and c[-1] == 0.0
    t = np.linspace(0, 1, 2001)
    v = np.zeros_like(t)
    for ck in c:
        v = v * t + ck
    assert np.abs(v - np.log2(1 + t)).max() < 2e-4                                 # a degree-8 minimax-style fit of log2(1 + t)
