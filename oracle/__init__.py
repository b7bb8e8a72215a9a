"""CPU oracle for the VMAF hot path -- TEST INFRASTRUCTURE, not the product.

ctypes front-end over ``oracle/liboracle.so`` (built from ``oracle/vmaf_oracle.c`` and
``oracle/vmaf_float_oracle.c`` by ``oracle/Makefile``).  PARITY UNPINNED: see the header of
``vmaf_oracle.c``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; nothing under
``pqa2_b200/`` does.

Restates what the reference reaches through ``ffmpeg -lavfi libvmaf`` at
``app/vmaf_analyzer.py:373-419``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (idempotent)."""
    srcs = [os.path.join(_HERE, f) for f in ("vmaf_oracle.c", "vmaf_float_oracle.c", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "libvmaf_spec.h"))      # the constants both sides share
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i, pd, d = C.c_void_p, C.c_int, C.c_ssize_t, C.c_double
        L.orc_motion_blur.argtypes = [vp, i, i, i, pd, vp]
        L.orc_motion_blur.restype = None
        L.orc_motion_sad.argtypes = [vp, vp, i, i]
        L.orc_motion_sad.restype = C.c_uint64
        L.orc_motion_score.argtypes = [C.c_uint64, i, i]
        L.orc_motion_score.restype = d
        L.orc_vif.argtypes = [vp, vp, i, i, i, pd, d, vp, vp, vp, vp]
        L.orc_vif.restype = i
        L.orc_vif_finish.argtypes = [vp, vp, vp]
        L.orc_vif_finish.restype = None
        L.orc_vif_log2_table.restype = C.POINTER(C.c_uint16)
        L.orc_vif_filter.argtypes = [i]
        L.orc_vif_filter.restype = C.POINTER(C.c_uint16)
        L.orc_adm.argtypes = [vp, vp, i, i, i, pd, d, d, i, vp, vp, vp, vp, vp, vp]
        L.orc_adm.restype = i
        L.orc_adm_rfactor.argtypes = [i, d, i, vp]
        L.orc_adm_rfactor.restype = None
        L.orc_sse.argtypes = [vp, vp, i, i, i, pd]
        L.orc_sse.restype = C.c_uint64
        L.orc_psnr_from_sse.argtypes = [C.c_uint64, i, i, i]
        L.orc_psnr_from_sse.restype = d
        L.orc_ffssim_plane.argtypes = [vp, vp, i, i, i, pd]
        L.orc_ffssim_plane.restype = d
        L.orc_svr_predict.argtypes = [vp, i, vp, vp, vp, vp, i, d, d]
        L.orc_svr_predict.restype = d
        _lib = L
    return _lib


def _plane(a: np.ndarray, bpc: int) -> np.ndarray:
    want = np.uint8 if bpc == 8 else np.uint16
    a = np.ascontiguousarray(a)
    if a.dtype != want:
        raise TypeError(f"plane dtype {a.dtype} does not match bpc {bpc}")
    return a


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------
# integer motion
# ------------------------------------------------------------------------------------------
def motion_blur(luma: np.ndarray, bpc: int = 8) -> np.ndarray:
    luma = _plane(luma, bpc)
    h, w = luma.shape
    out = np.empty((h, w), np.uint16)
    lib().orc_motion_blur(_p(luma), bpc, w, h, luma.strides[0], _p(out))
    return out


def motion_sad(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.uint16)
    b = np.ascontiguousarray(b, np.uint16)
    h, w = a.shape
    return int(lib().orc_motion_sad(_p(a), _p(b), w, h))


def motion_score(sad: int, w: int, h: int) -> float:
    return float(lib().orc_motion_score(int(sad), w, h))


# ------------------------------------------------------------------------------------------
# integer VIF / ADM
# ------------------------------------------------------------------------------------------
VIF_ACC_NAMES = ("num_log", "den_log", "num_non_log", "den_non_log", "accum_x", "accum_x2", "num_accum_x")


def vif(ref: np.ndarray, dis: np.ndarray, bpc: int = 8, enhn_gain_limit: float = 100.0) -> dict:
    ref, dis = _plane(ref, bpc), _plane(dis, bpc)
    h, w = ref.shape
    assert dis.shape == ref.shape and dis.strides[0] == ref.strides[0]
    acc = np.zeros((4, 7), np.int64)
    num = np.zeros(4)
    den = np.zeros(4)
    score = np.zeros(4)
    rc = lib().orc_vif(_p(ref), _p(dis), bpc, w, h, ref.strides[0], float(enhn_gain_limit),
                       _p(acc), _p(num), _p(den), _p(score))
    if rc != 0:
        raise ValueError("oracle vif: unsupported size")
    return {"acc": acc, "num": num, "den": den, "score": score}


def vif_finish(acc7: np.ndarray):
    acc7 = np.ascontiguousarray(acc7, np.int64)
    n, d = C.c_double(), C.c_double()
    lib().orc_vif_finish(_p(acc7), C.byref(n), C.byref(d))
    return n.value, d.value


def adm(ref: np.ndarray, dis: np.ndarray, bpc: int = 8, enhn_gain_limit: float = 100.0,
        norm_view_dist: float = 3.0, ref_display_height: int = 1080, want_bands: bool = False) -> dict:
    ref, dis = _plane(ref, bpc), _plane(dis, bpc)
    h, w = ref.shape
    assert dis.shape == ref.shape and dis.strides[0] == ref.strides[0]
    cm = np.zeros((4, 3), np.int64)
    dn = np.zeros((4, 3), np.uint64)
    ns = np.zeros(4)
    ds = np.zeros(4)
    adm2 = C.c_double()
    bands = None
    if want_bands:
        bands = np.zeros((4, (h + 1) // 2, (w + 1) // 2), np.int32)
    rc = lib().orc_adm(_p(ref), _p(dis), bpc, w, h, ref.strides[0], float(enhn_gain_limit),
                       float(norm_view_dist), int(ref_display_height), _p(cm), _p(dn), _p(ns), _p(ds),
                       C.byref(adm2), _p(bands) if bands is not None else None)
    if rc != 0:
        raise ValueError("oracle adm: unsupported size")
    with np.errstate(divide="ignore", invalid="ignore"):
        scale_scores = ns / ds
    out = {"cm": cm, "den": dn, "num_scale": ns, "den_scale": ds, "adm2": adm2.value,
           "scale_scores": scale_scores}
    if bands is not None:
        out["ref_bands_s0"] = bands
    return out


def adm_rfactor(scale: int, norm_view_dist: float = 3.0, ref_display_height: int = 1080) -> np.ndarray:
    rf = np.zeros(3, np.float32)
    lib().orc_adm_rfactor(scale, norm_view_dist, ref_display_height, _p(rf))
    return rf


# ------------------------------------------------------------------------------------------
# PSNR / FFmpeg SSIM
# ------------------------------------------------------------------------------------------
def sse(a: np.ndarray, b: np.ndarray, bpc: int = 8) -> int:
    a, b = _plane(a, bpc), _plane(b, bpc)
    h, w = a.shape
    return int(lib().orc_sse(_p(a), _p(b), bpc, w, h, a.strides[0]))


def psnr_from_sse(sse_v: int, bpc: int, w: int, h: int) -> float:
    return float(lib().orc_psnr_from_sse(int(sse_v), bpc, w, h))


def ffssim_plane(a: np.ndarray, b: np.ndarray, bpc: int = 8) -> float:
    a, b = _plane(a, bpc), _plane(b, bpc)
    h, w = a.shape
    return float(lib().orc_ffssim_plane(_p(a), _p(b), bpc, w, h, a.strides[0]))


# ------------------------------------------------------------------------------------------
# SVR
# ------------------------------------------------------------------------------------------
def svr_predict(feat, slopes, intercepts, sv, coef, gamma: float, rho: float) -> float:
    """Un-clipped, un-transformed, de-normalised score (libvmaf predict.c order)."""
    feat = np.ascontiguousarray(feat, np.float64)
    slopes = np.ascontiguousarray(slopes, np.float64)
    intercepts = np.ascontiguousarray(intercepts, np.float64)
    sv = np.ascontiguousarray(sv, np.float64)
    coef = np.ascontiguousarray(coef, np.float64)
    n_sv, n_feat = sv.shape
    return float(lib().orc_svr_predict(_p(feat), n_feat, _p(slopes), _p(intercepts), _p(sv), _p(coef),
                                       n_sv, float(gamma), float(rho)))


# ------------------------------------------------------------------------------------------
# whole-frame convenience: the integer feature row libvmaf would log for one pair
# ------------------------------------------------------------------------------------------
def integer_features(ref: np.ndarray, dis: np.ndarray, bpc: int = 8, vif_egl: float = 100.0,
                     adm_egl: float = 100.0) -> dict:
    v = vif(ref, dis, bpc, vif_egl)
    a = adm(ref, dis, bpc, adm_egl)
    out = {"integer_adm2": a["adm2"]}
    for s in range(4):
        out[f"integer_adm_scale{s}"] = float(a["scale_scores"][s])
        out[f"integer_vif_scale{s}"] = float(v["score"][s])
    return out


# ------------------------------------------------------------------------------------------
# float extractors (vmaf_float_* models, float_ssim, float_ms_ssim) -- vmaf_float_oracle.c
# ------------------------------------------------------------------------------------------
_f_ready = False


def _flib():
    global _f_ready
    L = lib()
    if not _f_ready:
        vp, i, pd, d, f = C.c_void_p, C.c_int, C.c_ssize_t, C.c_double, C.c_float
        L.orc_f_picture_copy.argtypes = [vp, i, i, i, pd, f, vp]
        L.orc_f_picture_copy.restype = None
        L.orc_f_vif.argtypes = [vp, vp, i, i, d, vp, vp]
        L.orc_f_motion_blur.argtypes = [vp, i, i, vp]
        L.orc_f_motion_blur.restype = None
        L.orc_f_motion_sad.argtypes = [vp, vp, i, i]
        L.orc_f_motion_sad.restype = d
        L.orc_f_adm.argtypes = [vp, vp, i, i, d, d, i, vp, vp, vp, vp, vp]
        L.orc_f_ssim.argtypes = [vp, vp, i, i]
        L.orc_f_ssim.restype = d
        L.orc_f_ms_ssim.argtypes = [vp, vp, i, i, vp]
        L.orc_f_ms_ssim.restype = d
        L.orc_f_log2_approx.argtypes = [f]
        L.orc_f_log2_approx.restype = f
        L.orc_f_vif_filter.argtypes = [i]
        L.orc_f_vif_filter.restype = C.POINTER(C.c_float)
        _f_ready = True
    return L


def picture_copy(luma: np.ndarray, bpc: int, offset: float) -> np.ndarray:
    """libvmaf picture_copy(): float32 luma, (v / 2^(bpc-8)) + offset."""
    luma = _plane(luma, bpc)
    h, w = luma.shape
    out = np.empty((h, w), np.float32)
    _flib().orc_f_picture_copy(_p(luma), bpc, w, h, luma.strides[0], offset, _p(out))
    return out


def f_vif(ref_f: np.ndarray, dis_f: np.ndarray, enhn_gain_limit: float = 100.0) -> dict:
    h, w = ref_f.shape
    num, den = np.zeros(4), np.zeros(4)
    _flib().orc_f_vif(_p(np.ascontiguousarray(ref_f, np.float32)), _p(np.ascontiguousarray(dis_f, np.float32)), w, h,
                      float(enhn_gain_limit), _p(num), _p(den))
    return {"num": num, "den": den, "score": num / den}


def f_motion_blur(ref_f: np.ndarray) -> np.ndarray:
    h, w = ref_f.shape
    out = np.empty((h, w), np.float32)
    _flib().orc_f_motion_blur(_p(np.ascontiguousarray(ref_f, np.float32)), w, h, _p(out))
    return out


def f_motion_sad(a: np.ndarray, b: np.ndarray) -> float:
    h, w = a.shape
    return float(_flib().orc_f_motion_sad(_p(np.ascontiguousarray(a, np.float32)), _p(np.ascontiguousarray(b, np.float32)), w, h))


def f_adm(ref_f: np.ndarray, dis_f: np.ndarray, enhn_gain_limit: float = 100.0, norm_view_dist: float = 3.0,
          ref_display_height: int = 1080) -> dict:
    h, w = ref_f.shape
    ns, ds = np.zeros((4, 3)), np.zeros((4, 3))
    num, den = np.zeros(4), np.zeros(4)
    adm2 = C.c_double()
    _flib().orc_f_adm(_p(np.ascontiguousarray(ref_f, np.float32)), _p(np.ascontiguousarray(dis_f, np.float32)), w, h,
                      float(enhn_gain_limit), float(norm_view_dist), int(ref_display_height), _p(ns), _p(ds), _p(num),
                      _p(den), C.byref(adm2))
    return {"num_sum": ns, "den_sum": ds, "num_scale": num, "den_scale": den, "adm2": adm2.value,
            "scale_scores": num / den}


def f_ssim(ref0: np.ndarray, dis0: np.ndarray) -> float:
    """ref0/dis0: float luma with offset 0 (range [0, 255])."""
    h, w = ref0.shape
    return float(_flib().orc_f_ssim(_p(np.ascontiguousarray(ref0, np.float32)), _p(np.ascontiguousarray(dis0, np.float32)), w, h))


def f_ms_ssim(ref0: np.ndarray, dis0: np.ndarray):
    h, w = ref0.shape
    lcs = np.zeros(15)
    v = float(_flib().orc_f_ms_ssim(_p(np.ascontiguousarray(ref0, np.float32)), _p(np.ascontiguousarray(dis0, np.float32)),
                                    w, h, _p(lcs)))
    return v, lcs.reshape(5, 3)


def float_features(ref: np.ndarray, dis: np.ndarray, bpc: int = 8, prev_ref: np.ndarray | None = None,
                   vif_egl: float = 100.0, adm_egl: float = 100.0, psnr: bool = False, ssim: bool = False,
                   ms_ssim: bool = False) -> dict:
    """The float feature row libvmaf would log for one pair (motion needs the previous ref frame)."""
    rf, df = picture_copy(ref, bpc, -128.0), picture_copy(dis, bpc, -128.0)
    v = f_vif(rf, df, vif_egl)
    a = f_adm(rf, df, adm_egl)
    out = {"adm2": a["adm2"], "vif": v, "adm": a}
    for s in range(4):
        out[f"adm_scale{s}"] = float(a["scale_scores"][s])
        out[f"vif_scale{s}"] = float(v["score"][s])
    if prev_ref is not None:
        out["motion"] = f_motion_sad(f_motion_blur(rf), f_motion_blur(picture_copy(prev_ref, bpc, -128.0)))
    else:
        out["motion"] = 0.0
    if psnr:
        h, w = ref.shape
        out["psnr_y"] = psnr_from_sse(sse(ref, dis, bpc), bpc, w, h)
    if ssim or ms_ssim:
        r0, d0 = picture_copy(ref, bpc, 0.0), picture_copy(dis, bpc, 0.0)
        if ssim:
            out["float_ssim"] = f_ssim(r0, d0)
        if ms_ssim:
            out["float_ms_ssim"] = f_ms_ssim(r0, d0)[0]
    return out
