"""ctypes binding of ``libb200vmaf.so`` (C ABI: ``include/b200vmaf.h``).

This is the thin layer the north star asks for between the Python host code and the sm_100a
kernels.  It replaces the ``subprocess.Popen([ffmpeg, ..., "-lavfi", "libvmaf=..."])`` boundary of
the reference (``app/vmaf_analyzer.py:411-455``).  There is no CPU fallback: if the library is
missing it is built with nvcc, and if that fails the import raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200VMAF_LIB: load another build of the library (kernel experiments); never set in normal use
LIB_PATH = os.environ.get("B200VMAF_LIB") or os.path.join(_HERE, "libb200vmaf.so")

BV_RAW_WORDS = 64
BV_MAX_BATCH = 32

FEAT_MOTION = 0x001
FEAT_VIF = 0x002
FEAT_ADM = 0x004
FEAT_PSNR_Y = 0x008
FEAT_PSNR_UV = 0x010
FEAT_FFSSIM = 0x020
FEAT_FLOAT_VIF = 0x040
FEAT_FLOAT_ADM = 0x080
FEAT_FLOAT_MOTION = 0x100
FEAT_FLOAT_SSIM = 0x200
FEAT_FLOAT_MS_SSIM = 0x400
FEAT_VMAF_INT = FEAT_MOTION | FEAT_VIF | FEAT_ADM
FEAT_VMAF_FLOAT = FEAT_FLOAT_MOTION | FEAT_FLOAT_VIF | FEAT_FLOAT_ADM

FRAME_LEAD_IN = 0x1
FRAME_SKIP_SPATIAL = 0x2
FRAME_FIRST = 0x4

MODEL_ENABLE_TRANSFORM = 0x1
MODEL_DISABLE_CLIP = 0x2

RAW_SAD, RAW_VIF, RAW_ADM_CM, RAW_ADM_DEN, RAW_SSE = 0, 1, 29, 41, 53

ERR_ARG, ERR_CUDA, ERR_CANCELLED, ERR_ORDER, ERR_UNSUPPORTED = -1, -2, -3, -4, -5


class BvOpts(C.Structure):
    _fields_ = [("vif_enhn_gain_limit", C.c_double), ("adm_enhn_gain_limit", C.c_double),
                ("adm_norm_view_dist", C.c_double), ("adm_ref_display_height", C.c_int),
                ("batch_frames", C.c_int), ("fast_float", C.c_int), ("reserved", C.c_int * 5)]


class BvFrameFeatures(C.Structure):
    _fields_ = [("frame_index", C.c_int64), ("flags", C.c_uint32), ("valid_mask", C.c_uint32),
                ("raw", C.c_int64 * BV_RAW_WORDS),
                ("motion", C.c_double),
                ("vif_num", C.c_double * 4), ("vif_den", C.c_double * 4), ("vif_scale", C.c_double * 4),
                ("adm_num", C.c_double * 4), ("adm_den", C.c_double * 4), ("adm_scale", C.c_double * 4),
                ("adm2", C.c_double),
                ("psnr_y", C.c_double), ("psnr_cb", C.c_double), ("psnr_cr", C.c_double),
                ("ffssim", C.c_double * 3),
                ("f_motion", C.c_double),
                ("f_vif_num", C.c_double * 4), ("f_vif_den", C.c_double * 4), ("f_vif_scale", C.c_double * 4),
                ("f_adm_num", C.c_double * 4), ("f_adm_den", C.c_double * 4), ("f_adm_scale", C.c_double * 4),
                ("f_adm2", C.c_double),
                ("float_ssim", C.c_double), ("float_ms_ssim", C.c_double)]


EXPORTS = (
    "bv_abi_version", "bv_device_count", "bv_create", "bv_destroy", "bv_last_error", "bv_pinned_alloc",
    "bv_pinned_free", "bv_host_register", "bv_host_unregister", "bv_device_alloc", "bv_device_free", "bv_device_upload", "bv_sizeof_frame_features",
    "bv_submit", "bv_submit_device", "bv_wait_uploads", "bv_flush", "bv_frames_done", "bv_fetch", "bv_cancel",
    "bv_reset", "bv_kick", "bv_batch_frames",
    "bv_kernel_launches", "bv_set_profiling", "bv_kernel_slots", "bv_kernel_name", "bv_kernel_ms", "bv_kernel_count",
    "bv_timer_mark", "bv_timer_elapsed_ms",
    "bv_model_create", "bv_model_free", "bv_predict", "bv_predict_device", "bv_luma_stats_device",
)

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the CUDA library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and not os.environ.get("B200VMAF_LIB"):
        from . import build as _build
        try:
            _build.build()
        except Exception:
            if not os.path.exists(LIB_PATH):
                raise
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing and could not be built; this engine has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i, u, i64, d, sz = C.c_void_p, C.c_int, C.c_uint, C.c_int64, C.c_double, C.c_size_t
    pp = C.POINTER(C.c_void_p)
    psz = C.POINTER(C.c_size_t)
    pd = C.POINTER(C.c_double)
    L.bv_abi_version.restype = i
    L.bv_device_count.restype = i
    L.bv_create.argtypes = [i, i, i, i, i, u, C.POINTER(BvOpts)]
    L.bv_create.restype = vp
    L.bv_destroy.argtypes = [vp]
    L.bv_destroy.restype = None
    L.bv_last_error.argtypes = [vp]
    L.bv_last_error.restype = C.c_char_p
    L.bv_pinned_alloc.argtypes = [pp, sz]
    L.bv_pinned_free.argtypes = [vp]
    L.bv_host_register.argtypes = [vp, sz, i]
    L.bv_host_unregister.argtypes = [vp]
    L.bv_device_alloc.argtypes = [i, pp, sz]
    L.bv_device_free.argtypes = [i, vp]
    L.bv_device_upload.argtypes = [i, vp, vp, sz]
    L.bv_sizeof_frame_features.restype = sz
    L.bv_submit.argtypes = [vp, i64, pp, psz, pp, psz, u]
    L.bv_submit_device.argtypes = [vp, i64, pp, psz, pp, psz, u]
    L.bv_wait_uploads.argtypes = [vp]
    L.bv_flush.argtypes = [vp]
    L.bv_frames_done.argtypes = [vp]
    L.bv_frames_done.restype = i64
    L.bv_fetch.argtypes = [vp, i64, i64, C.POINTER(BvFrameFeatures)]
    L.bv_cancel.argtypes = [vp]
    L.bv_reset.argtypes = [vp]
    L.bv_kick.argtypes = [vp]
    L.bv_batch_frames.argtypes = [vp]
    L.bv_batch_frames.restype = i
    L.bv_kernel_launches.argtypes = [vp]
    L.bv_kernel_launches.restype = i64
    L.bv_set_profiling.argtypes = [vp, i]
    L.bv_kernel_slots.restype = i
    L.bv_kernel_name.argtypes = [i]
    L.bv_kernel_name.restype = C.c_char_p
    L.bv_kernel_ms.argtypes = [vp, i, i]
    L.bv_kernel_ms.restype = d
    L.bv_kernel_count.argtypes = [vp, i, i]
    L.bv_kernel_count.restype = d
    L.bv_timer_mark.argtypes = [vp, i]
    L.bv_timer_elapsed_ms.argtypes = [vp]
    L.bv_timer_elapsed_ms.restype = d
    L.bv_model_create.argtypes = [i, i, pd, pd, d, d, pd, pd, pd, i, pd, u]
    L.bv_model_create.restype = vp
    L.bv_model_free.argtypes = [vp]
    L.bv_model_free.restype = None
    L.bv_predict.argtypes = [vp, pd, i64, u, pd]
    L.bv_predict_device.argtypes = [vp, i, pd, i64, u, pd]
    L.bv_luma_stats_device.argtypes = [i, vp, sz, sz, i, i, i, i, C.POINTER(C.c_uint), C.POINTER(C.c_uint64)]
    if L.bv_sizeof_frame_features() != C.sizeof(BvFrameFeatures):
        raise RuntimeError("bv_frame_features ABI mismatch between libb200vmaf.so and the ctypes binding")
    _lib = L
    return L
