"""The drop-in boundary used the way INTEGRATION.md (Option B) shows a maintainer of the reference would bind it: raw
ctypes against libb200vmaf.so and include/b200vmaf.h only -- no pqa2_b200 Python wrapper on the path -- replacing the
ffmpeg child of app/vmaf_analyzer.py:411-455.  The result must equal what the engine (the same ABI behind
pqa2_b200.engine) produces for the same frames."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_raw_ctypes_binding_matches_engine():
    from pqa2_b200 import _lib, engine, model as M, synth            # only for the struct layout, the frames and the check
    lib = C.CDLL(os.path.join(ROOT, "pqa2_b200", "libb200vmaf.so"))

    class BvOpts(C.Structure):
        _fields_ = [("vif_enhn_gain_limit", C.c_double), ("adm_enhn_gain_limit", C.c_double),
                    ("adm_norm_view_dist", C.c_double), ("adm_ref_display_height", C.c_int),
                    ("batch_frames", C.c_int), ("reserved", C.c_int * 6)]

    lib.bv_create.restype = C.c_void_p
    lib.bv_create.argtypes = [C.c_int] * 5 + [C.c_uint, C.POINTER(BvOpts)]
    lib.bv_submit.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                              C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_uint]
    lib.bv_flush.argtypes = [C.c_void_p]
    lib.bv_fetch.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.bv_destroy.argtypes = [C.c_void_p]
    lib.bv_destroy.restype = None
    lib.bv_last_error.restype = C.c_char_p
    lib.bv_last_error.argtypes = [C.c_void_p]
    lib.bv_sizeof_frame_features.restype = C.c_size_t
    lib.bv_model_create.restype = C.c_void_p
    pd = C.POINTER(C.c_double)
    lib.bv_model_create.argtypes = [C.c_int, C.c_int, pd, pd, C.c_double, C.c_double, pd, pd, pd, C.c_int, pd, C.c_uint]
    lib.bv_predict.argtypes = [C.c_void_p, pd, C.c_int64, C.c_uint, pd]
    lib.bv_model_free.argtypes = [C.c_void_p]
    assert lib.bv_sizeof_frame_features() == C.sizeof(_lib.BvFrameFeatures)

    w, h, n = 352, 288, 5
    BV_FEAT_VMAF_INT, BV_FEAT_PSNR_Y, BV_FRAME_FIRST = 0x7, 0x8, 0x4
    opts = BvOpts(100.0, 100.0, 3.0, 1080, 0)
    ctx = lib.bv_create(0, w, h, 8, 420, BV_FEAT_VMAF_INT | BV_FEAT_PSNR_Y, C.byref(opts))
    assert ctx, lib.bv_last_error(None)
    frames = [synth.frame_pair(31, f, w, h, 8) for f in range(n)]

    def planes(fr):
        return ((C.c_void_p * 3)(*[a.ctypes.data for a in fr]), (C.c_size_t * 3)(*[a.strides[0] for a in fr]))

    for i, (ref, dis) in enumerate(frames):
        rp, rs = planes(ref)
        dp, ds = planes(dis)
        assert lib.bv_submit(ctx, i, rp, rs, dp, ds, BV_FRAME_FIRST if i == 0 else 0) == 0, lib.bv_last_error(ctx)
    assert lib.bv_flush(ctx) == 0
    feats = (_lib.BvFrameFeatures * n)()
    assert lib.bv_fetch(ctx, 0, n, feats) == 0
    lib.bv_destroy(ctx)

    # motion2[i] = min(motion[i], motion[i+1]); features in the model's order; SVR through bv_model_create / bv_predict
    mj = json.load(open(os.path.join(ROOT, "pqa2_b200", "models", "vmaf_v0.6.1.bvm.json")))
    mdl = M.resolve_model("vmaf_v0.6.1").main                                          # parsed libsvm text (dense SVs)
    motion = [0.0] + [feats[i].motion for i in range(1, n)]
    motion2 = [min(motion[i], motion[i + 1]) if i + 1 < n else motion[i] for i in range(n)]
    x = np.array([[feats[i].adm2, motion2[i]] + list(feats[i].vif_scale) for i in range(n)], np.float64)
    sv, coef = np.ascontiguousarray(mdl.sv, np.float64), np.ascontiguousarray(mdl.coef, np.float64)
    sl, ic = np.ascontiguousarray(mdl.slopes, np.float64), np.ascontiguousarray(mdl.intercepts, np.float64)
    clip = np.array(mdl.score_clip or [0.0, 100.0], np.float64)
    tp = np.zeros(3, np.float64)
    hm = lib.bv_model_create(6, sv.shape[0], sv.ctypes.data_as(pd), coef.ctypes.data_as(pd), float(mdl.gamma), float(mdl.rho),
                             sl.ctypes.data_as(pd), ic.ctypes.data_as(pd), clip.ctypes.data_as(pd), 1, tp.ctypes.data_as(pd), 0)
    assert hm and mj is not None
    out = np.empty(n, np.float64)
    assert lib.bv_predict(hm, x.ctypes.data_as(pd), n, 0, out.ctypes.data_as(pd)) == 0
    lib.bv_model_free(hm)

    class Clip(engine.FrameSource):
        width, height, bpc, chroma, nb_frames, fps = w, h, 8, 420, n, 30.0

        def read_into(self, i, r, d, luma_only):
            for p in range(1 if luma_only else 3):
                r[p][...] = frames[i][0][p]
                d[p][...] = frames[i][1][p]

    res = engine.analyze(Clip(), M.resolve_model("vmaf_v0.6.1"), engine.EngineOptions(psnr=True, svr_on_device=False))
    want = [fr["metrics"]["vmaf"] for fr in res["frames"]]
    assert out.tolist() == want
    assert [feats[i].psnr_y for i in range(n)] == [fr["metrics"]["psnr_y"] for fr in res["frames"]]
