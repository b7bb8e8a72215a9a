"""Host-side driver of one ``bv_ctx`` (one GPU): frames in, per-frame feature rows out.

Replaces what happens inside the ffmpeg child of the reference between ``vmaf_read_pictures`` and
the feature collector (``app/vmaf_analyzer.py:417``)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib as L


class BvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200vmaf error {code}: {msg}")
        self.code = code


def _planes_arg(planes: Sequence, n: int):
    ptrs = (C.c_void_p * 3)()
    strides = (C.c_size_t * 3)()
    for k in range(3):
        if k < len(planes) and planes[k] is not None and k < n:
            p = planes[k]
            if isinstance(p, np.ndarray):
                ptrs[k] = p.ctypes.data
                strides[k] = p.strides[0]
            else:                                 # (device_ptr, pitch)
                ptrs[k] = int(p[0])
                strides[k] = int(p[1])
        else:
            ptrs[k] = None
            strides[k] = 0
    return ptrs, strides


class FeatureExtractor:
    """One GPU context.  Frames must be submitted in order; results are fetched by submission
    ordinal.  ``features`` is a mask of ``_lib.FEAT_*``."""

    def __init__(self, width: int, height: int, bpc: int = 8, chroma: int = 420,
                 features: int = L.FEAT_VMAF_INT, device: int = 0, vif_enhn_gain_limit: float = 100.0,
                 adm_enhn_gain_limit: float = 100.0, adm_norm_view_dist: float = 3.0,
                 adm_ref_display_height: int = 1080, batch_frames: int = 0, fast_float: bool = False):
        self.lib = L.load()
        self.width, self.height, self.bpc, self.chroma = width, height, bpc, chroma
        self.features, self.device = features, device
        o = L.BvOpts()
        o.vif_enhn_gain_limit = vif_enhn_gain_limit
        o.adm_enhn_gain_limit = adm_enhn_gain_limit
        o.adm_norm_view_dist = adm_norm_view_dist
        o.adm_ref_display_height = adm_ref_display_height
        o.batch_frames = batch_frames
        o.fast_float = 1 if fast_float else 0
        self._ctx = self.lib.bv_create(device, width, height, bpc, chroma, features, C.byref(o))
        if not self._ctx:
            raise BvError(L.ERR_CUDA, (self.lib.bv_last_error(None) or b"bv_create failed").decode())
        self._keep = []
        self.submitted = 0
        self._dtype = np.uint8 if bpc == 8 else np.uint16
        self._nplanes = 3 if (features & (L.FEAT_PSNR_UV | L.FEAT_FFSSIM)) and chroma not in (0, 400) else 1

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.bv_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise BvError(rc, (self.lib.bv_last_error(self._ctx) or b"").decode())

    # -- submission ----------------------------------------------------------------------------
    def submit(self, frame_index: int, ref_planes: Sequence[np.ndarray], dis_planes: Sequence[np.ndarray],
               flags: int = 0):
        """Host planes ([Y] or [Y, U, V]); copies are asynchronous, arrays are kept alive until
        ``wait_uploads``/``flush``."""
        for p in list(ref_planes)[: self._nplanes] + list(dis_planes)[: self._nplanes]:
            if p.dtype != self._dtype or p.strides[1] != p.itemsize:
                raise TypeError(f"plane must be {self._dtype} with unit column stride")
        rp, rs = _planes_arg(ref_planes, self._nplanes)
        dp, ds = _planes_arg(dis_planes, self._nplanes)
        self._keep.append((ref_planes, dis_planes))
        self._check(self.lib.bv_submit(self._ctx, frame_index, rp, rs, dp, ds, flags))
        self.submitted += 1

    def submit_device(self, frame_index: int, ref_planes, dis_planes, flags: int = 0):
        """Planes as (device_ptr, pitch_bytes) tuples, resident on this context's GPU."""
        rp, rs = _planes_arg(ref_planes, self._nplanes)
        dp, ds = _planes_arg(dis_planes, self._nplanes)
        self._check(self.lib.bv_submit_device(self._ctx, frame_index, rp, rs, dp, ds, flags))
        self.submitted += 1

    def wait_uploads(self):
        self._check(self.lib.bv_wait_uploads(self._ctx))
        self._keep.clear()

    def flush(self):
        self._check(self.lib.bv_flush(self._ctx))
        self._keep.clear()

    def kick(self):
        """Launch the frames submitted so far (a partial group) without waiting for them."""
        self._check(self.lib.bv_kick(self._ctx))

    def cancel(self):
        self.lib.bv_cancel(self._ctx)

    def reset(self):
        """Ready for the next clip of the same geometry (drops stored results and the motion state)."""
        self._check(self.lib.bv_reset(self._ctx))
        self._keep.clear()
        self.submitted = 0

    # -- results -------------------------------------------------------------------------------
    def fetch(self, first: int = 0, count: int | None = None):
        if count is None:
            count = self.submitted - first
        arr = (L.BvFrameFeatures * count)()
        if count:
            self._check(self.lib.bv_fetch(self._ctx, first, count, arr))
        return arr

    def frames_done(self) -> int:
        return int(self.lib.bv_frames_done(self._ctx))

    @property
    def batch_frames(self) -> int:
        """Frame pairs per launch group of this context (the auto choice when none was requested)."""
        return int(self.lib.bv_batch_frames(self._ctx))

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.bv_kernel_launches(self._ctx))

    def set_profiling(self, on: bool):
        self.lib.bv_set_profiling(self._ctx, 1 if on else 0)

    def kernel_profile(self, reset: bool = True) -> dict:
        """{kernel name: (total ms, launches timed)} since the last reset (needs set_profiling(True))."""
        out = {}
        for k in range(self.lib.bv_kernel_slots()):
            nm = self.lib.bv_kernel_name(k)
            if not nm:
                continue
            cnt = self.lib.bv_kernel_count(self._ctx, k, 1 if reset else 0)
            ms = self.lib.bv_kernel_ms(self._ctx, k, 1 if reset else 0)
            if cnt > 0:
                out[nm.decode()] = (float(ms), int(cnt))
        return out

    def timer_mark(self, which: int):
        self._check(self.lib.bv_timer_mark(self._ctx, which))

    def timer_elapsed_ms(self) -> float:
        return float(self.lib.bv_timer_elapsed_ms(self._ctx))


class DeviceBuffer:
    """cudaMalloc'd bytes on one GPU (resident clips for the bench and the tests)."""

    def __init__(self, nbytes: int, device: int = 0):
        self.lib = L.load()
        self.device, self.nbytes = device, nbytes
        p = C.c_void_p()
        rc = self.lib.bv_device_alloc(device, C.byref(p), nbytes)
        if rc != 0:
            raise BvError(rc, "cudaMalloc failed")
        self.ptr = p.value

    def upload(self, offset: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        rc = self.lib.bv_device_upload(self.device, self.ptr + offset, arr.ctypes.data, arr.nbytes)
        if rc != 0:
            raise BvError(rc, "cudaMemcpy H2D failed")

    def free(self):
        if self.ptr:
            self.lib.bv_device_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array backed by cudaHostAlloc memory (freed when the array's base is collected)."""
    lib = L.load()
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = lib.bv_pinned_alloc(C.byref(p), max(nbytes, 1))
    if rc != 0:
        raise BvError(rc, "cudaHostAlloc failed")

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib.bv_pinned_free(self.ptr)
            except Exception:
                pass

    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
    buf._owner = _Owner(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
