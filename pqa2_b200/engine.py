"""Clip-level VMAF engine: frames -> per-frame libvmaf metric rows -> pooled report.

Does, on B200s, what the ffmpeg child does for the reference between ``Popen``
(``app/vmaf_analyzer.py:446-455``) and the JSON log it leaves behind (``:640-641``):
pairs frames, runs the extractors (CUDA, through the C ABI), applies libvmaf's temporal rule for
motion2, fuses features with the model's SVR, pools, and returns the log as a dict.

Multi-GPU: frames are split into contiguous chunks, one per device, each with a
one-frame lead-in that only feeds the motion state; per-frame raw ``motion`` comes back from every
shard and motion2 / SVR / pooling run on the host over the concatenated rows.  No collective."""
from __future__ import annotations

import math
import threading
import time
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from . import report
from .extractor import FeatureExtractor, pinned_empty
from .model import VmafModel
from .yuvio import EndOfClip


@dataclass
class EngineOptions:
    n_subsample: int = 1
    psnr: bool = False               # libvmaf `psnr=1`   -> psnr_y (FFmpeg's libvmaf filter maps the option to the `psnr`
                                     # extractor with enable_chroma=false: SURVEY.md Appendix A.8)
    psnr_chroma: bool = False        # `feature=name=psnr` without that flag -> psnr_y, psnr_cb, psnr_cr
    ssim: bool = False               # libvmaf `ssim=1`   -> float_ssim
    ms_ssim: bool = False            # libvmaf `ms_ssim=1`-> float_ms_ssim
    ffmpeg_psnr: bool = False        # FFmpeg `psnr` filter stats (all planes)
    ffmpeg_ssim: bool = False        # FFmpeg `ssim` filter stats (all planes)
    enable_transform: bool = False
    disable_clip: bool = False
    report_motion: bool = True       # libvmaf logs integer_motion next to integer_motion2
    devices: tuple = (0,)
    batch_frames: int = 0
    svr_on_device: bool = True
    extra_features: int = 0          # extra _lib.FEAT_* bits
    fast_float: bool = False         # float models only, opt-in: contracted / folded-tap stencils (bv_opts.fast_float)
    dynamic_chunk: int = 512         # frames per chunk when several contexts share a long clip (>= 4 chunks each); 0 = fixed shares
    contexts_per_device: int = 0     # contexts (frame shards side by side) on EACH GPU of `devices`; 0 = auto: 3 for a clip
                                     # long enough to be dealt in chunks (>= 1536 frames per GPU), else 1.  One context runs its ~25 kernels per
                                     # launch group back to back on one stream, and the small pyramid levels, the
                                     # reductions and every kernel's last wave leave SMs idle; a second and third
                                     # context's kernels fill them: +6-8 % frames/s on one B200, same bits
    reader_threads: int = 0          # file readers per shard (pinned-ring path); 0 = auto (cores / shards, 2 .. 16)
    float_motion: bool = False       # `feature=name=motion` (app/vmaf_analyzer.py:388-402): libvmaf's float motion
                                     # extractor next to the model's own features -> `motion`, `motion2` in the log


class FrameSource:
    """Protocol: width/height/bpc/chroma/nb_frames/fps + read_into(i, ref_planes, dis_planes)."""
    width: int
    height: int
    bpc: int
    chroma: int
    nb_frames: int
    fps: float = 30.0

    def open(self):            # per-thread handle (file descriptors are not shared between shards)
        return self

    def close(self):
        pass

    def read_into(self, i: int, ref_planes, dis_planes, luma_only: bool) -> None:
        raise NotImplementedError


class FileSource(FrameSource):
    """A reference / distorted pair of clip files.  Frames are read into a pinned ring by reader threads
    (``_Prefetcher``: 12 threads deliver ~36 GB/s from the page cache on the B200 box).  ``mapped=True`` instead maps raw
    clips (.y4m / planar .yuv) and registers the mapping with CUDA (``yuvio.MappedClip``), so the shards hand page-cache
    views straight to ``bv_submit`` with no CPU copy (39 GB/s, PCIe-bound) -- but cudaHostRegister pins pages at only
    ~6 GB/s (0.3 s for a 1.9 GB pair; tmpfs only, ext4 / overlay mappings are refused), so it pays off only for a source
    that is analysed several times (several models over one clip pair); it is opt-in."""

    def __init__(self, ref_info, dis_info, mapped: bool = False):
        from .yuvio import ClipReader
        self._ri, self._di = ref_info, dis_info
        self.width, self.height, self.bpc, self.chroma = ref_info.width, ref_info.height, ref_info.bpc, ref_info.chroma
        self.nb_frames = min(ref_info.nb_frames, dis_info.nb_frames)
        self.fps = dis_info.fps or ref_info.fps or 30.0
        self._r = self._d = None
        self._ClipReader = ClipReader
        self._want_mapped = mapped
        self._maps = None                     # (MappedClip ref, MappedClip dis) shared by every shard, or False
        self._lock = threading.Lock()
        self.sequential = getattr(ref_info, "decoder", "raw") != "raw" or getattr(dis_info, "decoder", "raw") != "raw"
        self.parallel_reads = not self.sequential     # raw files: any frame can be read by any reader thread

    def _mapped(self):
        with self._lock:
            if self._maps is None:
                self._maps = False
                if self._want_mapped and not self.sequential:
                    from .yuvio import MappedClip
                    r = MappedClip.open(self._ri)
                    d = MappedClip.open(self._di) if r is not None else None
                    if r is not None and d is not None:
                        self._maps = (r, d)
                    elif r is not None:
                        r.close()
            return self._maps

    @property
    def zero_copy(self) -> bool:
        return bool(self._mapped())

    def get(self, i, luma_only):
        r, d = self._maps
        return r.planes(i, luma_only), d.planes(i, luma_only)

    def open(self):
        if self._mapped():
            return self                       # views of the shared mappings: nothing per thread
        s = FileSource(self._ri, self._di, mapped=False)
        s._maps = False
        s._r, s._d = self._ClipReader(self._ri), self._ClipReader(self._di)
        return s

    def close(self):
        for r in (self._r, self._d):
            if r:
                r.close()
        self._r = self._d = None

    def release(self):
        """Drop the shared mappings (after the last analyze() on this source)."""
        with self._lock:
            if self._maps:
                for m in self._maps:
                    m.close()
            self._maps = None

    def read_into(self, i, ref_planes, dis_planes, luma_only):
        self._r.read_into(i, ref_planes, luma_only)
        self._d.read_into(i, dis_planes, luma_only)


class _Prefetcher:
    """Reader threads that fill a pinned ring ahead of the submitting thread (file reads release the GIL).  Ring slot
    ``o % len(ring)`` may be filled for ordinal ``o`` once ``release(limit)`` has promised that every ordinal below
    ``limit - len(ring)`` was uploaded."""

    def __init__(self, src: FrameSource, first_handle, ring, frame_ids, luma_only: bool, n_threads: int):
        self.ring, self.ids, self.luma_only = ring, frame_ids, luma_only
        self.cond = threading.Condition()
        self.next, self.limit, self.stop = 0, min(len(ring), len(frame_ids)), False
        self.filled = [False] * len(frame_ids)
        self.end_at, self.error = None, None
        self.handles = [first_handle] + [src.open() for _ in range(max(1, n_threads) - 1)]
        self.threads = [threading.Thread(target=self._work, args=(h,), daemon=True) for h in self.handles]
        for t in self.threads:
            t.start()

    def _work(self, handle):
        while True:
            with self.cond:
                while not self.stop and (self.next >= self.limit or self.next >= len(self.ids)):
                    if self.next >= len(self.ids):
                        return
                    self.cond.wait()
                if self.stop:
                    return
                o = self.next
                self.next += 1
            rp, dp = self.ring[o % len(self.ring)]
            try:
                handle.read_into(self.ids[o], rp, dp, self.luma_only)
            except EndOfClip:
                with self.cond:
                    self.end_at = o if self.end_at is None else min(self.end_at, o)
                    self.stop = True
                    self.cond.notify_all()
                return
            except Exception as e:            # noqa: BLE001  (re-raised on the submitting thread)
                with self.cond:
                    self.error = e
                    self.stop = True
                    self.cond.notify_all()
                return
            with self.cond:
                self.filled[o] = True
                self.cond.notify_all()

    def get(self, o: int):
        with self.cond:
            while not self.filled[o]:
                if self.error is not None:
                    raise self.error
                if self.end_at is not None and o >= self.end_at:
                    raise EndOfClip(f"clip ends at ordinal {self.end_at}")
                self.cond.wait()
        return self.ring[o % len(self.ring)]

    def release(self, limit: int):
        with self.cond:
            self.limit = max(self.limit, min(limit, len(self.ids)))
            self.cond.notify_all()

    def close(self, own_first: bool = False):
        with self.cond:
            self.stop = True
            self.cond.notify_all()
        for t in self.threads:
            t.join()
        for h in self.handles[0 if own_first else 1:]:
            h.close()


class SynthSource(FrameSource):
    """Procedural clip (pqa2_b200.synth) -- tests and bench."""

    def __init__(self, width, height, bpc=8, nb_frames=8, seed=1, chroma=420, q=4, strength=1, fps=30.0):
        self.width, self.height, self.bpc, self.chroma = width, height, bpc, chroma
        self.nb_frames, self.seed, self.q, self.strength, self.fps = nb_frames, seed, q, strength, fps

    def read_into(self, i, ref_planes, dis_planes, luma_only):
        from . import synth
        rp, dp = synth.frame_pair(self.seed, i, self.width, self.height, self.bpc, chroma=not luma_only,
                                  q=self.q, strength=self.strength)
        for k in range(len(rp)):
            ref_planes[k][...] = rp[k]
            dis_planes[k][...] = dp[k]


class PlaneList(list):
    """[Y, U, V] views of ONE pinned buffer (``flat``, uint8): a reader can fill a whole frame with a single read."""
    flat = None


def pinned_planes(shapes, dtype) -> PlaneList:
    item = np.dtype(dtype).itemsize
    sizes = [int(np.prod(s)) * item for s in shapes]
    flat = pinned_empty((sum(sizes),), np.uint8)
    out, off = PlaneList(), 0
    for s, n in zip(shapes, sizes):
        out.append(flat[off:off + n].view(dtype).reshape(s))
        off += n
    out.flat = flat
    return out


def _plane_shapes(src: FrameSource):
    w, h = src.width, src.height
    if src.chroma in (0, 400):
        return [(h, w)]
    cw = (w + 1) // 2 if src.chroma in (420, 422) else w
    ch = (h + 1) // 2 if src.chroma == 420 else h
    return [(h, w), (ch, cw), (ch, cw)]


def feature_mask(model: VmafModel, opt: EngineOptions) -> int:
    m = L.FEAT_VMAF_FLOAT if model.is_float else L.FEAT_VMAF_INT
    if opt.psnr or opt.psnr_chroma:
        m |= L.FEAT_PSNR_Y
    if opt.psnr_chroma:
        m |= L.FEAT_PSNR_UV                      # inert for luma-only sources
    if opt.ssim:
        m |= L.FEAT_FLOAT_SSIM
    if opt.ms_ssim:
        m |= L.FEAT_FLOAT_MS_SSIM
    if opt.ffmpeg_psnr:
        m |= L.FEAT_PSNR_Y | L.FEAT_PSNR_UV
    if opt.ffmpeg_ssim:
        m |= L.FEAT_FFSSIM
    if opt.float_motion:
        m |= L.FEAT_FLOAT_MOTION
    return m | opt.extra_features


def shard_ranges(n_frames: int, n_shards: int, weights=None):
    """Contiguous chunks [start, end) per shard.  ``weights`` (one positive number per shard, e.g. the
    frames/s each GPU sustained in a calibration pass) makes the chunk lengths proportional to them: on a box whose GPUs
    do not see the same host-to-device bandwidth, equal chunks leave the fast ones idle while the slow ones finish."""
    n_shards = max(1, min(n_shards, max(n_frames, 1)))
    if weights is None or len(weights) != n_shards or not all(w > 0 for w in weights):
        return [(g * n_frames // n_shards, (g + 1) * n_frames // n_shards) for g in range(n_shards)]
    total = float(sum(weights))
    cuts, acc = [0], 0.0
    for w in weights[:-1]:
        acc += w
        cuts.append(min(n_frames, max(cuts[-1], int(round(n_frames * acc / total)))))
    cuts.append(n_frames)
    return [(cuts[g], cuts[g + 1]) for g in range(n_shards)]


_FIRST_KICK = 8
_EARLY_MIN = 64            # rows worth building ahead of the drain (below that the tail is short anyway)
_EARLY_STEP = 512          # long clips: build the log entries of finished frames every so many submitted frames
_MIN_CHUNK = 128           # shortest chunk worth a lead-in frame and a pipeline fill of its own (four launch groups)


def _hand_over(fx, rows, early, handed: int, done: int, lead: int) -> int:
    """Rows of the ordinals [handed, done) -- complete on the GPU and read back already -- go into `rows`, and `early`
    is told how far the clip is known now.  Returns the new `handed`.  The submitting thread does this between two
    launch groups: it waits for a free ring slot most of the time anyway, and the GPU has up to three groups queued."""
    if done - max(handed, lead) < _EARLY_MIN:
        return handed
    out = np.ctypeslib.as_array(fx.fetch(handed, done - handed))
    skip = max(lead - handed, 0)
    rows.put(out["frame_index"][skip:], out[skip:])
    early(int(out["frame_index"][-1]) + 1)
    return done


class _Cancelled(Exception):
    pass


def _run_shard(src: FrameSource, model: VmafModel, opt: EngineOptions, device: int, start: int, end: int,
               mask: int, rows: list, progress, cancel: threading.Event, errors: list, ctx_holder: list,
               session: "Engine | None" = None, shard: int = 0, lead_in: bool | None = None, early=None):
    handle = None
    fx = None
    pre = None
    try:
        handle = src.open()
        luma_only = not (mask & (L.FEAT_PSNR_UV | L.FEAT_FFSSIM)) or src.chroma in (0, 400)
        shapes = _plane_shapes(src)[: 1 if luma_only else 3]
        dtype = np.uint8 if src.bpc == 8 else np.uint16
        # one context (and one pinned ring) per shard: two shards on the same GPU must never share a bv_ctx
        # (include/b200vmaf.h: one host thread per ctx)
        key = (device, shard, src.width, src.height, src.bpc, src.chroma if not luma_only else 0, mask,
               model.vif_enhn_gain_limit, model.adm_enhn_gain_limit, opt.batch_frames, bool(opt.fast_float))
        fx = session._extractor(key) if session is not None else None
        if fx is None:
            fx = FeatureExtractor(src.width, src.height, src.bpc, src.chroma if not luma_only else 0, mask, device,
                                  vif_enhn_gain_limit=model.vif_enhn_gain_limit,
                                  adm_enhn_gain_limit=model.adm_enhn_gain_limit, batch_frames=opt.batch_frames,
                                  fast_float=opt.fast_float)
            if session is not None:
                session._keep_extractor(key, fx)
        else:
            fx.reset()
        ctx_holder.append(fx)
        B = fx.batch_frames
        zero_copy = bool(getattr(handle, "zero_copy", False))
        ring = None
        if not zero_copy:
            ring = session._ring(key, shapes, dtype, 2 * B) if session is not None else \
                [(pinned_planes(shapes, dtype), pinned_planes(shapes, dtype)) for _ in range(2 * B)]
        lead = (1 if start > 0 else 0) if lead_in is None else (1 if lead_in and start > 0 else 0)
        ordinal = 0
        handed = 0                      # ordinals whose rows were already handed to `early`
        if not isinstance(rows, Rows):
            early = None
        ids = list(range(start - lead, end))
        if not zero_copy:
            # raw files: several readers (page-cache copies run beside each other); a decoder is one sequential stream
            n_readers = 1
            if getattr(src, "parallel_reads", False):
                import os
                # auto: three quarters of the host's cores shared out over the shards, 2 .. 12 readers each.  One reader
                # copies ~4 GB/s out of the page cache; measured on the 16-core B200 box (tools/ingest_probe.py): 4 / 8 /
                # 12 / 16 readers deliver 21 / 32 / 36 / 26 GB/s -- past 12 they fight the submitting thread for cores
                n_readers = opt.reader_threads or max(2, min(12, ((os.cpu_count() or 4) * 3 // 4) // max(1, len(opt.devices))))
            pre = _Prefetcher(src, handle, ring, ids, luma_only, n_readers)
        # pictures large enough for a reduced group size are upload-bound over PCIe: only there does a short last
        # launch shorten the call; at <= 1440p the GPU is the bottleneck and short groups would only run less efficiently
        tail_split = B < L.BV_MAX_BATCH
        for i in ids:
            if cancel.is_set():
                fx.cancel()
                raise _Cancelled()
            if zero_copy:
                rp, dp = handle.get(i, luma_only)     # caller-owned (pinned / registered) planes, valid until we return
            else:
                if ordinal and ordinal % B == 0:
                    fx.wait_uploads()       # everything submitted so far is on the GPU: those ring slots are free again
                    pre.release(ordinal + len(ring))
                try:
                    rp, dp = pre.get(ordinal)
                except EndOfClip:
                    # the container promised more frames than it decodes to (its frame count is an estimate): the
                    # clip ends here, as it does for ffmpeg + libvmaf, which stop at the shorter input
                    errors.append(("truncated", i))
                    break
            flags = 0
            if i < start:
                flags |= L.FRAME_LEAD_IN
            if ordinal == 0:
                flags |= L.FRAME_FIRST      # nothing precedes the first frame this context sees
            if opt.n_subsample > 1 and i % opt.n_subsample != 0:
                flags |= L.FRAME_SKIP_SPATIAL
            fx.submit(i, rp, dp, flags)
            ordinal += 1
            if early is not None and ordinal - handed >= _EARLY_STEP and ordinal % B == 0:
                handed = _hand_over(fx, rows, early, handed, min(int(fx.frames_done()), ordinal), lead)
            left = end - 1 - i              # frames still to submit
            if (ordinal == min(_FIRST_KICK, B // 2) and B > 1) or (tail_split and left >= 2 and left in (B // 2, B // 4)):
                # start the GPU on a short first group (the pipeline fills sooner) and split the tail into halving
                # groups: when uploads are the bottleneck (2160p over PCIe) the drain after the last upload is the
                # compute time of the LAST launch only
                fx.kick()
            if progress:
                progress(1)
        if early is not None and ordinal:
            # Up to three launch groups are still on the GPU when the last frame has been submitted (a few ms of work).
            # Start the last (partial) group now and hand every row that is already complete to `early`, which builds
            # those frames' log entries while the GPU drains -- instead of doing all of it after the drain.
            fx.kick()
            handed = _hand_over(fx, rows, early, handed, min(int(fx.frames_done()), ordinal), lead)
        fx.flush()
        if ordinal > handed:
            out = np.ctypeslib.as_array(fx.fetch(handed, ordinal - handed))     # structured view of the bv_frame_features records
            skip = max(lead - handed, 0)
            if isinstance(rows, Rows):
                rows.put(out["frame_index"][skip:], out[skip:])
            else:
                for k in range(skip, len(out)):
                    rows[int(out["frame_index"][k])] = RowView(out, k)
    except _Cancelled:
        errors.append(("cancelled", None))
    except Exception as e:            # surfaced by analyze(): the caller decides how to report
        errors.append(("error", e))
        cancel.set()                  # a failed shard stops its siblings instead of letting them run to the end
    finally:
        if pre is not None:
            pre.close()
        if fx is not None and session is None:
            fx.close()
        if handle is not None and handle is not src:
            handle.close()


class Engine:
    """A session that keeps one CUDA context (and one pinned staging ring) per (device, geometry, features)
    alive between clips.  The reference pays an ffmpeg process start per clip (app/vmaf_analyzer.py:446);
    a sweep of many clips (BASELINE.json configs[4]) reuses the contexts instead of re-allocating
    ~0.6 GB of device buffers each time."""

    def __init__(self):
        self._fx = {}
        self._rings = {}
        self._lock = threading.Lock()

    def _extractor(self, key):
        with self._lock:
            return self._fx.get(key)

    def _keep_extractor(self, key, fx):
        with self._lock:
            self._fx[key] = fx

    def _ring(self, key, shapes, dtype, n):
        with self._lock:
            r = self._rings.get(key)
            if r is None:
                r = self._rings[key] = [(pinned_planes(shapes, dtype), pinned_planes(shapes, dtype)) for _ in range(n)]
            return r

    def analyze(self, src, model, opt=None, progress_cb=None, cancel=None, frame_range=None):
        return analyze(src, model, opt, progress_cb, cancel, frame_range, session=self)

    def retain(self, width: int, height: int, bpc: int):
        """Free every context and ring that was built for another picture geometry (call between analyses, never
        during one): a long-lived session that sees many resolutions would otherwise keep 1-7 GB per GPU for each."""
        with self._lock:
            for key in [k for k in self._fx if k[2:5] != (width, height, bpc)]:
                self._fx.pop(key).close()
                self._rings.pop(key, None)

    def close(self):
        with self._lock:
            for fx in self._fx.values():
                fx.close()
            self._fx.clear()
            self._rings.clear()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


ROW_DTYPE = np.dtype(L.BvFrameFeatures)        # numpy view of bv_frame_features (include/b200vmaf.h)

_ROW_KEYS = {"valid": "valid_mask", "raw": "raw", "motion": "motion", "vif": "vif_scale", "adm2": "adm2",
             "adm_scale": "adm_scale", "adm_num": "adm_num", "adm_den": "adm_den", "vif_num": "vif_num",
             "vif_den": "vif_den", "ffssim": "ffssim", "f_motion": "f_motion", "f_vif": "f_vif_scale",
             "f_adm2": "f_adm2", "f_adm_scale": "f_adm_scale", "float_ssim": "float_ssim",
             "float_ms_ssim": "float_ms_ssim", "frame_index": "frame_index", "flags": "flags"}


class RowView:
    """One frame of a Rows block with the dict keys the host code and the callers use."""
    __slots__ = ("_a", "_i")

    def __init__(self, arr, i):
        self._a, self._i = arr, i

    def __getitem__(self, key):
        r = self._a[self._i]
        if key == "psnr":
            return [float(r["psnr_y"]), float(r["psnr_cb"]), float(r["psnr_cr"])]
        v = r[_ROW_KEYS[key]]
        if isinstance(v, np.ndarray):
            return v.tolist()
        return v.item()

    def keys(self):
        return list(_ROW_KEYS) + ["psnr"]


class Rows:
    """Per-frame feature rows of a clip as ONE structured array (bv_frame_features records), filled shard by
    shard.  Indexing yields dict-like RowViews; build_frames() reads whole columns."""

    def __init__(self, n: int = 0, arr=None):
        self.arr = arr if arr is not None else np.zeros(n, ROW_DTYPE)
        self.present = np.zeros(len(self.arr), bool) if arr is None else np.ones(len(self.arr), bool)

    def __len__(self):
        return len(self.arr)

    def __getitem__(self, i):
        if isinstance(i, slice):
            out = Rows(arr=self.arr[i])
            out.present = self.present[i]
            return out
        if not self.present[i]:
            return None
        return RowView(self.arr, i)

    def __iter__(self):
        return (self[i] for i in range(len(self.arr)))

    def put(self, frame_indices, records):
        self.arr[frame_indices] = records
        self.present[frame_indices] = True


def _egl_suffix(v: float) -> str:
    return "" if v == 100.0 else "_egl_%g" % v


def motion2_from_motion(motion: list) -> list:
    """libvmaf integer_motion.c / float_motion.c: motion2[i] = min(motion[i], motion[i+1]); the last
    frame keeps its own motion (flush); motion[0] = 0."""
    n = len(motion)
    return [min(motion[i], motion[i + 1]) if i + 1 < n else motion[i] for i in range(n)]


def percentile(sorted_vals, p: float) -> float:
    """Linear-interpolated percentile (libvmaf predict.c percentile())."""
    n = len(sorted_vals)
    if n == 1:
        return sorted_vals[0]
    pos = p / 100.0 * (n - 1)
    lo = int(math.floor(pos))
    hi = min(lo + 1, n - 1)
    return sorted_vals[lo] + (sorted_vals[hi] - sorted_vals[lo]) * (pos - lo)


def _pool_column(col: np.ndarray) -> dict:
    """report.pool() on a column: the same left-to-right double sums (np.cumsum is strictly sequential)."""
    n = len(col)
    if n == 0:
        return {"min": 0.0, "max": 0.0, "mean": 0.0, "harmonic_mean": 0.0}
    return {"min": float(col.min()), "max": float(col.max()), "mean": float(np.cumsum(col)[-1]) / n,
            "harmonic_mean": n / float(np.cumsum(1.0 / (col + 1.0))[-1]) - 1.0}


def _build_frames_block(rows: Rows, model: VmafModel, opt: EngineOptions, device: int | None,
                        pooled_out: dict | None = None, clip_start: bool = True, cols_out: dict | None = None) -> list:
    """build_frames() on a Rows block: whole columns at a time (the per-row dict path costs ~50 us/frame in
    Python, which at several thousand frames/s is as much as the GPU work itself).  ``clip_start=False``: the block
    continues a clip (its first row keeps its motion); ``cols_out`` receives the numeric columns of the scored frames
    (name -> (values, present mask or None)) so that a caller who builds a clip in two blocks can pool over both."""
    a = rows.arr
    is_f = model.is_float
    pre = "" if is_f else "integer_"
    vs, as_ = _egl_suffix(model.vif_enhn_gain_limit), _egl_suffix(model.adm_enhn_gain_limit)
    n = len(a)
    motion = np.array(a["f_motion"] if is_f else a["motion"], np.float64)
    if n and clip_start:
        motion[0] = 0.0
    motion2 = motion.copy()
    if n > 1:
        motion2[:-1] = np.minimum(motion[:-1], motion[1:])
    valid = a["valid_mask"]
    scored = np.nonzero((valid & (L.FEAT_FLOAT_VIF if is_f else L.FEAT_VIF)) != 0)[0]
    adm2 = np.asarray(a["f_adm2"] if is_f else a["adm2"])
    adm_s = np.asarray(a["f_adm_scale"] if is_f else a["adm_scale"])
    vif = np.asarray(a["f_vif_scale"] if is_f else a["vif_scale"])
    base = {"adm2": adm2, "motion2": motion2, "motion": motion}
    for k in range(4):
        base[f"vif_scale{k}"] = vif[:, k]
        base[f"adm_scale{k}"] = adm_s[:, k]
    keys = model.main.metric_keys
    feats = np.stack([base[k[len("integer_"):] if k.startswith("integer_") else k][scored] for k in keys], axis=1) \
        if len(scored) else np.zeros((0, len(keys)))
    cols = [(f"{pre}adm2{as_}", adm2)] + [(f"{pre}adm_scale{k}{as_}", adm_s[:, k]) for k in range(4)]
    cols.append((f"{pre}motion2", motion2))
    if opt.report_motion:
        cols.append((f"{pre}motion", motion))
    cols += [(f"{pre}vif_scale{k}{vs}", vif[:, k]) for k in range(4)]
    opt_cols = []
    if opt.psnr or opt.psnr_chroma or opt.ffmpeg_psnr:
        opt_cols.append(("psnr_y", a["psnr_y"], L.FEAT_PSNR_Y))
    if opt.psnr_chroma:
        # libvmaf's psnr extractor logs all three planes unless enable_chroma=false (integer_psnr.c) -- which is what
        # FFmpeg's `psnr=1` sets; the chroma columns are therefore opt-in
        opt_cols.append(("psnr_cb", a["psnr_cb"], L.FEAT_PSNR_UV))
        opt_cols.append(("psnr_cr", a["psnr_cr"], L.FEAT_PSNR_UV))
    if opt.float_motion and not is_f:
        fm = np.array(a["f_motion"], np.float64)
        if n and clip_start:
            fm[0] = 0.0
        fm2 = fm.copy()
        if n > 1:
            fm2[:-1] = np.minimum(fm[:-1], fm[1:])
        cols += [("motion2", fm2), ("motion", fm)]
    opt_cols.append(("float_ssim", a["float_ssim"], L.FEAT_FLOAT_SSIM))
    opt_cols.append(("float_ms_ssim", a["float_ms_ssim"], L.FEAT_FLOAT_MS_SSIM))
    names = [c[0] for c in cols]
    table = np.stack([np.asarray(c[1], np.float64)[scored] for c in cols], axis=1).tolist() if len(scored) else []
    extras = [(nm, np.asarray(col, np.float64)[scored].tolist(), ((valid[scored] & bit) != 0))
              for nm, col, bit in opt_cols]
    frames = []
    if len(scored):
        dev = device if opt.svr_on_device else None
        vmaf = model.main.predict(feats, opt.enable_transform, opt.disable_clip, device=dev).tolist()
        boots = None
        if model.bootstrap:
            boots = np.stack([b.predict(feats, False, True, device=dev) for b in model.bootstrap], axis=1)
        # the usual clip: every optional column is present on every scored frame (or on none) -> one dict per frame
        # straight from a row of values, keys in libvmaf's log order (model features, optional metrics, vmaf)
        uniform = [e for e in extras if e[2].all()]
        if boots is None and all(e[2].all() or not e[2].any() for e in extras):
            keys = names + [e[0] for e in uniform] + ["vmaf"]
            for i, row, ex, v in zip(scored.tolist(), table, zip(*[e[1] for e in uniform]) if uniform else [()] * len(table), vmaf):
                frames.append({"frameNum": i, "metrics": dict(zip(keys, row + list(ex) + [v]))})
        else:
            oks = [e[2].tolist() for e in extras]
            for j, i in enumerate(scored.tolist()):
                m = dict(zip(names, table[j]))
                for (nm, vals, _), ok in zip(extras, oks):
                    if ok[j]:
                        m[nm] = vals[j]
                m["vmaf"] = vmaf[j]
                if boots is not None:
                    m.update(_bootstrap_metrics(model, boots[j], opt))
                frames.append({"frameNum": i, "metrics": m})
        if (pooled_out is not None or cols_out is not None) and boots is None:
            # pooled_metrics straight from the columns (the per-frame dict walk of report.pooled_metrics costs ~1 ms
            # per 512 frames -- as much as a third of a launch group of GPU work)
            tab = np.asarray(table, np.float64)
            cols_ = {nm: (tab[:, k], None) for k, nm in enumerate(names)}
            for nm, vals, okm in extras:
                cols_[nm] = (np.asarray(vals, np.float64), okm)
            cols_["vmaf"] = (np.asarray(vmaf, np.float64), None)
            if cols_out is not None:
                cols_out.update(cols_)
            if pooled_out is not None:
                pooled_out["pooled"] = _pool_columns([cols_])
    return frames


def _pool_columns(parts: list) -> dict:
    """Pooled metrics over the concatenation of column sets (one per block of a clip, in clip order)."""
    pooled = {}
    for nm in parts[0]:
        vals = np.concatenate([p[nm][0] for p in parts])
        if parts[0][nm][1] is not None:
            ok = np.concatenate([p[nm][1] for p in parts])
            if not ok.any():
                continue
            vals = vals[ok]
        pooled[nm] = _pool_column(vals)
    return pooled


class _NoGc:
    """Two dicts per frame: with the cyclic collector on, a long clip triggers full collections in the middle of the build
    (measured: 10 ms .. 0.5 s for 3600 frames, run to run); nothing built there can be cyclic garbage."""

    def __enter__(self):
        import gc
        self._gc, self._was_on = gc, gc.isenabled()
        gc.disable()

    def __exit__(self, *a):
        if self._was_on:
            self._gc.enable()


def build_frames(rows, model: VmafModel, opt: EngineOptions, device: int | None, pooled_out: dict | None = None) -> list:
    """Per-frame feature rows -> libvmaf 'frames' list (metric names of SURVEY.md Appendix A.8)."""
    if isinstance(rows, Rows):
        with _NoGc():
            return _build_frames_block(rows, model, opt, device, pooled_out)
    is_f = model.is_float
    pre = "" if is_f else "integer_"
    vs, as_ = _egl_suffix(model.vif_enhn_gain_limit), _egl_suffix(model.adm_enhn_gain_limit)
    motion = [(r["f_motion"] if is_f else r["motion"]) for r in rows]
    if motion:
        motion[0] = 0.0
    motion2 = motion2_from_motion(motion)
    scored = [i for i, r in enumerate(rows) if r["valid"] & (L.FEAT_FLOAT_VIF if is_f else L.FEAT_VIF)]
    frames = []
    feats = np.zeros((len(scored), 6), np.float64)
    keys = model.main.metric_keys
    for n, i in enumerate(scored):
        r = rows[i]
        m = {}
        adm2 = r["f_adm2"] if is_f else r["adm2"]
        adm_s = r["f_adm_scale"] if is_f else r["adm_scale"]
        vif = r["f_vif"] if is_f else r["vif"]
        m[f"{pre}adm2{as_}"] = adm2
        for s in range(4):
            m[f"{pre}adm_scale{s}{as_}"] = adm_s[s]
        m[f"{pre}motion2"] = motion2[i]
        if opt.report_motion:
            m[f"{pre}motion"] = motion[i]
        for s in range(4):
            m[f"{pre}vif_scale{s}{vs}"] = vif[s]
        if r["valid"] & L.FEAT_PSNR_Y and (opt.psnr or opt.psnr_chroma or opt.ffmpeg_psnr):
            m["psnr_y"] = r["psnr"][0]
            if opt.psnr_chroma and r["valid"] & L.FEAT_PSNR_UV:
                m["psnr_cb"], m["psnr_cr"] = r["psnr"][1], r["psnr"][2]
        if r["valid"] & L.FEAT_FLOAT_SSIM:
            m["float_ssim"] = r["float_ssim"]
        if r["valid"] & L.FEAT_FLOAT_MS_SSIM:
            m["float_ms_ssim"] = r["float_ms_ssim"]
        base = {"adm2": adm2, "motion2": motion2[i], "motion": motion[i],
                "vif_scale0": vif[0], "vif_scale1": vif[1], "vif_scale2": vif[2], "vif_scale3": vif[3],
                "adm_scale0": adm_s[0], "adm_scale1": adm_s[1], "adm_scale2": adm_s[2], "adm_scale3": adm_s[3]}
        for j, k in enumerate(keys):
            feats[n, j] = base[k[len("integer_"):] if k.startswith("integer_") else k]
        frames.append({"frameNum": i, "metrics": m})
    if scored:
        dev = device if opt.svr_on_device else None
        vmaf = model.main.predict(feats, opt.enable_transform, opt.disable_clip, device=dev)
        boots = None
        if model.bootstrap:
            boots = np.stack([b.predict(feats, False, True, device=dev) for b in model.bootstrap], axis=1)
        for n, fr in enumerate(frames):
            fr["metrics"]["vmaf"] = float(vmaf[n])
            if boots is not None:
                fr["metrics"].update(_bootstrap_metrics(model, boots[n], opt))
    return frames


def _post(model: VmafModel, y: float, opt: EngineOptions) -> float:
    m = model.main
    if opt.enable_transform and m.transform:
        t = m.transform
        v = 0.0
        has = False
        if "p0" in t:
            v += t["p0"]; has = True
        if "p1" in t:
            v += t["p1"] * y; has = True
        if "p2" in t:
            v += t["p2"] * y * y; has = True
        out = v if has else y
        if t.get("out_lte_in") and out > y:
            out = y
        if t.get("out_gte_in") and out < y:
            out = y
        y = out
    if not opt.disable_clip and m.score_clip:
        y = min(max(y, m.score_clip[0]), m.score_clip[1])
    return y


def _bootstrap_metrics(model: VmafModel, scores: np.ndarray, opt: EngineOptions) -> dict:
    """Bagging / stddev / 95 % CI over the bootstrap models' un-clipped scores."""
    s = [float(x) for x in scores]
    n = len(s)
    mean = sum(s) / n
    std = math.sqrt(sum((x - mean) ** 2 for x in s) / n)
    srt = sorted(s)
    lo, hi = percentile(srt, 2.5), percentile(srt, 97.5)
    delta = 0.01
    slope = (_post(model, mean + delta, opt) - _post(model, mean - delta, opt)) / (2 * delta) \
        if opt.enable_transform else 1.0
    return {"vmaf_bagging": _post(model, mean, opt), "vmaf_stddev": std * slope,
            "vmaf_ci_p95_lo": _post(model, lo, opt), "vmaf_ci_p95_hi": _post(model, hi, opt)}


def analyze(src: FrameSource, model: VmafModel, opt: EngineOptions | None = None, progress_cb=None,
            cancel: threading.Event | None = None, frame_range: tuple | None = None,
            session: "Engine | None" = None) -> dict:
    """Scores a clip; returns the libvmaf log as a dict plus ``rows`` (raw per-frame features).

    Raises on engine errors; returns ``None`` if cancelled."""
    opt = opt or EngineOptions()
    cancel = cancel or threading.Event()
    first, last = frame_range or (0, src.nb_frames)
    n = last - first
    if n <= 0:
        raise ValueError("no frames to analyze")
    mask = feature_mask(model, opt)
    rows = Rows(src.nb_frames)
    devices = list(opt.devices) or [0]
    chunk = opt.dynamic_chunk
    sequential = bool(getattr(src, "sequential", False))
    per_dev = opt.contexts_per_device
    if per_dev <= 0:
        per_dev = 3 if (chunk > 0 and not sequential and n >= 4 * 3 * len(devices) * _MIN_CHUNK) else 1
    if per_dev > 1 and not sequential:
        devices = [d for d in devices for _ in range(per_dev)]          # contexts of one GPU take neighbouring chunks
        if chunk > 0:                   # at least four chunks per context, so that their fills and drains interleave
            chunk = min(chunk, max(_MIN_CHUNK, n // (4 * len(devices))))
    ranges = [(first + a, first + b) for a, b in shard_ranges(n, len(devices))]
    if len(devices) != len(opt.devices):
        from dataclasses import replace
        opt = replace(opt, devices=tuple(devices))          # the shards share the host's reader threads by this count
    done = [0]
    lock = threading.Lock()

    def progress(k):
        if progress_cb:
            with lock:
                done[0] += k
                d = done[0]
            progress_cb(d, n)

    errors: list = []
    holders: list = []
    t0 = time.perf_counter()
    threads = []
    head = {"frames": [], "cols": [], "upto": first}

    def build_block(hi, final):
        """Log entries of the frames [head.upto, hi) from rows that are complete (frame hi - 1 waits for its successor's
        motion unless the clip ends there).  Called on the shard thread while the GPU works on later frames (`build_early`)
        and once more after the last row has arrived."""
        s0 = head["upto"]
        if hi - s0 < (1 if final else 2):
            return
        cols: dict = {}
        with _NoGc():
            frs = _build_frames_block(rows[s0:hi], model, opt, devices[0], None, s0 == first, cols)
        keep = len(frs) if final else sum(1 for fr in frs if fr["frameNum"] < hi - 1 - s0)
        for fr in frs[:keep]:
            fr["frameNum"] += s0
        head["frames"] += frs[:keep]
        if cols:
            head["cols"].append({nm: (v[:keep], None if ok is None else ok[:keep]) for nm, (v, ok) in cols.items()})
        head["upto"] = hi if final else hi - 1

    def build_early(hi):
        build_block(hi, False)

    # Long clips on several GPUs: the devices pull chunks of `dynamic_chunk` frames from a shared counter instead of
    # each getting one fixed share.  The GPUs of one box do not see the same host-to-device bandwidth under load (four
    # of the eight B200s share a host bridge: ~24 vs ~36 GB/s each, tools/h2d_concurrent.py), and with equal shares the
    # fast ones idle while the slow ones finish.  Every chunk after the first carries its own lead-in frame, so the
    # results are the same bits; sequential sources (container decode) keep their single shard.
    own_session = None
    if len(devices) > 1 and chunk > 0 and n >= 4 * len(devices) * chunk and not sequential:
        if session is None:
            session = own_session = Engine()            # contexts must survive from chunk to chunk
        nxt = [first]
        finished: dict = {}                             # chunk start -> chunk end, for the chunks whose rows have arrived
        arrived = threading.Condition(lock)

        def worker(k, dev):
            while not cancel.is_set():
                with lock:
                    a = nxt[0]
                    nxt[0] = b = min(last, a + chunk)
                if a >= last:
                    return
                n_err = len(errors)
                _run_shard(src, model, opt, dev, a, b, mask, rows, progress, cancel, errors, holders, session, k, a > first)
                with arrived:
                    if len(errors) == n_err:
                        finished[a] = b
                    arrived.notify_all()

        for k, dev in enumerate(devices):
            th = threading.Thread(target=worker, args=(k, dev), daemon=True)
            th.start()
            threads.append(th)
        # this thread has nothing to do but wait: it builds the log entries of the clip's finished prefix meanwhile
        prefix = first
        try:
            while any(th.is_alive() for th in threads) and not model.bootstrap:
                with arrived:
                    arrived.wait(0.02)
                    while prefix in finished:
                        prefix = finished.pop(prefix)
                if prefix - head["upto"] >= _EARLY_STEP and not errors:
                    build_block(prefix, False)
        except Exception as e:            # noqa: BLE001  (e.g. the SVR call failed): stop the workers, then report it
            errors.append(("error", e))
            cancel.set()
    else:
        live = [(k, dev, a, b) for k, (dev, (a, b)) in enumerate(zip(devices, ranges)) if b > a]
        for k, dev, a, b in live:
            # a frame_range is scored as libvmaf would score the trimmed clip: its first frame has no predecessor
            # (motion = 0), so only the shards after the first one get a lead-in frame
            th = threading.Thread(target=_run_shard, args=(src, model, opt, dev, a, b, mask, rows, progress, cancel,
                                                           errors, holders, session, k, a > first,
                                                           build_early if len(live) == 1 and not model.bootstrap else None),
                                  daemon=True)
            th.start()
            threads.append(th)
    for th in threads:
        th.join()
    if own_session is not None:
        own_session.close()
    for k, e in errors:               # an engine error wins over the cancellation it triggered in the sibling shards
        if k == "error":
            raise e
    if any(k == "cancelled" for k, _ in errors) or cancel.is_set():
        return None
    cut = [e for k, e in errors if k == "truncated"]
    if cut:
        last = min(cut)
        n = last - first
        if n <= 0:
            raise EOFError("no frame could be decoded")
    rows_used = rows[first:last]
    if head["upto"] > last:             # a truncated clip that ends before what was built ahead: start over
        head.update(frames=[], cols=[], upto=first)
    build_block(last, True)
    frames = head["frames"]
    pooled_out: dict = {}
    if head["cols"] and not model.bootstrap:
        pooled_out["pooled"] = _pool_columns(head["cols"])
    dt = time.perf_counter() - t0
    pooled = pooled_out.get("pooled") or report.pooled_metrics(frames)
    return {"version": report.VERSION, "fps": n / dt if dt > 0 else 0.0, "frames": frames,
            "pooled_metrics": pooled, "aggregate_metrics": {}, "rows": rows_used,
            "model": model.name, "elapsed_s": dt, "n_frames": n}


def analyze_batch(clips: list, model: VmafModel, opt: EngineOptions | None = None, devices=None,
                  progress_cb=None, concurrency: int = 3) -> list:
    """Many clips over many GPUs (BASELINE.json configs[4]: a sweep of 64 1080p clip pairs on 8 B200).

    Whole clips are the unit here -- no lead-in frames, no cross-GPU state: every GPU has ``concurrency`` workers, each
    with its own session (one context, reused for every clip of the same geometry), which take the next clip from a
    shared list; three clips in flight per GPU fill the few ms a single clip leaves idle while its pipeline fills and while
    its last launch groups drain and are scored.  Returns one libvmaf log dict per clip, in input order; a clip that
    fails yields ``{"error": str}`` in its slot (the reference's per-clip error convention)."""
    opt = opt or EngineOptions()
    devices = list(devices if devices is not None else opt.devices) or [0]
    results: list = [None] * len(clips)
    lock = threading.Lock()
    done = [0]
    nxt = [0]

    def worker(dev: int):
        from dataclasses import replace
        o = replace(opt, devices=(dev,))
        with Engine() as sess:
            while True:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= len(clips):
                    return
                try:
                    results[k] = sess.analyze(clips[k], model, o)
                except Exception as e:            # noqa: BLE001
                    results[k] = {"error": str(e)}
                if progress_cb:
                    with lock:
                        done[0] += 1
                        d = done[0]
                    progress_cb(d, len(clips))

    per_dev = max(1, int(concurrency))
    threads = [threading.Thread(target=worker, args=(d,), daemon=True) for _ in range(per_dev) for d in devices]
    threads = threads[:max(1, len(clips))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return results
