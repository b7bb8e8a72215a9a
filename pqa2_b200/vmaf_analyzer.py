"""Drop-in ``VMAFAnalyzer``: the reference's call surface (``app/vmaf_analyzer.py:18-616``) on the
B200 engine instead of an ``ffmpeg -lavfi libvmaf`` child process.

Same constructor, setters, ``analyze_videos`` / ``terminate_analysis`` / ``get_video_metadata``
methods, same four signals, same error convention (never raise to the caller: log, emit
``error_occurred(str)``, return ``None`` -- reference ``:259-269``, ``:520-524``, ``:603-609``), same
output files (``<test>_<ts>_vmaf.json`` in libvmaf's JSON layout, ``_psnr.txt`` / ``_ssim.txt`` in FFmpeg's
stats-file layout; reference
``:304-311``) and the same results dict (reference ``:919-932``).

Differences, all forced by the environment:
* PyQt5 is optional.  Without it the four signals are plain objects with ``connect`` / ``emit``.
* Inputs are raw planar video (``.y4m`` or headerless ``.yuv`` with a ``_WxH`` name hint) or, when cv2 is
  installed, containers (``.mp4`` ...) decoded by its bundled libavformat / libavcodec (``avdec``: Y, U, V as ffmpeg
  would deliver them; through ``cv2.VideoCapture`` -- luma only -- if those libraries cannot be driven directly).
* ``threads`` (libvmaf ``n_threads``) is accepted and ignored: the work runs on the GPUs in
  ``devices`` (default: every visible B200), frame-sharded with a one-frame lead-in per shard.
"""
from __future__ import annotations

import logging
import os
import threading
from datetime import datetime

from . import _lib as L
from . import engine, report
from . import model as M
from . import yuvio

logger = logging.getLogger(__name__)

try:                                        # pragma: no cover - PyQt5 is absent in the build image
    from PyQt5.QtCore import QObject, pyqtSignal
    _HAVE_QT = True
except Exception:                           # noqa: BLE001
    _HAVE_QT = False

    class QObject:                          # type: ignore[no-redef]
        def __init__(self, *a, **k):
            pass

    class _BoundSignal:
        def __init__(self):
            self._slots = []

        def connect(self, fn):
            self._slots.append(fn)

        def disconnect(self, fn=None):
            self._slots = [s for s in self._slots if fn is not None and s is not fn]

        def emit(self, *args):
            for s in list(self._slots):
                s(*args)

    class pyqtSignal:                       # type: ignore[no-redef]
        """Descriptor giving every instance its own connect/emit object, like a Qt signal."""

        def __init__(self, *types):
            self._name = None

        def __set_name__(self, owner, name):
            self._name = "_sig_" + name

        def __get__(self, obj, owner=None):
            if obj is None:
                return self
            s = obj.__dict__.get(self._name)
            if s is None:
                s = obj.__dict__[self._name] = _BoundSignal()
            return s


class VMAFAnalyzer(QObject):
    """VMAF analyzer with the reference's signals (app/vmaf_analyzer.py:20-23)."""
    analysis_progress = pyqtSignal(int)      # 0-100 %
    analysis_complete = pyqtSignal(dict)     # results dict
    error_occurred = pyqtSignal(str)
    status_update = pyqtSignal(str)

    def __init__(self):
        super().__init__()
        self.output_directory = None
        self.test_name = None
        self._process_lock = threading.Lock()
        self._cancel = None                  # threading.Event of the running analysis
        self._terminate_requested = False
        self.threads = 4                     # libvmaf n_threads: kept for API compatibility, unused on GPU
        self.pool_method = "mean"            # mean | min | harmonic_mean
        self.enable_motion_score = False
        self.enable_temporal_features = False
        self.feature_subsample = 1
        self.psnr_enabled = True
        self.ssim_enabled = True
        self.devices = None                  # None -> all visible GPUs
        # CUDA contexts (one per GPU and picture geometry) stay alive between analyses of this analyzer: the reference
        # pays an ffmpeg process start per call (:446), a second call here starts on warm contexts
        self._engine = engine.Engine()
        self.last_ingest = None              # "mapped" (mmap + cudaHostRegister) or "ring" (reader threads), for reports

    # ---- option setters (reference :44-137) --------------------------------------------------
    def set_options_from_manager(self, options_manager):
        if not options_manager:
            logger.warning("No options manager provided, using default settings")
            return
        try:
            s = options_manager.get_setting("vmaf")
            self.threads = s.get("threads", 4)
            self.feature_subsample = s.get("feature_subsample", 1)
            self.pool_method = s.get("pool_method", "mean")
            self.enable_motion_score = s.get("enable_motion_score", False)
            self.enable_temporal_features = s.get("enable_temporal_features", False)
            self.psnr_enabled = s.get("psnr_enabled", True)
            self.ssim_enabled = s.get("ssim_enabled", True)
        except Exception as e:               # noqa: BLE001  (reference :68-69 logs and carries on)
            logger.error(f"Error setting VMAF options from manager: {e}")

    set_options_manager = set_options_from_manager          # the reference has both names (:44, :77)

    def set_output_directory(self, output_dir):
        self.output_directory = output_dir

    def set_test_name(self, test_name):
        self.test_name = test_name

    def set_advanced_options(self, pool_method="mean", enable_motion_score=False, enable_temporal_features=False,
                             feature_subsample=1, psnr_enabled=True, ssim_enabled=True):
        self.pool_method = pool_method
        self.enable_motion_score = enable_motion_score
        self.enable_temporal_features = enable_temporal_features
        self.feature_subsample = feature_subsample
        self.psnr_enabled = psnr_enabled
        self.ssim_enabled = ssim_enabled

    def set_devices(self, devices):
        """GPU ordinals to shard frames over (extension; the reference has `threads` instead)."""
        self.devices = tuple(devices) if devices is not None else None

    def close(self):
        """Free the CUDA contexts kept between analyses (also happens when the analyzer is collected)."""
        self._engine.close()

    def terminate_analysis(self):
        """Cancel the running analysis (reference :139-151 kills the ffmpeg child)."""
        self._terminate_requested = True
        ev = self._cancel
        if ev is not None:
            ev.set()

    # ---- metadata (reference :162-240 shells out to ffprobe) -----------------------------------
    def get_video_metadata(self, video_path, ffprobe_exe=None):
        try:
            info = yuvio.probe(video_path)
            fps = info.fps
            return {"path": video_path, "duration": (info.nb_frames / fps) if fps else 0.0, "frame_rate": fps,
                    "width": info.width, "height": info.height, "pix_fmt": info.pix_fmt, "codec_name": info.codec_name,
                    "bit_rate": int(info.frame_bytes * 8 * fps), "nb_frames": info.nb_frames}
        except Exception as e:               # noqa: BLE001  (reference :233-240 returns None on any failure)
            logger.error(f"Error getting video metadata for {video_path}: {e}")
            return None

    # ---- the hot path -------------------------------------------------------------------------
    def analyze_videos(self, reference_path, distorted_path, model="vmaf_v0.6.1", duration=None):
        """Score a reference/distorted pair; returns the results dict or None (errors are emitted)."""
        with self._process_lock:
            try:
                return self._analyze(reference_path, distorted_path, model)
            except Exception as e:           # noqa: BLE001  (reference :603-609)
                msg = f"Error in VMAF analysis: {e}"
                logger.exception(msg)
                self.error_occurred.emit(msg)
                return None
            finally:
                self._cancel = None

    def _analyze(self, reference_path, distorted_path, model):
        self._terminate_requested = False
        self._cancel = threading.Event()
        if model is None:
            model = "vmaf_v0.6.1"
        self.status_update.emit(f"Analyzing videos with model: {model}")
        for label, p in (("Reference", reference_path), ("Distorted", distorted_path)):
            if not os.path.exists(p):
                msg = f"{label} video not found: {p}"
                logger.error(msg)
                self.error_occurred.emit(msg)
                return None

        output_dir = self.output_directory or os.path.dirname(reference_path)
        timestamp = datetime.now().strftime("%Y%m%d_%H%M%S")
        test_name = self.test_name or "Test"
        parent_dir = os.path.dirname(reference_path)
        if test_name and test_name in parent_dir:
            test_dir = parent_dir
        else:
            test_dir = os.path.join(output_dir, f"{test_name}_{timestamp}")
            os.makedirs(test_dir, exist_ok=True)
        json_path = os.path.join(test_dir, f"{test_name}_{timestamp}_vmaf.json")
        psnr_path = os.path.join(test_dir, f"{test_name}_{timestamp}_psnr.txt")
        ssim_path = os.path.join(test_dir, f"{test_name}_{timestamp}_ssim.txt")

        ref_info, dis_info = yuvio.probe(reference_path), yuvio.probe(distorted_path)
        if (ref_info.width, ref_info.height, ref_info.bpc, ref_info.chroma) != \
                (dis_info.width, dis_info.height, dis_info.bpc, dis_info.chroma):
            msg = (f"Reference and distorted videos differ in format: {ref_info.width}x{ref_info.height} "
                   f"{ref_info.pix_fmt} vs {dis_info.width}x{dis_info.height} {dis_info.pix_fmt}")
            self.error_occurred.emit(msg)
            return None
        vm = M.resolve_model(model)

        devices = self.devices
        if devices is None:
            n = L.load().bv_device_count()
            if n < 1:
                raise RuntimeError("no CUDA device: the B200 VMAF engine has no CPU fallback")
            devices = tuple(range(n))
        # libvmaf options the reference builds (:373-386): n_subsample always; pool != mean adds psnr=1, ssim=1
        extra = self.pool_method != "mean"
        opt = engine.EngineOptions(n_subsample=max(1, int(self.feature_subsample)), psnr=extra, ssim=extra,
                                   ffmpeg_psnr=bool(self.psnr_enabled) and ref_info.chroma != 400,
                                   ffmpeg_ssim=bool(self.ssim_enabled) and ref_info.chroma != 400, devices=devices,
                                   # :388-402: every `feature=` item the reference appends ends in libvmaf's float
                                   # `motion` extractor (the vif_scaleN / adm2 items name no extractor of their own
                                   # and the last `feature=` wins), so both flags add `motion` / `motion2` to the log
                                   float_motion=bool(self.enable_motion_score or self.enable_temporal_features))
        if ref_info.decoder != "raw" or dis_info.decoder != "raw":
            # container decode is sequential (seeking is not frame-exact): one shard, one decoder per file
            devices = tuple(devices)[:1]
        src = engine.FileSource(ref_info, dis_info)
        last = [-1]

        def on_progress(done, total):
            pct = min(95, int(100 * done / max(total, 1)))      # reference :483: capped at 95 while running
            if pct != last[0]:
                last[0] = pct
                self.analysis_progress.emit(pct)

        self._engine.retain(ref_info.width, ref_info.height, ref_info.bpc)
        # a GPU pays for its context and pipeline fill only with a few launch groups of its own: at least 128 frames each
        devices = tuple(devices)[:max(1, min(len(devices), src.nb_frames // 128))]
        opt.devices = devices
        self.status_update.emit(f"Scoring {src.nb_frames} frames on {len(devices)} GPU(s)")
        try:
            self.last_ingest = "mapped" if src.zero_copy else "ring"
            res = self._engine.analyze(src, vm, opt, progress_cb=on_progress, cancel=self._cancel)
        finally:
            src.release()
        if res is None or self._terminate_requested:
            self.status_update.emit("VMAF analysis terminated by user")      # reference :514-518
            return None

        report.write_libvmaf_json(json_path, res["frames"], res["pooled_metrics"], res["fps"])
        psnr_log = None
        if opt.ffmpeg_psnr:
            rows = []
            shapes = ref_info.plane_shapes()
            for r in res["rows"]:
                if r is None or not (r["valid"] & L.FEAT_PSNR_UV):
                    continue
                sse = r["raw"][L.RAW_SSE:L.RAW_SSE + 3]
                rows.append({"mse": [sse[k] / float(shapes[k][0] * shapes[k][1]) for k in range(3)],
                             "areas": [a * b for a, b in shapes]})
            if rows:
                report.write_ffmpeg_psnr_stats(psnr_path, rows, ref_info.bpc)
                psnr_log = psnr_path
        ssim_log = None
        if opt.ffmpeg_ssim:                   # the reference's third ffmpeg pass (:1057-1064): FFmpeg `ssim` stats file
            shapes = ref_info.plane_shapes()
            rows = [{"ssim": list(r["ffssim"]), "weights": [a * b for a, b in shapes]}
                    for r in res["rows"] if r is not None and (r["valid"] & L.FEAT_FFSSIM)]
            if rows:
                report.write_ffmpeg_ssim_stats(ssim_path, rows)
                ssim_log = ssim_path

        pooled = res["pooled_metrics"]
        vmaf_score = pooled["vmaf"]["mean"] if "vmaf" in pooled else None     # reference :652-653 reads the mean
        import json
        with open(json_path) as f:
            raw_results = json.load(f)
        results = {
            "vmaf_score": vmaf_score,
            "psnr_score": os.path.basename(psnr_log) if psnr_log else "Not Available",     # reference :830-831
            "ssim_score": os.path.basename(ssim_log) if ssim_log else "Not Available",
            "json_path": json_path,
            "psnr_log": psnr_log,
            "ssim_log": ssim_log,
            "reference_video": os.path.basename(reference_path),
            "distorted_video": os.path.basename(distorted_path),
            "raw_results": raw_results,
            "model": vm.name,
            "width": ref_info.width,
            "height": ref_info.height,
            # extras (not in the reference dict)
            "pooled_score": pooled["vmaf"].get(self.pool_method, vmaf_score) if "vmaf" in pooled else None,
            "frames_per_second": res["fps"],
            "n_gpus": len(devices),
        }
        self.analysis_progress.emit(100)                                                   # reference :935
        self.status_update.emit(f"VMAF score: {vmaf_score:.6f}" if vmaf_score is not None else "VMAF done")
        self.analysis_complete.emit(results)                                               # reference :963
        return results


class VMAFAnalysisThread(threading.Thread):
    """Caller-side wrapper (reference app/ui/tabs/analysis_tab.py:585-640: a QThread that owns an analyzer
    and forwards its four signals)."""

    def __init__(self, reference_path, distorted_path, model="vmaf_v0.6.1", duration=None):
        super().__init__(daemon=True)
        self.reference_path, self.distorted_path, self.model, self.duration = reference_path, distorted_path, model, duration
        self.analyzer = VMAFAnalyzer()
        self.analysis_progress = self.analyzer.analysis_progress
        self.analysis_complete = self.analyzer.analysis_complete
        self.error_occurred = self.analyzer.error_occurred
        self.status_update = self.analyzer.status_update
        self.results = None

    def set_output_directory(self, d):
        self.analyzer.set_output_directory(d)

    def set_test_name(self, n):
        self.analyzer.set_test_name(n)

    def run(self):
        self.results = self.analyzer.analyze_videos(self.reference_path, self.distorted_path, self.model, self.duration)
