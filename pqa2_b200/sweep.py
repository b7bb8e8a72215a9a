"""A sweep of clip pairs over the box's GPUs with the per-test artefacts the reference's History tab indexes
(BASELINE.json configs[4]; SURVEY.md §8 rows e and f3).

The reference runs one ffmpeg child per pair, one after another, and the GUI then writes, per test directory,
``<test>_<stamp>_vmaf.json`` (``app/vmaf_analyzer.py:304``), ``<test>_<stamp>_metadata.json``
(``app/ui/tabs/analysis_tab.py:765-811``) and on request the per-test / combined CSV exports
(``app/ui/tabs/results_tab.py:3518-3696``).  Here whole clips are dealt round-robin to the devices
(``engine.analyze_batch``: three clips in flight per GPU, a session each, no lead-in frames, no collective) and the same files are written
from the returned logs."""
from __future__ import annotations

import os
import time

from . import engine, report
from .model import VmafModel


def run_sweep(clips: list, model: VmafModel, out_dir: str, names: list | None = None, opt=None, devices=None,
              stamp: str | None = None, progress_cb=None) -> dict:
    """``clips``: FrameSource objects (``engine.FileSource`` pairs, ``alignment.AlignedSource``, ...).

    Returns ``{"results": [per-clip summary dicts], "combined_csv": path}``; a failed clip carries ``error``."""
    opt = opt or engine.EngineOptions(psnr=True, ssim=True)
    stamp = stamp or time.strftime("%Y%m%d_%H%M%S")
    names = names or [f"clip{k:03d}" for k in range(len(clips))]
    if len(names) != len(clips):
        raise ValueError("one name per clip")
    os.makedirs(out_dir, exist_ok=True)
    logs = engine.analyze_batch(clips, model, opt, devices, progress_cb)
    rows, results = [], []
    for name, src, log in zip(names, clips, logs):
        test_dir = os.path.join(out_dir, f"{name}_{stamp}")
        os.makedirs(test_dir, exist_ok=True)
        if log is None or "error" in log:
            results.append({"test_name": name, "test_dir": test_dir, "error": (log or {}).get("error", "cancelled")})
            continue
        json_path = os.path.join(test_dir, f"{name}_{stamp}_vmaf.json")
        report.write_libvmaf_json(json_path, log["frames"], log["pooled_metrics"], log["fps"])
        pooled = log["pooled_metrics"]
        ref_name = os.path.basename(getattr(getattr(src, "_ri", None), "path", "") or "synthetic")
        dis_name = os.path.basename(getattr(getattr(src, "_di", None) or getattr(src, "_ci", None), "path", "") or "synthetic")
        report.write_result_csv(os.path.join(test_dir, f"{name}_data_{stamp}.csv"), name, log, ref_name, dis_name)
        fps = float(getattr(src, "fps", 0.0) or 0.0)
        dur = log["n_frames"] / fps if fps else None
        summary = {"test_name": name, "timestamp": time.strftime("%Y-%m-%d %H:%M:%S"), "test_dir": test_dir,
                   "json_path": json_path, "vmaf_score": pooled["vmaf"]["mean"],
                   "psnr_score": report._mean_of(pooled, "psnr", "psnr_y"),
                   "ssim_score": report._mean_of(pooled, "ssim", "ssim_y", "float_ssim"),
                   "reference": ref_name, "reference_video": ref_name, "distorted_video": dis_name,
                   "duration": ("%.2fs" % dur) if dur else "", "n_frames": log["n_frames"], "model": log["model"]}
        report.write_metadata_json(os.path.join(test_dir, f"{name}_{stamp}_metadata.json"), summary,
                                   {"width": src.width, "height": src.height, "fps": fps, "frame_count": log["n_frames"],
                                    "duration_seconds": dur},
                                   {"model": log["model"], "n_subsample": opt.n_subsample, "psnr": opt.psnr, "ssim": opt.ssim,
                                    "ms_ssim": opt.ms_ssim}, name)
        rows.append(summary)
        results.append(summary)
    combined = report.write_combined_csv(os.path.join(out_dir, f"combined_results_{stamp}.csv"), rows)
    return {"results": results, "combined_csv": combined}
