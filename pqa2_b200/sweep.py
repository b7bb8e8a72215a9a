"""A sweep of clip pairs over the box's GPUs with the per-test artefacts the reference's History tab indexes
(BASELINE.json configs[4]; SURVEY.md §8 rows e and f3).

The reference runs one ffmpeg child per pair, one after another, and the GUI then writes, per test directory,
``<test>_<stamp>_vmaf.json`` (``app/vmaf_analyzer.py:304``), ``<test>_<stamp>_metadata.json``
(``app/ui/tabs/analysis_tab.py:765-811``) and on request the per-test / combined CSV exports
(``app/ui/tabs/results_tab.py:3518-3696``).  Here whole clips are dealt round-robin to the devices
(``engine.analyze_batch``: one session per GPU, no lead-in frames, no collective) and the same files are written
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


def read_pairs(path: str) -> list:
    """Pairs file: CSV rows ``name,reference,distorted`` (a header row with those words is skipped); relative paths are
    resolved against the file's directory."""
    import csv
    base = os.path.dirname(os.path.abspath(path))
    out = []
    with open(path, newline="") as f:
        for row in csv.reader(f):
            row = [c.strip() for c in row]
            if len(row) < 3 or not row[0] or row[0].startswith("#") or row[1].lower() in ("reference", "ref"):
                continue
            out.append((row[0], os.path.join(base, row[1]), os.path.join(base, row[2])))
    return out


def main(argv=None) -> int:
    """python -m pqa2_b200.sweep PAIRS.csv [--out DIR] [--model vmaf_v0.6.1] [--gpus 0,1,...] [--align-bookends]"""
    import argparse
    import sys
    from . import _lib as L
    from . import model as M
    from . import yuvio
    ap = argparse.ArgumentParser(prog="python -m pqa2_b200.sweep", description=main.__doc__)
    ap.add_argument("pairs")
    ap.add_argument("--out", default="sweep_results")
    ap.add_argument("--model", default="vmaf_v0.6.1")
    ap.add_argument("--gpus", default=None)
    ap.add_argument("--align-bookends", action="store_true", help="every distorted clip is a capture with white bookends")
    a = ap.parse_args(argv)
    pairs = read_pairs(a.pairs)
    if not pairs:
        print("error: no pairs in", a.pairs, file=sys.stderr)
        return 1
    n = L.load().bv_device_count()
    if n < 1:
        print("error: no CUDA device (the B200 VMAF engine has no CPU fallback)", file=sys.stderr)
        return 1
    devices = [int(x) for x in a.gpus.split(",")] if a.gpus else list(range(n))
    clips, names = [], []
    for name, ref, dis in pairs:
        if a.align_bookends:
            from . import alignment
            res = alignment.align_by_bookends(ref, dis, device=devices[0])
            clips.append(res["source"])
        else:
            clips.append(engine.FileSource(yuvio.probe(ref), yuvio.probe(dis)))
        names.append(name)
    out = run_sweep(clips, M.resolve_model(a.model), a.out, names, devices=devices)
    for r in out["results"]:
        print(r["test_name"], "error: " + r["error"] if "error" in r else "%.6f" % r["vmaf_score"])
    print("combined:", out["combined_csv"])
    return 0 if all("error" not in r for r in out["results"]) else 2


if __name__ == "__main__":
    import sys
    sys.exit(main())
