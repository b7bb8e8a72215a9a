"""Pooling and the on-disk artefacts the reference's consumers read.

* libvmaf JSON log (``output.c`` format) -- read back by the reference at
  ``app/vmaf_analyzer.py:640-690``, ``app/report_generator.py:296-311``,
  ``app/ui/tabs/results_tab.py:3000-3028``.
* FFmpeg ``psnr`` / ``ssim`` stats files (one line per frame) -- the files the
  reference's second and third ffmpeg passes write (``app/vmaf_analyzer.py:1027-1034``, ``:1057-1064``).
* per-frame CSV as the Results tab exports it (``app/ui/tabs/results_tab.py:3006-3028``)."""
from __future__ import annotations

import csv
import math

# What the log's "version" field says (libvmaf writes its own version there; the reference shows it as the model /
# engine string, app/vmaf_analyzer.py:834-838).  It names this engine and states what its numbers were checked against.
VERSION = "b200vmaf-0.2 (restates libvmaf 3.0.0; parity checked against the in-repo CPU oracle only)"


def pool(values) -> dict:
    """libvmaf feature_collector pooling: min / max / mean / harmonic_mean = n / sum(1/(v+1)) - 1."""
    vals = [float(v) for v in values]
    n = len(vals)
    if n == 0:
        return {"min": 0.0, "max": 0.0, "mean": 0.0, "harmonic_mean": 0.0}
    s = 0.0
    hs = 0.0
    for v in vals:
        s += v
        hs += 1.0 / (v + 1.0)
    return {"min": min(vals), "max": max(vals), "mean": s / n, "harmonic_mean": n / hs - 1.0}


def pooled_metrics(frames: list) -> dict:
    names = []
    for fr in frames:
        for k in fr["metrics"]:
            if k not in names:
                names.append(k)
    return {k: pool([fr["metrics"][k] for fr in frames if k in fr["metrics"]]) for k in names}


def _f6(v: float) -> str:
    if isinstance(v, float) and (math.isinf(v) or math.isnan(v)):
        return "null"
    return "%.6f" % v


def write_libvmaf_json(path: str, frames: list, pooled: dict, fps: float, version: str = VERSION,
                       extra: dict | None = None) -> None:
    """Same layout as libvmaf's vmaf_write_output_json (2-space indent, %.6f values)."""
    with open(path, "w") as f:
        f.write("{\n")
        f.write('  "version": "%s",\n' % version)
        f.write('  "fps": %.2f,\n' % fps)
        if extra:
            import json
            for k, v in extra.items():
                f.write('  %s: %s,\n' % (json.dumps(k), json.dumps(v)))
        f.write('  "frames": [')
        for i, fr in enumerate(frames):
            f.write("\n    {\n")
            f.write('      "frameNum": %d,\n' % fr["frameNum"])
            f.write('      "metrics": {\n')
            items = list(fr["metrics"].items())
            for j, (k, v) in enumerate(items):
                f.write('        "%s": %s%s\n' % (k, _f6(v), "," if j + 1 < len(items) else ""))
            f.write("      }\n")
            f.write("    }%s" % ("," if i + 1 < len(frames) else ""))
        f.write("\n  ],\n")
        f.write('  "pooled_metrics": {')
        items = list(pooled.items())
        for i, (k, p) in enumerate(items):
            f.write('\n    "%s": {\n' % k)
            f.write('      "min": %s,\n' % _f6(p["min"]))
            f.write('      "max": %s,\n' % _f6(p["max"]))
            f.write('      "mean": %s,\n' % _f6(p["mean"]))
            f.write('      "harmonic_mean": %s\n' % _f6(p["harmonic_mean"]))
            f.write("    }%s" % ("," if i + 1 < len(items) else ""))
        f.write("\n  },\n")
        f.write('  "aggregate_metrics": {\n  }\n')
        f.write("}\n")


def _db(mse: float, peak: float) -> float:
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)


def write_ffmpeg_psnr_stats(path: str, rows: list, bpc: int) -> dict:
    """rows: per frame {'mse': [y, u, v], 'areas': [..]}.  FFmpeg vf_psnr.c stats_file format."""
    peak = float((1 << bpc) - 1)
    tot = 0.0
    with open(path, "w") as f:
        for n, r in enumerate(rows, 1):
            mse = r["mse"]
            areas = r["areas"]
            avg = sum(m * a for m, a in zip(mse, areas)) / sum(areas)
            tot += avg
            names = ("y", "u", "v")[: len(mse)]
            s = "n:%d mse_avg:%.2f " % (n, avg)
            s += " ".join("mse_%s:%.2f" % (c, m) for c, m in zip(names, mse))
            s += " psnr_avg:%.2f " % _db(avg, peak)
            s += " ".join("psnr_%s:%.2f" % (c, _db(m, peak)) for c, m in zip(names, mse))
            f.write(s + " \n")
    mean_mse = tot / max(len(rows), 1)
    return {"average": _db(mean_mse, peak), "mse_avg": mean_mse}


def write_ffmpeg_ssim_stats(path: str, rows: list) -> dict:
    """rows: per frame {'ssim': [y, u, v], 'weights': [..]}.  FFmpeg vf_ssim.c stats_file format."""
    tot = 0.0
    with open(path, "w") as f:
        for n, r in enumerate(rows, 1):
            ss, wt = r["ssim"], r["weights"]
            allv = sum(s * w for s, w in zip(ss, wt)) / sum(wt)
            tot += allv
            names = ("Y", "U", "V")[: len(ss)]
            s = "n:%d " % n + " ".join("%s:%f" % (c, v) for c, v in zip(names, ss))
            db = float("inf") if allv >= 1.0 else -10.0 * math.log10(1.0 - allv)
            f.write("%s All:%f (%f)\n" % (s, allv, db))
    return {"average": tot / max(len(rows), 1)}


def write_frames_csv(path: str, frames: list) -> None:
    """Per-frame CSV as ResultsTab.export_csv_data writes it: 'Frame Number' + sorted metric keys, %.4f."""
    keys = sorted({k for fr in frames for k in fr["metrics"]})
    with open(path, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["Frame Number"] + keys)
        for fr in frames:
            wr.writerow([fr["frameNum"]] + ["%.4f" % fr["metrics"].get(k, 0.0) for k in keys])


def _mean_of(pooled: dict, *names):
    for n in names:
        if n in pooled and "mean" in pooled[n]:
            return pooled[n]["mean"]
    return None


def _fmt4(v) -> str:
    if v is None:
        return "N/A"
    return v if isinstance(v, str) else "%.4f" % v


def write_result_csv(path: str, test_name: str, data: dict, reference_path: str = "Unknown",
                     distorted_path: str = "Unknown", date: str | None = None) -> str:
    """One test as the Results tab exports it (``ResultsTab.export_result_csv``,
    ``app/ui/tabs/results_tab.py:3518-3616``): summary row, file rows, then the per-frame table."""
    import datetime as _dt
    pooled = data.get("pooled_metrics", {})
    with open(path, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["Test Name", "Date", "VMAF Score", "PSNR Score", "SSIM Score"])
        wr.writerow([test_name, date or _dt.datetime.now().strftime("%Y-%m-%d %H:%M:%S"),
                     _fmt4(_mean_of(pooled, "vmaf")), _fmt4(_mean_of(pooled, "psnr", "psnr_y")),
                     _fmt4(_mean_of(pooled, "ssim", "ssim_y"))])
        wr.writerow([])
        wr.writerow(["Reference File", reference_path])
        wr.writerow(["Distorted File", distorted_path])
        frames = data.get("frames") or []
        if frames:
            wr.writerow([])
            keys = sorted(frames[0].get("metrics", {}).keys())
            wr.writerow(["Frame Number"] + keys)
            for fr in frames:
                m = fr.get("metrics", {})
                wr.writerow([fr.get("frameNum", "N/A")] +
                            [("%.4f" % m[k]) if isinstance(m.get(k), (int, float)) else "N/A" for k in keys])
    return path


def write_combined_csv(path: str, rows: list) -> str:
    """Many tests in one table (``ResultsTab.export_combined_csv``, ``results_tab.py:3644-3696``).  ``rows``: dicts
    with test_name, timestamp, vmaf_score, psnr_score, ssim_score, reference, duration, test_dir."""
    with open(path, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["Test Name", "Date/Time", "VMAF Score", "PSNR Score", "SSIM Score", "Reference", "Duration",
                     "Test Directory"])
        for r in rows:
            wr.writerow([r.get("test_name", ""), r.get("timestamp", ""), _fmt4(r.get("vmaf_score")),
                         _fmt4(r.get("psnr_score")), _fmt4(r.get("ssim_score")), r.get("reference", ""),
                         r.get("duration", ""), r.get("test_dir", "")])
    return path


def write_metadata_json(path: str, results: dict, video: dict, settings: dict, test_name: str = "test",
                        capture: dict | None = None) -> str:
    """``<test>_<stamp>_metadata.json`` next to the log (``AnalysisTab`` writes it, ``analysis_tab.py:765-811``); the
    History tab indexes tests by these files.  ``results``: the analyzer's result dict; ``video``: width / height /
    fps / frame_count / duration_seconds."""
    import json
    import platform
    w, h = video.get("width"), video.get("height")
    meta = {
        "test_name": test_name,
        "vmaf_score": results.get("vmaf_score"),
        "reference_video": results.get("reference_video"),
        "distorted_video": results.get("distorted_video"),
        "psnr_file": results.get("psnr_score") if results.get("psnr_log") else None,
        "ssim_file": results.get("ssim_score") if results.get("ssim_log") else None,
        "json_result": results.get("json_path") and results["json_path"].replace("\\", "/").split("/")[-1],
        "video_details": {"resolution": f"{w}x{h}" if w and h else "Unknown", "width": w, "height": h,
                          "fps": video.get("fps"), "frame_count": video.get("frame_count"),
                          "duration_seconds": video.get("duration_seconds")},
        "analysis_settings": settings,
        "capture_details": capture or {},
        "system_info": {"engine_version": VERSION, "os": platform.system(), "os_version": platform.version(),
                        "processor": platform.processor()},
    }
    with open(path, "w") as f:
        json.dump(meta, f, indent=4)
    return path
