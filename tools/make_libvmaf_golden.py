"""One-command parity pin: run the REFERENCE's own command on seeded synthetic clips and store libvmaf's JSON logs.

    python tools/make_libvmaf_golden.py [--ffmpeg /path/to/ffmpeg] [--out tests/golden/libvmaf] [--keep-clips DIR]

For every case below the script writes the synthetic ref / dis pair as .y4m (pqa2_b200.synth, the generator every test
uses), runs exactly what app/vmaf_analyzer.py:411-419 runs --

    ffmpeg -hide_banner -loglevel info -i DIS.y4m -i REF.y4m \
        -lavfi libvmaf=log_path=LOG:log_fmt=json:model=version=<model>:n_threads=1:n_subsample=1[:psnr=1:ssim=1] -f null -

(first input = distorted, second = reference; n_threads=1 so that the log is reproducible) -- and stores the log as
tests/golden/libvmaf/<case>.json together with ffmpeg's and libvmaf's version strings.  tests/test_libvmaf_golden.py
then compares the CPU oracle (always) and the CUDA engine (-m gpu) with those logs: integer-model features bit-exact
at libvmaf's six printed decimals, float models within 1e-4 per frame / 1e-5 pooled.

An ffmpeg with the libvmaf filter is looked for in --ffmpeg, $B200VMAF_FFMPEG, baseline/_ref/ (bin/ffmpeg or ffmpeg) and
PATH.  This image has none, so here the script only lists what it would do and exits 3; the
fixtures directory stays empty and the golden test skips with that reason.  Nothing is fabricated."""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name, seed, w, h, bpc, frames, model, extra libvmaf options, synth kwargs
CASES = [
    ("qcif_8bit_v061", 3, 176, 144, 8, 3, "vmaf_v0.6.1", "psnr=1:ssim=1", {}),
    ("wqvga_10bit_v061", 8, 416, 240, 10, 3, "vmaf_v0.6.1", "psnr=1", {}),
    ("odd_333x251_v061", 3, 333, 251, 8, 2, "vmaf_v0.6.1", "", {}),
    ("strong_blur_v061", 5, 480, 270, 8, 2, "vmaf_v0.6.1", "", {"strength": 3, "q": 8}),        # VIF non-log / low-gain branches
    ("neg_480x270", 5, 480, 270, 8, 2, "vmaf_v0.6.1neg", "", {}),                                # both NEG gain limits (sharpened patch)
    ("boot_480x270", 5, 480, 270, 8, 2, "vmaf_b_v0.6.3", "", {}),
    ("hd1080_8bit_v061", 11, 1920, 1080, 8, 2, "vmaf_v0.6.1", "psnr=1:ssim=1", {}),
    ("uhd2160_10bit_4k", 12, 3840, 2160, 10, 2, "vmaf_4k_v0.6.1", "", {}),
    ("uhd2160_10bit_neg", 12, 3840, 2160, 10, 1, "vmaf_v0.6.1neg", "", {}),
    ("float_qcif", 3, 176, 144, 8, 3, "vmaf_float_v0.6.1", "psnr=1:ssim=1:ms_ssim=1", {}),
    ("float_640x360", 21, 640, 360, 8, 6, "vmaf_float_v0.6.1", "ssim=1:ms_ssim=1", {}),
    ("float_hd1080", 11, 1920, 1080, 8, 2, "vmaf_float_v0.6.1", "psnr=1:ssim=1:ms_ssim=1", {}),
]


def find_ffmpeg(explicit: str | None) -> str | None:
    cands = [explicit, os.environ.get("B200VMAF_FFMPEG"), os.path.join(ROOT, "baseline", "_ref", "bin", "ffmpeg"),
             os.path.join(ROOT, "baseline", "_ref", "ffmpeg"), shutil.which("ffmpeg")]
    for c in cands:
        if not c or not os.path.exists(c):
            continue
        try:
            out = subprocess.run([c, "-hide_banner", "-filters"], capture_output=True, text=True, timeout=30).stdout
        except Exception:
            continue
        if " libvmaf " in out:
            return c
    return None


def case_command(ffmpeg: str, dis: str, ref: str, log: str, model: str, extra: str) -> list:
    opts = [f"log_path={log}", "log_fmt=json", f"model=version={model}", "n_threads=1", "n_subsample=1"]
    if extra:
        opts.append(extra)
    return [ffmpeg, "-hide_banner", "-loglevel", "info", "-i", dis, "-i", ref, "-lavfi", "libvmaf=" + ":".join(opts),
            "-f", "null", "-"]


def write_pair(d: str, name: str, seed, w, h, bpc, n, kw) -> tuple:
    from pqa2_b200 import synth, yuvio
    frames = [synth.frame_pair(seed, f, w, h, bpc, **kw) for f in range(n)]
    rp, dp = os.path.join(d, f"{name}_ref.y4m"), os.path.join(d, f"{name}_dis.y4m")
    yuvio.write_y4m(rp, (fr[0] for fr in frames), w, h, bpc)
    yuvio.write_y4m(dp, (fr[1] for fr in frames), w, h, bpc)
    return rp, dp


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ffmpeg")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "libvmaf"))
    ap.add_argument("--keep-clips", help="write the .y4m pairs here and keep them")
    ap.add_argument("--only", help="comma-separated case names")
    a = ap.parse_args()
    ffmpeg = find_ffmpeg(a.ffmpeg)
    only = set(a.only.split(",")) if a.only else None
    cases = [c for c in CASES if not only or c[0] in only]
    if ffmpeg is None:
        print("no ffmpeg with the libvmaf filter found (looked at --ffmpeg, $B200VMAF_FFMPEG, baseline/_ref/, PATH).")
        print("would run, per case:")
        for name, seed, w, h, bpc, n, model, extra, kw in cases:
            print("  " + " ".join(case_command("ffmpeg", f"{name}_dis.y4m", f"{name}_ref.y4m", f"{name}.json", model, extra)))
        return 3
    os.makedirs(a.out, exist_ok=True)
    ver = subprocess.run([ffmpeg, "-hide_banner", "-version"], capture_output=True, text=True).stdout.splitlines()[0]
    work = a.keep_clips or tempfile.mkdtemp(prefix="b200vmaf_golden_")
    os.makedirs(work, exist_ok=True)
    try:
        for name, seed, w, h, bpc, n, model, extra, kw in cases:
            rp, dp = write_pair(work, name, seed, w, h, bpc, n, kw)
            log = os.path.join(work, f"{name}.json")
            r = subprocess.run(case_command(ffmpeg, dp, rp, log, model, extra), capture_output=True, text=True)
            if r.returncode != 0 or not os.path.exists(log):
                print(f"{name}: ffmpeg failed ({r.returncode})\n{r.stderr[-2000:]}")
                return 1
            data = json.load(open(log))
            out = {"case": {"name": name, "seed": seed, "w": w, "h": h, "bpc": bpc, "n": n, "model": model,
                            "libvmaf_options": extra, "synth": kw},
                   "source": "ffmpeg -lavfi libvmaf (reference command, app/vmaf_analyzer.py:411-419), n_threads=1",
                   "ffmpeg": ver, "libvmaf_version": data.get("version"), "log": data}
            with open(os.path.join(a.out, f"{name}.json"), "w") as f:
                json.dump(out, f, indent=1)
            print(f"{name}: vmaf mean {data['pooled_metrics']['vmaf']['mean']:.6f} ({n} frames) -> {a.out}/{name}.json")
    finally:
        if not a.keep_clips:
            shutil.rmtree(work, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
