"""Command line: score a reference/distorted pair of raw clips the way the reference's ffmpeg call does.

    python -m pqa2_b200 REF.y4m DIST.y4m [--model vmaf_v0.6.1] [--out DIR] [--name TEST] [--pool mean]
                        [--subsample N] [--no-psnr] [--no-ssim] [--gpus 0,1,...] [--align-bookends]

Equivalent of `ffmpeg -i DIST -i REF -lavfi libvmaf=log_path=...:log_fmt=json:model=version=<m>:n_subsample=<N>`
plus the two `psnr` / `ssim` passes (reference app/vmaf_analyzer.py:373-419, :996-1092)."""
from __future__ import annotations

import argparse
import sys

from .vmaf_analyzer import VMAFAnalyzer


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m pqa2_b200", description=__doc__.split("\n")[0])
    ap.add_argument("reference")
    ap.add_argument("distorted")
    ap.add_argument("--model", default="vmaf_v0.6.1")
    ap.add_argument("--out", default=None, help="output directory (default: next to the reference clip)")
    ap.add_argument("--name", default="Test")
    ap.add_argument("--pool", default="mean", choices=["mean", "min", "harmonic_mean"])
    ap.add_argument("--subsample", type=int, default=1)
    ap.add_argument("--no-psnr", action="store_true")
    ap.add_argument("--no-ssim", action="store_true")
    ap.add_argument("--gpus", default=None, help="comma-separated GPU ordinals (default: all)")
    ap.add_argument("--align-bookends", action="store_true",
                    help="DIST is a capture with white bookends: find the content window on the GPU and trim/re-time it "
                         "losslessly before scoring (reference app/bookend_alignment.py)")
    a = ap.parse_args(argv)

    an = VMAFAnalyzer()
    if a.out:
        an.set_output_directory(a.out)
    an.set_test_name(a.name)
    an.set_advanced_options(pool_method=a.pool, feature_subsample=a.subsample, psnr_enabled=not a.no_psnr,
                            ssim_enabled=not a.no_ssim)
    if a.gpus:
        an.set_devices([int(x) for x in a.gpus.split(",")])
    an.error_occurred.connect(lambda m: print("error:", m, file=sys.stderr))
    an.status_update.connect(lambda m: print(m, file=sys.stderr))
    ref, dis = a.reference, a.distorted
    if a.align_bookends:
        from . import alignment
        al = alignment.align_bookend_videos(ref, dis, a.out)
        if al is None:
            print("error: Failed to detect white bookends in captured video", file=sys.stderr)
            return 1
        ref, dis = al["aligned_reference"], al["aligned_captured"]
        print("aligned: %d frames, captured frames %d..%d" % (al["plan"].n_frames, al["plan"].cap_frames[0],
                                                               al["plan"].cap_frames[-1]), file=sys.stderr)
    res = an.analyze_videos(ref, dis, a.model)
    if res is None:
        return 1
    print("VMAF score: %.6f" % res["vmaf_score"])
    print("log:", res["json_path"])
    for k in ("psnr_log", "ssim_log"):
        if res.get(k):
            print(k + ":", res[k])
    return 0


if __name__ == "__main__":
    sys.exit(main())
