"""Writes tests/golden/*.json: regression vectors = the CPU oracle's outputs on seeded synthetic frames.

These are NOT outputs of the reference (ffmpeg + libvmaf cannot run in this image and the reference ships no
golden vectors); they freeze the oracle so that an accidental change to oracle/ or to the
synthetic generator shows up, and give the GPU tests a second, file-based comparison point.

    python tests/tools/make_golden.py

Lives under tests/ because it imports the oracle (only tests/, smoke() and bench.py's CPU legs may)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pqa2_b200 import synth  # noqa: E402

CASES = [dict(seed=3, w=176, h=144, bpc=8, n=3), dict(seed=8, w=208, h=120, bpc=10, n=2),
         dict(seed=5, w=161, h=97, bpc=8, n=2), dict(seed=9, w=242, h=178, bpc=12, n=2)]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for c in CASES:
        w, h, bpc = c["w"], c["h"], c["bpc"]
        frames = [synth.frame_pair(c["seed"], f, w, h, bpc) for f in range(c["n"])]
        rows, prev_blur, prev_ref = [], None, None
        for rp, dp in frames:
            blur = oracle.motion_blur(rp[0], bpc)
            sad = 0 if prev_blur is None else oracle.motion_sad(blur, prev_blur)
            prev_blur = blur
            v, a = oracle.vif(rp[0], dp[0], bpc), oracle.adm(rp[0], dp[0], bpc)
            fl = oracle.float_features(rp[0], dp[0], bpc, prev_ref=prev_ref, psnr=True, ssim=True,
                                       ms_ssim=min(w, h) >= 176)
            prev_ref = rp[0]
            rows.append({
                "luma_crc": [int(np.bitwise_xor.reduce(rp[0].astype(np.uint32).ravel() * np.uint32(2654435761))),
                             int(np.bitwise_xor.reduce(dp[0].astype(np.uint32).ravel() * np.uint32(2654435761)))],
                "sad": int(sad), "vif_acc": v["acc"].tolist(), "adm_cm": a["cm"].tolist(),
                "adm_den": [[int(x) for x in r] for r in a["den"]], "adm2": a["adm2"],
                "vif_score": [float(x) for x in v["score"]],
                "sse": [int(oracle.sse(rp[k], dp[k], bpc)) for k in range(3)],
                "ffssim": [oracle.ffssim_plane(rp[k], dp[k], bpc) for k in range(3)],
                "float": {k: fl[k] for k in fl if isinstance(fl[k], float)},
            })
        path = os.path.join(out_dir, "oracle_%dx%d_%dbit_seed%d.json" % (w, h, bpc, c["seed"]))
        with open(path, "w") as f:
            json.dump({"case": c, "source": "oracle/ (CPU restatement; not reference output)", "frames": rows}, f, indent=1)
        print(path)


def fullsize():
    """One frame pair at each BASELINE picture size: raw integer accumulators only (the float rows would take minutes)."""
    cases = []
    for seed, w, h, bpc, egl in ((11, 1920, 1080, 8, 100.0), (12, 3840, 2160, 10, 100.0), (12, 3840, 2160, 10, 1.0)):
        rp, dp = synth.frame_pair(seed, 0, w, h, bpc, chroma=False)
        v, a = oracle.vif(rp[0], dp[0], bpc, egl), oracle.adm(rp[0], dp[0], bpc, egl)
        cases.append({"seed": seed, "w": w, "h": h, "bpc": bpc, "egl": egl, "vif_acc": v["acc"].tolist(),
                      "adm_cm": a["cm"].tolist(), "adm_den": [[int(x) for x in r] for r in a["den"]], "adm2": a["adm2"],
                      "sse_y": int(oracle.sse(rp[0], dp[0], bpc))})
    path = os.path.join(ROOT, "tests", "golden", "fullsize_oracle.json")
    with open(path, "w") as f:
        json.dump({"source": "oracle/ (CPU restatement; not reference output)", "cases": cases}, f, indent=1)
    print(path)


if __name__ == "__main__":
    main()
    fullsize()
