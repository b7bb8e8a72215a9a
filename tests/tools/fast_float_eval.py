"""bv_opts.fast_float (contracted multiply-add + folded symmetric taps in the float stencils) against the oracle:
max |dVMAF| per frame and on the pooled mean, per-feature deltas, on every float test clip plus 64 frames of 1080p.
Writes a markdown table (default gpurun_out/r02_fast_float.md).  North-star tolerance: 1e-4 per frame, 1e-5 pooled."""
import multiprocessing as mp
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def _oracle_job(a):
    import oracle
    from pqa2_b200 import synth
    seed, f, w, h, bpc = a
    rp, dp = synth.frame_pair(seed, f, w, h, bpc, chroma=False)
    prev = synth.frame_pair(seed, f - 1, w, h, bpc, chroma=False)[0][0] if f > 0 else None
    r = oracle.float_features(rp[0], dp[0], bpc, prev_ref=prev, psnr=True, ssim=True, ms_ssim=min(w, h) >= 176)
    return {k: r.get(k) for k in ("adm2", "motion", "vif_scale0", "vif_scale1", "vif_scale2", "vif_scale3",
                                   "adm_scale0", "adm_scale1", "adm_scale2", "adm_scale3", "float_ssim", "float_ms_ssim")}


def main():
    from pqa2_b200 import _lib as L, engine, model as M, synth
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_fast_float.md"
    model = M.resolve_model("vmaf_float_v0.6.1")
    cases = [(3, 176, 144, 8, 3), (3, 322, 242, 8, 2), (3, 333, 251, 8, 2), (3, 640, 360, 8, 5), (3, 416, 240, 10, 3),
             (3, 960, 540, 8, 1), (4, 352, 288, 8, 2), (4, 1280, 720, 8, 2), (4, 335, 253, 8, 2), (21, 640, 360, 8, 6),
             (100, 1920, 1080, 8, 64)]
    lines = ["# fast_float vs the oracle (tests/tools/fast_float_eval.py)", "",
             "`fast` = bv_opts.fast_float (FMA-contracted, folded-tap stencils); `faithful` = the default build (libvmaf's scalar "
             "operation order).  Deltas are against oracle/ (CPU restatement) through the same SVR (vmaf_float_v0.6.1).", "",
             "clip | frames | max abs dVMAF/frame faithful | fast | pooled-mean dVMAF faithful | fast | max abs d(feature) fast | d float_ssim fast | d float_ms_ssim fast",
             "--- | --- | --- | --- | --- | --- | --- | --- | ---"]
    worst = {"frame": 0.0, "pooled": 0.0}
    with mp.get_context("spawn").Pool(os.cpu_count() or 1) as pool:
        for seed, w, h, bpc, n in cases:
            rows = pool.map(_oracle_job, [(seed, f, w, h, bpc) for f in range(n)])
            motion = [r["motion"] for r in rows]
            motion[0] = 0.0
            m2 = engine.motion2_from_motion(motion)
            feats = np.array([[r["adm2"], m2[i], r["vif_scale0"], r["vif_scale1"], r["vif_scale2"], r["vif_scale3"]]
                              for i, r in enumerate(rows)])
            want = model.main.predict(feats, False, False, device=None)
            ms = min(w, h) >= 176
            res = {}
            for fast in (False, True):
                o = engine.EngineOptions(psnr=True, ssim=True, ms_ssim=ms, fast_float=fast)
                res[fast] = engine.analyze(engine.SynthSource(w, h, bpc, n, seed=seed, chroma=0), model, o)
            d = {}
            for fast in (False, True):
                got = np.array([fr["metrics"]["vmaf"] for fr in res[fast]["frames"]])
                d[fast] = (float(np.max(np.abs(got - want))), float(abs(got.mean() - want.mean())))
            fr = res[True]["frames"]
            dfeat = max(abs(fr[i]["metrics"][k] - rows[i][k]) for i in range(n)
                        for k in ("adm2", "vif_scale0", "vif_scale1", "vif_scale2", "vif_scale3"))
            dss = max(abs(fr[i]["metrics"]["float_ssim"] - rows[i]["float_ssim"]) for i in range(n))
            dms = max(abs(fr[i]["metrics"]["float_ms_ssim"] - rows[i]["float_ms_ssim"]) for i in range(n)) if ms else float("nan")
            worst["frame"] = max(worst["frame"], d[True][0]); worst["pooled"] = max(worst["pooled"], d[True][1])
            lines.append(f"seed {seed} {w}x{h} {bpc}-bit | {n} | {d[False][0]:.2e} | {d[True][0]:.2e} | {d[False][1]:.2e} | "
                         f"{d[True][1]:.2e} | {dfeat:.2e} | {dss:.2e} | {dms:.2e}")
            print(lines[-1], flush=True)
    lines += ["", f"Worst case fast: {worst['frame']:.2e} per frame (tolerance 1e-4), {worst['pooled']:.2e} pooled (tolerance 1e-5)."]
    # throughput of the two builds, resident 1080p, the bench's headline features
    import bench
    class A: frames_per_step = 0
    cx = bench.Ctx(A())
    cx.args.fast_float = False
    wl = bench.WORKLOADS["1080p-float"]
    pool_ = bench.Pool(1920, 1080, 8, 64, 100, True, 0)
    for fast in (False, True):
        cx.args.fast_float = fast
        r = bench.measure_workload(cx, "1080p-float", wl, pool_, 10, 3, False, True)
        lines.append(f"{'fast' if fast else 'faithful'}: {r['value']:.0f} fps resident, {r['e2e']['value']:.0f} fps e2e (1080p, float model + psnr + ssim + ms-ssim); "
                     + ", ".join(f"{k} {v['ms_per_launch']:.3f}" for k, v in r["kernels"].items() if v["ms_per_launch"] > 0.15))
        print(lines[-1], flush=True)
    open(out_path, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
