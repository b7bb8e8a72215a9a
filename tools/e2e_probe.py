"""Where does the end-to-end time go?  Times the phases of one engine-style pass over pinned host frames."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from pqa2_b200 import _lib as L, engine, model as M, synth
from pqa2_b200.extractor import FeatureExtractor, pinned_empty

w, h, n, P = 1920, 1080, 512, 32
wl = sys.argv[1] if len(sys.argv) > 1 else "int"
model = M.resolve_model("vmaf_v0.6.1" if wl == "int" else "vmaf_float_v0.6.1")
opt = engine.EngineOptions(psnr=wl != "int", ssim=wl != "int", ms_ssim=wl != "int")
mask = engine.feature_mask(model, opt)
ref = [pinned_empty((h, w), np.uint8) for _ in range(P)]
dis = [pinned_empty((h, w), np.uint8) for _ in range(P)]
for i in range(P):
    rp, dp = synth.frame_pair(1, i, w, h, 8, chroma=False)
    ref[i][...] = rp[0]; dis[i][...] = dp[0]
fx = FeatureExtractor(w, h, 8, 0, mask)
for rep in range(3):
    fx.reset()
    t0 = time.perf_counter()
    for i in range(n):
        fx.submit(i, [ref[i % P]], [dis[i % P]], L.FRAME_FIRST if i == 0 else 0)
    t1 = time.perf_counter()
    fx.flush()
    t2 = time.perf_counter()
    out = np.ctypeslib.as_array(fx.fetch(0, n))
    rows = engine.Rows(n); rows.put(out["frame_index"], out)
    t3 = time.perf_counter()
    frames = engine.build_frames(rows, model, opt, 0)
    t4 = time.perf_counter()
    from pqa2_b200 import report
    report.pooled_metrics(frames)
    t5 = time.perf_counter()
    print(f"{wl} rep{rep}: submit {1e3*(t1-t0):.1f} ms ({1e6*(t1-t0)/n:.1f} us/frame)  flush {1e3*(t2-t1):.1f}  fetch {1e3*(t3-t2):.1f}  "
          f"build+svr {1e3*(t4-t3):.1f}  pool {1e3*(t5-t4):.1f}  total {1e3*(t5-t0):.1f} ms -> {n/(t5-t0):.0f} fps")
