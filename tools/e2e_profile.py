"""cProfile of one steady-state engine.analyze call (512 pinned 1080p frames): where the host time of the e2e path goes."""
import cProfile, pstats, sys, time
import numpy as np
sys.path.insert(0, '.')
from pqa2_b200 import engine, model as M, synth
from pqa2_b200.extractor import pinned_empty

w, h, n, P = 1920, 1080, 512, 32
wl = sys.argv[1] if len(sys.argv) > 1 else "int"
model = M.resolve_model("vmaf_v0.6.1" if wl == "int" else "vmaf_float_v0.6.1")
opt = engine.EngineOptions(psnr=wl != "int", ssim=wl != "int", ms_ssim=wl != "int")
ref = [pinned_empty((h, w), np.uint8) for _ in range(P)]
dis = [pinned_empty((h, w), np.uint8) for _ in range(P)]
for i in range(P):
    rp, dp = synth.frame_pair(1, i, w, h, 8, chroma=False)
    ref[i][...] = rp[0]; dis[i][...] = dp[0]

class Clip(engine.FrameSource):
    zero_copy = True
    width, height, bpc, chroma, nb_frames, fps = w, h, 8, 0, n, 30.0
    def get(self, i, luma_only):
        return [ref[i % P]], [dis[i % P]]

with engine.Engine() as sess:
    for _ in range(3):
        sess.analyze(Clip(), model, opt)
    t0 = time.perf_counter()
    for _ in range(5):
        sess.analyze(Clip(), model, opt)
    print(f"{wl}: {1e3 * (time.perf_counter() - t0) / 5:.2f} ms per call -> {5 * n / (time.perf_counter() - t0):.0f} fps")
    import threading
    pr = cProfile.Profile()
    # profile the shard thread too: run the shard inline by calling _run_shard through analyze with profiling enabled in threads
    threading.setprofile(lambda *a: None)
    pr.enable()
    sess.analyze(Clip(), model, opt)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
