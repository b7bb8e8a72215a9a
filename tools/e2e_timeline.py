"""Where one host-buffer engine.analyze() call spends its wall time: submit loop, drain (flush), feature read-back, frame
building + SVR + pooling.  Usage: python tools/e2e_timeline.py [1080p-float|1080p-int|4k-int] [contexts on GPU 0] [frames] [chunk]"""
import sys, time
sys.path.insert(0, '.')
import bench
from pqa2_b200 import engine, model as M, extractor

wname = sys.argv[1] if len(sys.argv) > 1 else "1080p-float"
wl = bench.WORKLOADS[wname]
pool = bench.Pool(wl["w"], wl["h"], wl["bpc"], wl["pool"], 100, False, 0, resident=False)
model = M.resolve_model(wl["model"])
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # contexts on GPU 0 (frame shards side by side); 0 = the engine's choice
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 512
opt = engine.EngineOptions(devices=(0,), contexts_per_device=ndev, dynamic_chunk=chunk, psnr=wl["psnr"], ssim=wl["ssim"],
                           ms_ssim=wl["ms_ssim"])
T = {}
def wrap(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        t = time.perf_counter(); r = f(self, *a, **k); T[name] = T.get(name, 0.0) + time.perf_counter() - t; return r
    setattr(cls, name, g)
for nm in ("submit", "flush", "fetch", "kick", "wait_uploads", "reset"):
    wrap(extractor.FeatureExtractor, nm)
bf = engine.build_frames
def timed_bf(*a, **k):
    t = time.perf_counter(); r = bf(*a, **k); T["build_frames"] = T.get("build_frames", 0.0) + time.perf_counter() - t; return r
engine.build_frames = timed_bf
n = int(sys.argv[3]) if len(sys.argv) > 3 else 512
with engine.Engine() as sess:
    for rep in range(4):
        T.clear()
        t0 = time.perf_counter()
        res = sess.analyze(pool.clip(n), model, opt)
        dt = time.perf_counter() - t0
        rest = dt - sum(T.values())
        print(f"{wname} x{ndev} chunk {chunk} n {n} call {rep}: {1e3 * dt:7.2f} ms = {n / dt:7.1f} fps | " +
              " ".join(f"{k} {1e3 * v:.2f}" for k, v in T.items()) + f" | other {1e3 * rest:.2f}")
