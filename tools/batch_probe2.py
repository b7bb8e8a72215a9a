"""Why do some clips run at half speed end to end?  Per-clip wall time for clip length / pool / walk variants, and the
submit timeline (time stamp every 32 frames) of one slow and one fast case."""
import sys, time
sys.path.insert(0, ".")
import bench
from pqa2_b200 import engine, model as M

class A: frames_per_step = 0
cx = bench.Ctx(A())
model = M.resolve_model("vmaf_v0.6.1")
opt = engine.EngineOptions(devices=(0,))
pools = {"P32 luma": bench.Pool(1920, 1080, 8, 32, 5, False, 0, resident=False),
         "P64 yuv": bench.Pool(1920, 1080, 8, 64, 6, True, 0, resident=False)}
with engine.Engine() as sess:
    for pname, pool in pools.items():
        sess.analyze(pool.clip(64), model, opt)
        for n, off, stride in ((300, 0, 1), (300, 5, 2), (300, 10, 3), (512, 0, 1), (1024, 0, 1), (2048, 0, 1), (300, 0, 1)):
            ts = []
            t0 = time.perf_counter()
            def cb(done, total):
                if done % 32 == 0:
                    ts.append(time.perf_counter() - t0)
            for rep in range(3):
                ts.clear()
                t0 = time.perf_counter()
                sess.analyze(pool.clip(n, offset=off, stride=stride), model, opt, progress_cb=cb)
                dt = time.perf_counter() - t0
            steps = [round(1e3 * (b - a), 1) for a, b in zip([0.0] + ts[:-1], ts)]
            print(f"{pname} n={n} off={off} stride={stride}: {1e3 * dt:.1f} ms -> {n / dt:.0f} fps; ms per 32 submits: {steps[:40]}", flush=True)
