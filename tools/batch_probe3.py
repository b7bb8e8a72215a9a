"""Is the configs[4] batch slow by itself, or because of what ran before it in bench.py?"""
import gc, sys, time
sys.path.insert(0, ".")
import bench
from pqa2_b200 import engine

class A:
    frames_per_step = 0; batch_clips = 16; sharded_frames = 512; fast_float = False
cx = bench.Ctx(A())
pool = bench.Pool(1920, 1080, 8, 64, 100, True, 0)
r = bench.measure_batch(cx, pool); print("batch alone:", round(r["value"]), "fps", flush=True)
orig = engine.Engine.analyze
def timed(self, *a, **k):
    t0 = time.perf_counter(); out = orig(self, *a, **k); timed.t.append(time.perf_counter() - t0); return out
timed.t = []
engine.Engine.analyze = timed
r = bench.measure_batch(cx, pool); print("batch again:", round(r["value"]), "fps; per clip ms", [round(1e3 * t, 1) for t in timed.t], flush=True)
h = bench.measure_workload(cx, "1080p-float", bench.WORKLOADS["1080p-float"], pool, 5, 3, True, True)
print("headline", round(h["value"]), round(h["e2e"]["value"]), flush=True)
timed.t = []
r = bench.measure_batch(cx, pool); print("batch after headline:", round(r["value"]), "fps; per clip ms", [round(1e3 * t, 1) for t in timed.t], flush=True)
i = bench.measure_workload(cx, "1080p-int", bench.WORKLOADS["1080p-int"], pool, 5, 3, False, True)
print("int", round(i["value"]), round(i["e2e"]["value"]), flush=True)
timed.t = []
r = bench.measure_batch(cx, pool); print("batch after int:", round(r["value"]), "fps; per clip ms", [round(1e3 * t, 1) for t in timed.t], flush=True)
gc.collect(); gc.freeze()
timed.t = []
r = bench.measure_batch(cx, pool); print("batch after gc.freeze:", round(r["value"]), "fps; per clip ms", [round(1e3 * t, 1) for t in timed.t], flush=True)
