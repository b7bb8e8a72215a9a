"""configs[4] on one GPU with 1 .. 4 clips in flight (dist.analyze_batch_distributed `concurrency`)."""
import sys, time
sys.path.insert(0, '.')
import bench
from pqa2_b200 import dist as D, engine, model as M

wl = bench.WORKLOADS["1080p-int"]
pool = bench.Pool(wl["w"], wl["h"], wl["bpc"], wl["pool"], 100, False, 0, resident=False)
model = M.resolve_model("vmaf_v0.6.1")
opt = engine.EngineOptions(devices=(0,))
clips = [pool.clip(300, offset=5 * k, stride=1 + k % 3) for k in range(64)]
for conc in (1, 2, 3, 4):
    sessions = [engine.Engine() for _ in range(conc)]
    D.analyze_batch_distributed([pool.clip(64) for _ in range(2 * conc)], model, opt, device=0, session=sessions, concurrency=conc)
    best = 1e9
    for rep in range(2):
        t0 = time.perf_counter()
        out = D.analyze_batch_distributed(clips, model, opt, device=0, session=sessions, concurrency=conc)
        best = min(best, time.perf_counter() - t0)
    print(f"{conc} clips in flight: {64 * 300 / best:8.1f} fps  sum of pooled means {sum(o['pooled_metrics']['vmaf']['mean'] for o in out):.9f}")
    for s in sessions:
        s.close()
