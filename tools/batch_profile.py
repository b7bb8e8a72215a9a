"""Where does the per-clip time of a sweep go?  cProfile of engine.Engine.analyze over short pinned clips."""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import bench
from pqa2_b200 import engine, model as M

class A: frames_per_step = 0
cx = bench.Ctx(A())
pool = bench.Pool(1920, 1080, 8, 32, 5, False, 0, resident=False)
model = M.resolve_model("vmaf_v0.6.1")
opt = engine.EngineOptions(devices=(0,))
with engine.Engine() as sess:
    sess.analyze(pool.clip(64), model, opt)
    for n in (300, 300, 512, 1024):
        t0 = time.perf_counter()
        for k in range(8):
            sess.analyze(pool.clip(n, offset=k), model, opt)
        dt = time.perf_counter() - t0
        print(f"{n} frames/clip: {1e3 * dt / 8:.1f} ms per clip -> {8 * n / dt:.0f} fps")
    class Inline:                      # run the shard on the profiled thread
        def __init__(self, target, args=(), daemon=None): self.t, self.a = target, args
        def start(self): self.t(*self.a)
        def join(self): pass
    real = engine.threading.Thread
    engine.threading.Thread = Inline
    pr = cProfile.Profile()
    pr.enable()
    for k in range(8):
        sess.analyze(pool.clip(300, offset=k), model, opt)
    pr.disable()
    engine.threading.Thread = real
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
    import threading
    # two clips at a time on one GPU (two contexts): does the fill/drain of one hide behind the other?
    def run(j):
        with engine.Engine() as s2:
            s2.analyze(pool.clip(64), model, opt)
            bar.wait()
            for k in range(8):
                s2.analyze(pool.clip(300, offset=k + j), model, opt)
    for nth in (1, 2, 3):
        bar = threading.Barrier(nth + 1)
        th = [threading.Thread(target=run, args=(j,)) for j in range(nth)]
        [t.start() for t in th]
        bar.wait(); t0 = time.perf_counter()
        [t.join() for t in th]
        dt = time.perf_counter() - t0
        print(f"{nth} concurrent sessions: {nth * 8 * 300 / dt:.0f} fps")
