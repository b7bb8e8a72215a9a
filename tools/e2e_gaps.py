"""Is the GPU idle between launch groups when frames come from the host?  Sum of per-kernel CUDA-event times vs the
device-side stopwatch over the same 512 host-submitted frames, next to the same frames submitted from device memory."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from pqa2_b200 import _lib as L, engine, model as M, synth
from pqa2_b200.extractor import FeatureExtractor, pinned_empty, DeviceBuffer

w, h, n, P = 1920, 1080, 512, 32
model = M.resolve_model("vmaf_v0.6.1")
mask = engine.feature_mask(model, engine.EngineOptions())
ref = [pinned_empty((h, w), np.uint8) for _ in range(P)]
dis = [pinned_empty((h, w), np.uint8) for _ in range(P)]
dev = DeviceBuffer(2 * P * w * h, 0)
for i in range(P):
    rp, dp = synth.frame_pair(1, i, w, h, 8, chroma=False)
    ref[i][...] = rp[0]; dis[i][...] = dp[0]
    dev.upload(2 * i * w * h, ref[i]); dev.upload((2 * i + 1) * w * h, dis[i])
fx = FeatureExtractor(w, h, 8, 0, mask)
for mode in ("device", "host", "device", "host"):
    for prof in (False, True):
        fx.reset(); fx.set_profiling(prof); fx.kernel_profile(reset=True)
        t0 = time.perf_counter()
        fx.timer_mark(0)
        for i in range(n):
            k = i % P
            if mode == "host":
                fx.submit(i, [ref[k]], [dis[k]], L.FRAME_FIRST if i == 0 else 0)
            else:
                fx.submit_device(i, [(dev.ptr + 2 * k * w * h, w)], [(dev.ptr + (2 * k + 1) * w * h, w)], L.FRAME_FIRST if i == 0 else 0)
        fx.timer_mark(1)
        fx.flush()
        wall = time.perf_counter() - t0
        gpu = fx.timer_elapsed_ms()
        ksum = sum(ms for ms, _ in fx.kernel_profile(reset=True).values()) if prof else float("nan")
        print(f"{mode:6s} prof={prof}: wall {1e3 * wall:6.2f} ms  device stopwatch {gpu:6.2f} ms  kernel sum {ksum:6.2f} ms")
