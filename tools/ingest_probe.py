"""File ingest: does cudaHostRegister accept an mmap of a raw clip here (tmpfs and the container's disk), and what do the two
ingest paths deliver -- page cache -> GPU DMA (yuvio.MappedClip) vs reader threads -> pinned ring (engine._Prefetcher)?"""
import os, shutil, sys, time
import numpy as np
sys.path.insert(0, ".")
from pqa2_b200 import engine, model as M, synth, yuvio

w, h, n, P = 1920, 1080, int(os.environ.get("N", "300")), 32
pool = [synth.frame_pair(3, i, w, h, 8, chroma=True) for i in range(P)]
model = M.resolve_model("vmaf_v0.6.1")
for base in ("/dev/shm", "/tmp"):
    d = os.path.join(base, "ingest_probe")
    try:
        os.makedirs(d, exist_ok=True)
        if shutil.disk_usage(d).free < 2 * n * 3.2e6 + 1e9:
            print(base, "not enough space"); continue
        rp, dp = os.path.join(d, "r.y4m"), os.path.join(d, "d.y4m")
        yuvio.write_y4m(rp, (pool[i % P][0] for i in range(n)), w, h)
        yuvio.write_y4m(dp, (pool[i % P][1] for i in range(n)), w, h)
        ri, di = yuvio.probe(rp), yuvio.probe(dp)
        for all_planes in (False, True):
            opt = engine.EngineOptions(ffmpeg_psnr=all_planes, ffmpeg_ssim=all_planes)
            for mapped, threads in ((True, 0), (False, 4), (False, 8), (False, 12), (False, 16)):
                opt.reader_threads = max(1, threads)
                src = engine.FileSource(ri, di, mapped=mapped)
                mode = getattr(src._mapped()[0], "mode", "?") if src.zero_copy else f"ring, {threads} reader threads"
                if mapped and not src.zero_copy:
                    print(f"{base}: cudaHostRegister refused the mapping"); continue
                with engine.Engine() as sess:
                    best = 0.0
                    for rep in range(3):
                        t0 = time.perf_counter()
                        res = sess.analyze(src, model, opt)
                        best = max(best, n / (time.perf_counter() - t0))
                src.release()
                print(f"{base} {'Y+Cb+Cr + stats' if all_planes else 'Y only        '} {mode:34s}: {best:7.0f} fps  "
                      f"({best * 2 * (ri.frame_bytes if all_planes else w * h) / 1e9:5.1f} GB/s)  vmaf {res['pooled_metrics']['vmaf']['mean']:.6f}", flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)
