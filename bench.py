#!/usr/bin/env python
"""bench.py -- VMAF frames/sec on B200 (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one batch of `frames_per_step` synthetic ref/dis frame
pairs per GPU (weak scaling: frame chunks are independent, there is no data-path collective --
SURVEY.md §8e).  `value` = frame pairs per second over all GPUs with the clips resident in HBM;
`e2e` = the same through the public engine call (pqa2_b200.engine.analyze) from pinned HOST frames,
host->device copies and the feature read-back inside the timed region: ONE call over the K steps' frames (20 x 512 = the
10k-frame clip of configs[1]); `e2e.one_call_per_step` is the same work as K separate 512-frame clips.

Headline workload (identical at every N): configs[1].  Next to it, in the same JSON line:
  workloads      (N = 1)  the integer v0.6.1 configurations BASELINE's metric names -- 1080p-int (configs[0] shape)
                          and 4k-int (configs[2]) -- each with value / e2e / roofline / cpu_baseline
  sharded_4k     (all N)  ONE 2160p 10-bit clip (configs[2]) split into contiguous frame chunks with a one-frame lead-in
                          over the N ranks (pqa2_b200.dist.analyze_distributed), rows gathered on rank 0; at N > 1 rank 0
                          re-scores the clip alone and the per-frame results must be bit-identical
  batch          (all N)  configs[4]: 64 clips of 300 1080p frames dealt over the ranks, one pooled report per clip
  file_e2e       (N = 1)  the drop-in call itself: VMAFAnalyzer.analyze_videos() on a 300-frame 1080p .y4m pair

Workloads (BASELINE.json `configs`):
  1080p-float  configs[1]: 1080p yuv420p 8-bit, vmaf_float_v0.6.1 + psnr + float_ssim + float_ms_ssim
  1080p-int    configs[0] shape: 1080p 8-bit, vmaf_v0.6.1 (integer extractors)
  4k-int       configs[2]: 2160p yuv420p10le, vmaf_4k_v0.6.1
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONTEXTS_PER_GPU = 3        # what engine.analyze picks for clips of >= 1536 frames per GPU (EngineOptions.contexts_per_device)
WORKLOADS = {
    "1080p-float": dict(w=1920, h=1080, bpc=8, model="vmaf_float_v0.6.1", psnr=True, ssim=True, ms_ssim=True,
                        frames_per_step=512, pool=64, cfg="configs[1]"),
    "1080p-int": dict(w=1920, h=1080, bpc=8, model="vmaf_v0.6.1", psnr=False, ssim=False, ms_ssim=False,
                      frames_per_step=512, pool=64, cfg="configs[0] shape"),
    "4k-int": dict(w=3840, h=2160, bpc=10, model="vmaf_4k_v0.6.1", psnr=False, ssim=False, ms_ssim=False,
                   frames_per_step=256, pool=16, cfg="configs[2]"),
}
SHARDED_FRAMES = 3600          # configs[2]: 2160p 60 fps, one minute
BATCH_CLIPS, BATCH_FRAMES = 64, 300      # configs[4]
FILE_FRAMES = 300              # configs[0]


# Algorithmic bytes per frame pair of each kernel at (w, h, bytes per sample): unique bytes the
# kernel must read + must write (DESIGN.md "Kernels").
def kernel_bytes(name: str, w: int, h: int, bps: int) -> float | None:
    px = w * h
    cpx = ((w + 1) // 2) * ((h + 1) // 2)
    lv = [((w >> s) * (h >> s)) for s in range(4)]                       # VIF level pixel counts
    ad = []
    cw, ch = w, h
    for _ in range(4):
        cw, ch = (cw + 1) // 2, (ch + 1) // 2
        ad.append(cw * ch)
    t = {
        "motion_blur": px * bps + px * 2,
        # SAD of consecutive blurred frames: every blur plane of the group is needed once (frame i is "current" for
        # pair i and "previous" for pair i+1 of the same launch; the second use hits L2 -- ncu: DRAM reads = 1 plane/frame)
        "motion_sad": px * 2,
        # scale 0 also writes pyramid level 1 and the motion feature's blurred reference (fused consumers of the raw luma)
        "vif_stat_s0": 2 * px * bps + 2 * lv[1] * 2 + px * 2, "vif_subsample_s1": 2 * px * bps + 2 * lv[1] * 2,
        "vif_stat_s1": 2 * lv[1] * 2, "vif_subsample_s2": 2 * lv[1] * 2 + 2 * lv[2] * 2,
        "vif_stat_s2": 2 * lv[2] * 2, "vif_subsample_s3": 2 * lv[2] * 2 + 2 * lv[3] * 2,
        "vif_stat_s3": 2 * lv[3] * 2,
        "adm_scale0": 2 * px * bps + 2 * ad[0] * 2, "adm_scale1": 2 * ad[0] * 2 + 2 * ad[1] * 4,
        "adm_scale2": 2 * ad[1] * 4 + 2 * ad[2] * 4, "adm_scale3": 2 * ad[2] * 4,
        "psnr_sse_y": 2 * px * bps, "psnr_sse_u": 2 * cpx * bps, "psnr_sse_v": 2 * cpx * bps,
        "ffssim_y": 2 * px * bps, "ffssim_u": 2 * cpx * bps, "ffssim_v": 2 * cpx * bps,
        # float extractors: fp32 pyramids / bands / blur
        "f_motion_blur": px * bps + px * 4, "f_motion_sad": px * 4,
        # f_vif_subsample_s1 also writes the motion feature's blurred reference (fused staging)
        "f_vif_stat_s0": 2 * px * bps, "f_vif_subsample_s1": 2 * px * bps + 2 * lv[1] * 4 + px * 4,
        "f_vif_stat_s1": 2 * lv[1] * 4, "f_vif_subsample_s2": 2 * lv[1] * 4 + 2 * lv[2] * 4,
        "f_vif_stat_s2": 2 * lv[2] * 4, "f_vif_subsample_s3": 2 * lv[2] * 4 + 2 * lv[3] * 4,
        "f_vif_stat_s3": 2 * lv[3] * 4,
        "f_adm_scale0": 2 * px * bps + 2 * ad[0] * 4, "f_adm_scale1": 2 * ad[0] * 4 + 2 * ad[1] * 4,
        "f_adm_scale2": 2 * ad[1] * 4 + 2 * ad[2] * 4, "f_adm_scale3": 2 * ad[2] * 4,
    }
    # float_ssim: box decimation by f = round(min(w, h) / 256), then the 11-tap maps on the small picture
    f = max(1, int(min(w, h) / 256.0 + 0.5))
    sp = (w // f + (w & 1)) * (h // f + (h & 1)) if f > 1 else px
    t["ssim_decimate"] = 2 * px * bps + 2 * sp * 4
    t["ssim_maps"] = 2 * sp * 4 if f > 1 else 2 * px * bps
    # float_ms_ssim: 5-level pyramid (9-tap low-pass, /2), maps on every level
    mw, mh = w, h
    for sc in range(5):
        if sc > 0:
            nw, nh = mw // 2 + (mw & 1), mh // 2 + (mh & 1)
            t[f"ms_ssim_lpf_s{sc}"] = (2 * mw * mh * (bps if sc == 1 else 4)) + 2 * nw * nh * 4
            if sc == 1 and f in (2, 4, 8):
                t["ms_ssim_lpf_s1"] += 2 * sp * 4          # it also writes float_ssim's decimated pair (fused staging)
            mw, mh = nw, nh
        t[f"ms_ssim_maps_s{sc}"] = 2 * mw * mh * (bps if sc == 0 else 4)
    return t.get(name)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


_TRAFFIC = None


def ncu_traffic(kernel: str | None, wname: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture of this workload (profiles/r02_traffic.json, else r01: {workload: {kernel: {"bytes": .., "frames": ..}}});
    kernel=None returns the workload's whole table."""
    global _TRAFFIC
    if _TRAFFIC is None:
        _TRAFFIC = {}
        for nm in ("r01_traffic.json", "r02_traffic.json"):          # later rounds override, workload by workload
            try:
                _TRAFFIC.update(json.load(open(os.path.join(ROOT, "profiles", nm))))
            except Exception:
                pass
    t = _TRAFFIC.get(wname)
    if t is None:
        return None
    return t if kernel is None else t.get(kernel)


def measured_peaks() -> tuple:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (CPU restatement of libvmaf's path; the reference's own
# ffmpeg+libvmaf cannot run in this image -- SURVEY.md §8c) on the host cores.
# --------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, f0, n, w, h, bpc, kind = args
    import oracle
    from pqa2_b200 import synth
    frames = [synth.frame_pair(seed, f0 + i, w, h, bpc, chroma=False) for i in range(n + 1)]
    t0 = time.perf_counter()
    prev = oracle.motion_blur(frames[0][0][0], bpc)
    for i in range(1, n + 1):
        rp, dp = frames[i]
        if kind == "float":
            oracle.float_features(rp[0], dp[0], bpc, prev_ref=frames[i - 1][0][0], psnr=True, ssim=True, ms_ssim=True)
        else:
            blur = oracle.motion_blur(rp[0], bpc)
            oracle.motion_sad(blur, prev)
            prev = blur
            oracle.vif(rp[0], dp[0], bpc)
            oracle.adm(rp[0], dp[0], bpc)
    return time.perf_counter() - t0


def cpu_baseline(wl: dict, frames_per_core: int = 0) -> dict:
    import multiprocessing as mp
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    kind = "float" if "float" in wl["model"] else "int"
    if frames_per_core <= 0:
        # ~10-20 s of CPU work per core: the C oracle needs ~0.6 s (float + ssim + ms-ssim) or ~0.3 s
        # (integer) per 1080p frame pair, 4x that at 4K
        per_frame = (0.6 if kind == "float" else 0.3) * (wl["w"] * wl["h"]) / (1920 * 1080)
        frames_per_core = max(2, min(24, int(round(12.0 / per_frame))))
    jobs = [(1, 10 * c, frames_per_core, wl["w"], wl["h"], wl["bpc"], kind) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, [(1, 0, 0, 64, 64, 8, kind)] * cores)           # start workers, load the oracle
        t0 = time.perf_counter()
        busy = pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
    n = cores * frames_per_core
    return {"value": n / max(busy), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{n} synthetic {wl['w']}x{wl['h']} {wl['bpc']}-bit frame pairs ({frames_per_core} per core), "
                      f"oracle/ ({'float' if kind == 'float' else 'integer'} extractors, C -O3, row-wise filter loops vectorised by gcc with an AVX2 clone); "
                      f"wall {dt:.1f}s incl. frame synthesis"}


def reference_arm(args, wl, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    times = []
    base = None
    for s in range(args.warmup + args.steps):
        base = cpu_baseline(wl, frames_per_core=max(1, min(4, int(round(3.0 / ((1.1 if 'float' in wl['model'] else 0.5) * wl['w'] * wl['h'] / (1920 * 1080)))))))
        if s >= args.warmup:
            times.append(base["value"])
    v = statistics.mean(times) if times else 0.0
    base["value"] = v
    out = {"impl": "reference", "metric": "vmaf_frames_per_sec", "value": v, "unit": "frames/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1000.0 * (os.cpu_count() or 1) / v if v else None, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": wl_dtype(wl), "data": "synthetic",
           "config": bench_config(wl, wname, args.frames_per_step or wl["frames_per_step"], args.gpus),
           "cpu_baseline": base, "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "CPU restatement of the libvmaf path (oracle/), all host cores; ffmpeg+libvmaf is not installable here"}
    print(json.dumps(out), flush=True)
    return 0


def bind_to_gpu_numa_node(local: int) -> str:
    """One process per GPU: run on (and therefore first-touch the pinned frame buffers from) the CPUs of the NUMA node the
    GPU hangs off, so that the H2D stream of rank r does not cross the socket interconnect.  Best effort; returns a note."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out                    # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "numa: single node"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa: rank bound to node {node} ({len(cpus)} cpus)"
    except Exception as e:          # noqa: BLE001
        return f"numa: not bound ({type(e).__name__})"
    return "numa: not bound"


def wl_dtype(wl) -> str:
    return "f32" if "float" in wl["model"] else "int32/int64 fixed-point"


def bench_config(wl, wname, fps, n_gpus, contexts=None) -> dict:
    contexts = CONTEXTS_PER_GPU if contexts is None else max(1, contexts)
    chroma = False          # libvmaf psnr=1 is psnr_y (enable_chroma=false through FFmpeg): every enabled feature reads luma
    bps = 1 if wl["bpc"] == 8 else 2
    mb = wl["pool"] * 2 * wl["w"] * wl["h"] * bps * (1.5 if chroma else 1.0) / 1e6
    return {"workload": f"{wname}: {wl['cfg']} -- {wl['w']}x{wl['h']} yuv420p {wl['bpc']}-bit, model {wl['model']}"
                        + (" + psnr" if wl["psnr"] else "") + (" + float_ssim" if wl["ssim"] else "")
                        + (" + float_ms_ssim" if wl["ms_ssim"] else ""),
            "frames_per_step_per_gpu": fps, "pool_frames_per_gpu": wl["pool"],
            "planes": "Y, Cb, Cr (psnr=1 reads all three)" if chroma else "Y (every enabled feature reads luma only)",
            "l2_policy": "inputs larger than L2: each step streams the whole resident pool "
                         f"({wl['pool']} frame pairs, {mb:.0f} MB) through the kernels",
            "contexts_per_gpu": contexts,
            "parallelism": f"frame-sharded x{n_gpus}, no collective; {contexts} context(s) side by side on each GPU, a "
                           "contiguous share of the frames each (the engine's choice for a clip of this length)"}


# --------------------------------------------------------------------------------------------
class Ctx:
    """Per-process state shared by the measurements: ranks, the optional process group, helpers."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.numa_note = None

    def init_dist(self):
        if self.world > 1:
            self.numa_note = bind_to_gpu_numa_node(self.local)
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, v: float) -> float:
        if self.dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{self.local}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


class Pool:
    """P synthetic frame pairs: pinned on the host (e2e) and, optionally, resident in HBM (value)."""

    def __init__(self, w, h, bpc, P, seed, chroma: bool, device: int, resident: bool = True):
        from concurrent.futures import ThreadPoolExecutor
        from pqa2_b200 import synth
        from pqa2_b200.extractor import DeviceBuffer, pinned_empty
        self.w, self.h, self.bpc, self.P, self.chroma = w, h, bpc, P, chroma
        self.bps = 1 if bpc == 8 else 2
        dtype = np.uint8 if bpc == 8 else np.uint16
        shapes = [(h, w)] + ([((h + 1) // 2, (w + 1) // 2)] * 2 if chroma else [])
        self.shapes = shapes
        self.plane_bytes = [a * b * self.bps for a, b in shapes]
        self.frame_bytes = sum(self.plane_bytes)
        self.ref = [[pinned_empty(s, dtype) for s in shapes] for _ in range(P)]
        self.dis = [[pinned_empty(s, dtype) for s in shapes] for _ in range(P)]

        def fill(i):
            rp, dp = synth.frame_pair(seed, i, w, h, bpc, chroma=chroma)
            for k in range(len(shapes)):
                self.ref[i][k][...] = rp[k]
                self.dis[i][k][...] = dp[k]

        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            list(ex.map(fill, range(P)))
        self.dev = None
        if resident:
            self.dev = DeviceBuffer(2 * P * self.frame_bytes, device)
            for i in range(P):
                for clip, planes in ((0, self.ref[i]), (1, self.dis[i])):
                    off = (2 * i + clip) * self.frame_bytes
                    for k, p in enumerate(planes):
                        self.dev.upload(off, p)
                        off += self.plane_bytes[k]

    def dev_planes(self, i: int, clip: int):
        off = (2 * (i % self.P) + clip) * self.frame_bytes
        out = []
        for k, (ph, pw) in enumerate(self.shapes):
            out.append((self.dev.ptr + off, pw * self.bps))
            off += self.plane_bytes[k]
        return out

    def clip(self, n: int, offset: int = 0, stride: int = 1):
        """A zero-copy engine.FrameSource of n frames cycling through the pinned pool."""
        from pqa2_b200 import engine
        pool = self

        class PinnedClip(engine.FrameSource):
            zero_copy = True

            def __init__(self):
                self.width, self.height, self.bpc = pool.w, pool.h, pool.bpc
                self.chroma, self.nb_frames, self.fps = (420 if pool.chroma else 0), n, 30.0

            def get(self, i, luma_only):
                k = (offset + i * stride) % pool.P
                return (pool.ref[k][:1], pool.dis[k][:1]) if luma_only else (pool.ref[k], pool.dis[k])

        return PinnedClip()

    def free(self):
        if self.dev is not None:
            self.dev.free()
        self.ref = self.dis = None


def measure_workload(cx: Ctx, wname: str, wl: dict, pool: Pool, steps: int, warmup: int, with_clocks: bool,
                     with_e2e: bool) -> dict:
    """Resident `value` (CUDA events on the compute stream), the per-kernel pass with the roofline of the top kernel, and the
    host-buffer `e2e` of one workload on this rank's GPU."""
    from pqa2_b200 import _lib as L
    from pqa2_b200 import engine, model as M
    from pqa2_b200.extractor import FeatureExtractor
    args, local, world = cx.args, cx.local, cx.world
    w, h, bpc = wl["w"], wl["h"], wl["bpc"]
    bps = pool.bps
    fps_step = args.frames_per_step or wl["frames_per_step"]
    model = M.resolve_model(wl["model"])
    fast = bool(getattr(args, "fast_float", False)) and model.is_float
    opt = engine.EngineOptions(psnr=wl["psnr"], ssim=wl["ssim"], ms_ssim=wl["ms_ssim"], devices=(local,), fast_float=fast)
    mask = engine.feature_mask(model, opt)
    fx = FeatureExtractor(w, h, bpc, 420 if pool.chroma else 0, mask, local, vif_enhn_gain_limit=model.vif_enhn_gain_limit,
                          adm_enhn_gain_limit=model.adm_enhn_gain_limit, fast_float=fast)
    counter = [0]
    # bytes of one frame of one clip that the enabled features actually read (luma only unless psnr / ssim stats want chroma)
    uses_chroma = pool.chroma and bool(mask & (L.FEAT_PSNR_UV | L.FEAT_FFSSIM))
    frame_bytes = pool.frame_bytes if uses_chroma else pool.plane_bytes[0]

    def run_step():
        for _ in range(fps_step):
            i = counter[0]
            fx.submit_device(i, pool.dev_planes(i, 0), pool.dev_planes(i, 1), L.FRAME_FIRST if i == 0 else 0)
            counter[0] += 1

    # ---- resident run: W warm-up steps, then exactly K timed steps.  As the engine does for a clip of this length
    # (EngineOptions.contexts_per_device), the steps' frames are dealt over CONTEXTS_PER_GPU contexts that run side by side
    # on this GPU, each a contiguous share submitted by its own host thread; the stopwatch is a pair of CUDA events on the
    # first context's compute stream: the first before anything is submitted, the second after every context has drained.
    others = [FeatureExtractor(w, h, bpc, 420 if pool.chroma else 0, mask, local, vif_enhn_gain_limit=model.vif_enhn_gain_limit,
                               adm_enhn_gain_limit=model.adm_enhn_gain_limit, fast_float=fast)
              for _ in range(max(1, args.contexts) - 1)]
    ctxs = [fx] + others

    def run_share(fxk, first, count):
        for i in range(first, first + count):
            fxk.submit_device(i, pool.dev_planes(i, 0), pool.dev_planes(i, 1), L.FRAME_FIRST if i == first else 0)
        fxk.kick()                          # the share's last (partial) launch group

    def run_all(total):
        share = (total + len(ctxs) - 1) // len(ctxs)
        ths = [threading.Thread(target=run_share, args=(c, k * share, max(0, min(share, total - k * share))))
               for k, c in enumerate(ctxs)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for c in ctxs:
            c.flush()

    run_all(fps_step * max(warmup, 3))
    for c in ctxs:
        c.reset()
    sampler = ClockSampler(local) if with_clocks else None
    cx.barrier()
    launches0 = sum(c.kernel_launches for c in ctxs)
    if sampler:
        sampler.start()
    fx.timer_mark(0)
    run_all(fps_step * steps)
    fx.timer_mark(1)
    fx.flush()
    ms = fx.timer_elapsed_ms()
    cx.barrier()
    clocks = sampler.stop() if sampler else None
    launches = sum(c.kernel_launches for c in ctxs) - launches0
    ms = cx.max_over_ranks(ms)
    value = fps_step * steps * world / (ms / 1000.0)
    for c in others:
        c.close()
    fx.reset()

    # ---- per-kernel profile over the same region (events around every launch; a separate pass so the
    #      event overhead stays out of `value`)
    fx.set_profiling(True)
    fx.kernel_profile(reset=True)
    for _ in range(min(steps, 10)):
        run_step()
    fx.flush()
    prof = fx.kernel_profile(reset=True)
    fx.set_profiling(False)
    peak, peak_src = measured_peaks()
    B = fx.batch_frames
    kernels = {}
    for nm, (kms, cnt) in prof.items():
        per_launch_ms = kms / cnt
        by = kernel_bytes(nm, w, h, bps)
        kernels[nm] = {"ms_per_launch": per_launch_ms, "launches": cnt,
                       "gbps": (by * B / (per_launch_ms * 1e-3) / 1e9) if by else None}
    tot_ms = sum(k["ms_per_launch"] for k in kernels.values()) or 1.0
    top = max(kernels, key=lambda n: kernels[n]["ms_per_launch"]) if kernels else None
    roofline = None
    floor = 2 * frame_bytes                   # every input sample of the pair read once
    if top:
        by = kernel_bytes(top, w, h, bps)
        ach = kernels[top]["gbps"]
        tr = ncu_traffic(top, wname)
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": (ach / peak) if ach else None,
                    "traffic": (tr["bytes"] * B / tr["frames"]) if tr else None,
                    "traffic_source": (tr.get("source") if tr else None), "peak_source": peak_src,
                    "pipe_note": (tr.get("pipe_note") if tr else None),
                    "algorithmic_bytes_per_launch": by * B if by else None,
                    "ms_per_launch": kernels[top]["ms_per_launch"], "frames_per_launch": B,
                    "share_of_step": kernels[top]["ms_per_launch"] / tot_ms,
                    "pipeline_floor_gbps": floor * value / world / 1e9,
                    "pipeline_floor_frac": floor * value / world / 1e9 / peak}
        # whole-pipeline DRAM traffic per frame pair (sum over the committed ncu capture of every kernel of this
        # workload) against the floor of reading every input sample once
        table = ncu_traffic(None, wname)
        if table and all(k in table for k in kernels):
            per_pair = sum(table[k]["bytes"] / table[k]["frames"] for k in kernels)
            roofline["pipeline_dram_bytes_per_pair"] = per_pair
            roofline["pipeline_dram_amplification"] = per_pair / floor
        # The stencil kernels sit on the instruction-issue roof, not on HBM: report that roof too.
        # Executed warp instructions per launch come from the committed ncu capture of this kernel; the rate is
        # measured live (CUDA-event launch time); peak = SMs x 4 schedulers x 1 warp instruction per clock.
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        if tr and tr.get("warp_inst"):
            wi = tr["warp_inst"] * B / tr["frames"]
            peak_issue = 148 * 4 * sm_mhz * 1e6
            roofline["issue"] = {"warp_instructions_per_launch": wi,
                                 "achieved_gwarp_inst_s": wi / (kernels[top]["ms_per_launch"] * 1e-3) / 1e9,
                                 "peak_gwarp_inst_s": peak_issue / 1e9,
                                 "frac": wi / (kernels[top]["ms_per_launch"] * 1e-3) / peak_issue,
                                 "thread_instructions_per_pixel": wi * 32 / (B * w * h)}
    fx.close()

    # ---- end to end through the public engine call: pinned host frames in, scores out
    e2e = None
    if with_e2e:
        # one engine session; the timed region is ONE engine.analyze call over the K steps' frames -- the call a user makes
        # for a clip of that length (configs[1] is one 10k-frame pair: 20 steps x 512 frames) -- with, for every step's
        # frames, H2D of every plane the enabled features read, the kernels and D2H of the feature rows as the launch groups
        # complete, and SVR fusion + log entries + pooling of the whole clip before the call returns.
        k_e2e = max(1, steps)
        with engine.Engine() as sess:
            # warm-up (allocations, first launches): a short clip (one context) and one long enough for the engine to
            # bring up the contexts it runs side by side on a long clip (EngineOptions.contexts_per_device)
            sess.analyze(pool.clip(fps_step), model, opt)
            sess.analyze(pool.clip(min(fps_step * k_e2e, max(fps_step, 2048))), model, opt)
            cx.barrier()
            t0 = time.perf_counter()
            res = sess.analyze(pool.clip(fps_step * k_e2e), model, opt)
            dt = time.perf_counter() - t0
            dt = cx.max_over_ranks(dt)
            # the same frames as separate short clips (one call per step): what pipeline fill, drain and scoring cost per call
            k_short = max(1, min(steps, 10))
            cx.barrier()
            t0 = time.perf_counter()
            for _ in range(k_short):
                res_s = sess.analyze(pool.clip(fps_step), model, opt)
            dt_s = cx.max_over_ranks(time.perf_counter() - t0)
        n_e2e = fps_step * k_e2e
        e2e = {"value": n_e2e * world / dt, "unit": "frames/s",
               "h2d_bytes_per_step": int(2 * frame_bytes * fps_step),
               "d2h_bytes_per_step": int((L.BV_RAW_WORDS * 8 + 64 * 8) * fps_step + 8 * fps_step),
               "frames": n_e2e, "steps": k_e2e, "calls": 1, "ms_per_step": 1000.0 * dt / k_e2e,
               "timer": "host wall clock around one engine.analyze call over steps x frames_per_step pinned host frames (H2D, "
                        "kernels, feature D2H per launch group, SVR, log entries, pooling), max over ranks",
               "pooled_vmaf_mean": res["pooled_metrics"]["vmaf"]["mean"],
               "one_call_per_step": {"value": fps_step * k_short * world / dt_s, "unit": "frames/s", "calls": k_short,
                                     "frames_per_call": fps_step,
                                     "pooled_vmaf_mean": res_s["pooled_metrics"]["vmaf"]["mean"]}}
        if cx.numa_note:
            e2e["host_placement"] = cx.numa_note
    return {"value": value, "ms_per_step": ms / steps, "steps": steps, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "e2e": e2e, "frames_per_launch": B,
            "kernels": {k: {"ms_per_launch": round(v["ms_per_launch"], 5), "gbps": v["gbps"] and round(v["gbps"], 1),
                            "frac_of_hbm_peak": v["gbps"] and round(v["gbps"] / peak, 4)} for k, v in kernels.items()}}


def measure_sharded_4k(cx: Ctx, pool: Pool) -> dict:
    """configs[2] as the north star partitions it: ONE 2160p 10-bit clip, contiguous frame chunks with a one-frame
    lead-in, one rank per GPU, per-frame rows gathered on rank 0 (motion2 across the chunk borders, SVR and pooling
    there).  Host frames in, pooled report out: an end-to-end, strong-scaling number."""
    from pqa2_b200 import dist as D
    from pqa2_b200 import engine, model as M
    wl = WORKLOADS["4k-int"]
    model = M.resolve_model(wl["model"])
    opt = engine.EngineOptions(devices=(cx.local,))
    n = cx.args.sharded_frames
    clip = pool.clip(n)
    times, shard_s, res = [], [], None
    with engine.Engine() as sess:
        def timed_shard(src, model_, opt_, device, start, end, mask):      # this rank's chunk: H2D + kernels + row read-back
            t0 = time.perf_counter()
            out = D._default_shard_fn(src, model_, opt_, device, start, end, mask, sess)
            shard_s.append(time.perf_counter() - t0)
            return out

        D.analyze_distributed(pool.clip(min(n, 64 * cx.world)), model, opt, device=cx.local, session=sess,
                              svr_device=cx.local)                 # warm-up: contexts, first launches, first gather
        # chunk lengths proportional to what each rank sustains while all ranks upload at once (dist.calibrate): the
        # box's GPUs do not see the same host-to-device bandwidth under load (tools/h2d_concurrent.py)
        weights = D.calibrate(clip, model, opt, device=cx.local, session=sess) if cx.world > 1 else None
        equal_times = []
        for policy in (("equal", None), ("weighted", weights)) if cx.world > 1 else (("equal", None),):
            for _ in range(3 if policy[0] == "weighted" or cx.world == 1 else 2):
                cx.barrier()
                t0 = time.perf_counter()
                r = D.analyze_distributed(clip, model, opt, device=cx.local, shard_fn=timed_shard, svr_device=cx.local,
                                          weights=policy[1])
                dt = time.perf_counter() - t0        # rank 0 returns after the gather, the SVR and the pooling
                (equal_times if policy[0] == "equal" and cx.world > 1 else times).append(cx.max_over_ranks(dt))
                res = r if r is not None else res
        # the same chunks with the frames already resident in HBM (device-timed, max over ranks): the compute path alone
        from pqa2_b200 import _lib as L
        from pqa2_b200.extractor import FeatureExtractor
        a, b = D.rank_range(n, cx.rank, cx.world)
        with FeatureExtractor(pool.w, pool.h, pool.bpc, 0, engine.feature_mask(model, opt), cx.local) as fx:
            for rep in range(2):
                fx.reset()
                cx.barrier()
                fx.timer_mark(0)
                for i in range(max(a - 1, 0), b):
                    fx.submit_device(i, pool.dev_planes(i, 0), pool.dev_planes(i, 1),
                                     (L.FRAME_FIRST if i == max(a - 1, 0) else 0) | (L.FRAME_LEAD_IN if i < a else 0))
                fx.timer_mark(1)
                fx.flush()
                resident_s = cx.max_over_ranks(fx.timer_elapsed_ms() / 1000.0)
        rec = None
        if cx.rank == 0:
            dt = min(times)
            per_frame = 2 * pool.frame_bytes
            rec = {"value": n / dt, "unit": "frames/s", "frames": n, "n_gpus": cx.world, "scaling": "strong",
                   "seconds": dt, "runs": [round(t, 4) for t in times], "rank0_shard_seconds": [round(t, 4) for t in shard_s],
                   "resident_value": n / resident_s,
                   "resident_note": "the same contiguous chunks + lead-in frames with the clip already in HBM, device-timed, "
                                    "max over ranks: the compute path without the host-to-device ingest",
                   "workload": f"configs[2]: ONE 3840x2160 yuv420p10le clip of {n} frames (a pinned pool of {pool.P} distinct "
                               f"frame pairs, cycled), {wl['model']}, contiguous frame chunks + one-frame lead-in per rank, "
                               "rows gathered on rank 0",
                   "h2d_bytes_total": int(per_frame * (n + cx.world - 1)), "h2d_gbps_per_gpu": per_frame * n / cx.world / dt / 1e9,
                   "pooled_vmaf_mean": res["pooled_metrics"]["vmaf"]["mean"],
                   "timer": "host wall clock around dist.analyze_distributed (H2D of every frame, kernels, row gather, "
                            "motion2 / SVR / pooling on rank 0); max over ranks, best of 2"}
            if cx.world > 1:
                rec["split"] = "chunk lengths proportional to each rank's calibrated rate (dist.calibrate)"
                rec["calibrated_fps_per_rank"] = [round(x, 1) for x in weights]
                rec["equal_split_value"] = n / min(equal_times)
                rec["equal_split_runs"] = [round(t, 4) for t in equal_times]
                # the same clip on rank 0's GPU alone: every per-frame value must be bit-identical (integer accumulators)
                t0 = time.perf_counter()
                one = sess.analyze(clip, model, opt)
                rec["n1_same_run_fps"] = n / (time.perf_counter() - t0)
                a, b = res["rows"].arr, one["rows"].arr
                same_rows = all(np.array_equal(a[f], b[f]) for f in ("raw", "motion", "vif_scale", "adm_scale", "adm2"))
                va = [fr["metrics"]["vmaf"] for fr in res["frames"]]
                vb = [fr["metrics"]["vmaf"] for fr in one["frames"]]
                m2a = [fr["metrics"]["integer_motion2"] for fr in res["frames"]]
                m2b = [fr["metrics"]["integer_motion2"] for fr in one["frames"]]
                rec["identical_to_n1"] = bool(same_rows and va == vb and m2a == m2b and
                                              res["pooled_metrics"] == one["pooled_metrics"])
                if not rec["identical_to_n1"]:
                    log("[bench] ERROR: frame-sharded results differ from the single-GPU run; the sharded value is withdrawn")
                    rec["value"] = None
    cx.barrier()
    return rec


def measure_batch(cx: Ctx, pool: Pool) -> dict:
    """configs[4]: a sweep of 64 clip pairs (300 frames of 1080p each), whole clips dealt over the ranks, one pooled
    report per clip gathered on rank 0."""
    from pqa2_b200 import dist as D
    from pqa2_b200 import engine, model as M
    model = M.resolve_model("vmaf_v0.6.1")
    opt = engine.EngineOptions(devices=(cx.local,))
    nclips, nfr = cx.args.batch_clips, BATCH_FRAMES
    clips = [pool.clip(nfr, offset=5 * k, stride=1 + k % 3) for k in range(nclips)]
    with engine.Engine() as sess, engine.Engine() as sess2, engine.Engine() as sess3:
        # warm-up: the contexts and first launches of the workers' sessions, and the process group's first object gather (a
        # fresh NCCL communicator sets up its channels lazily -- about a second at 8 ranks -- which is not the sweep's cost)
        D.analyze_batch_distributed([pool.clip(64) for _ in range(3 * cx.world)], model, opt, device=cx.local,
                                    session=[sess, sess2, sess3])
        cx.barrier()
        t0 = time.perf_counter()
        out = D.analyze_batch_distributed(clips, model, opt, device=cx.local, session=[sess, sess2, sess3])
        dt = cx.max_over_ranks(time.perf_counter() - t0)
    if cx.rank != 0:
        return None
    errs = [o for o in out if o is None or "error" in o]
    means = [o["pooled_metrics"]["vmaf"]["mean"] for o in out if o and "pooled_metrics" in o]
    return {"value": (nclips * nfr / dt) if not errs else None, "unit": "frames/s", "clips": nclips, "frames_per_clip": nfr,
            "clips_per_s": nclips / dt, "seconds": dt, "n_gpus": cx.world, "scaling": "strong", "failed_clips": len(errs),
            "workload": f"configs[4]: {nclips} clips x {nfr} frames 1920x1080 8-bit (pinned pool of {pool.P} frame pairs, a "
                        "different walk per clip), vmaf_v0.6.1, clip k on rank k % N, three clips in flight per rank (a session each)",
            "pooled_reports": len(means), "sum_of_pooled_vmaf_means": float(np.sum(np.array(means))) if means else None,
            "h2d_bytes_total": int(2 * pool.plane_bytes[0] * nclips * nfr),
            "timer": "host wall clock around dist.analyze_batch_distributed, max over ranks"}


def measure_file_e2e(cx: Ctx, pool: Pool) -> dict:
    """The call the reference makes (app/vmaf_analyzer.py:242): VMAFAnalyzer.analyze_videos(ref_path, dis_path) on a 300-frame
    1080p yuv420p .y4m pair (configs[0]), all planes, with the reference's defaults (psnr / ssim stats files on)."""
    from pqa2_b200 import yuvio
    from pqa2_b200.vmaf_analyzer import VMAFAnalyzer
    n = FILE_FRAMES
    need = 2 * n * (pool.frame_bytes + 6) + (64 << 20)
    base = None
    for cand in ("/dev/shm", os.environ.get("TMPDIR") or "/tmp"):
        try:
            if os.path.isdir(cand) and shutil.disk_usage(cand).free > need:
                base = cand
                break
        except OSError:
            continue
    if base is None:
        return {"value": None, "note": "no scratch space for the clip pair"}
    d = os.path.join(base, f"b200vmaf_bench_{os.getpid()}")
    os.makedirs(d, exist_ok=True)
    try:
        rp, dp = os.path.join(d, "ref_1920x1080.y4m"), os.path.join(d, "dis_1920x1080.y4m")
        t0 = time.perf_counter()
        yuvio.write_y4m(rp, (pool.ref[i % pool.P] for i in range(n)), pool.w, pool.h, pool.bpc)
        yuvio.write_y4m(dp, (pool.dis[i % pool.P] for i in range(n)), pool.w, pool.h, pool.bpc)
        t_write = time.perf_counter() - t0
        a = VMAFAnalyzer()
        a.set_output_directory(d)
        a.set_test_name("bench")
        a.set_devices((cx.local,))
        errs = []
        a.error_occurred.connect(errs.append)
        runs = []
        res = None
        for _ in range(3):
            t0 = time.perf_counter()
            res = a.analyze_videos(rp, dp, "vmaf_v0.6.1")
            runs.append(time.perf_counter() - t0)
            if res is None:
                return {"value": None, "note": "analyze_videos failed: " + "; ".join(errs)}
        ingest = "mmap + cudaHostRegister (page cache -> GPU, no CPU copy)" if getattr(a, "last_ingest", "") == "mapped" \
            else "reader threads (auto: 12 on a 16-core host) -> pinned ring -> cudaMemcpyAsync"
        return {"value": n / min(runs[1:]), "unit": "frames/s", "frames": n, "first_call_fps": n / runs[0],
                "runs_s": [round(t, 4) for t in runs], "ingest": ingest, "clip_dir": base,
                "bytes_read_per_call": int(2 * n * pool.frame_bytes),
                "call": "VMAFAnalyzer.analyze_videos(ref.y4m, dis.y4m, 'vmaf_v0.6.1') with the reference defaults: libvmaf JSON + "
                        "FFmpeg psnr and ssim stats files over Y, Cb, Cr; file read, H2D, kernels, D2H, SVR, pooling and the "
                        "three output files inside the timed region; best of calls 2-3 (call 1 also creates the CUDA context)",
                "vmaf_score": res["vmaf_score"], "psnr_log": bool(res["psnr_log"]), "ssim_log": bool(res["ssim_log"]),
                "write_clip_pair_s": round(t_write, 2)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def two_gpus_in_one_process_check(cx: Ctx, pool: Pool) -> dict:
    """The driver's GPU test box has one GPU, so tests/test_gpu_integer.py::test_two_gpus_in_one_process_match_one_gpu is
    skipped there; with N >= 2 GPUs visible rank 0 runs the same check here: one process, one thread + context per GPU."""
    from pqa2_b200 import engine, model as M
    model = M.resolve_model("vmaf_v0.6.1")
    out = {"devices": [0, 1]}
    # 96 frames: two fixed shares; 4096 frames: chunks of 512 pulled from a shared counter (engine.analyze dynamic_chunk)
    for n in (96, 4096):
        clip = pool.clip(n)
        one = engine.analyze(clip, model, engine.EngineOptions(devices=(0,)))
        t0 = time.perf_counter()
        two = engine.analyze(clip, model, engine.EngineOptions(devices=(0, 1)))
        dt = time.perf_counter() - t0
        same = [fr["metrics"] for fr in one["frames"]] == [fr["metrics"] for fr in two["frames"]] and \
            np.array_equal(one["rows"].arr["raw"], two["rows"].arr["raw"])
        out[f"identical_{n}_frames"] = bool(same)
        out[f"fps_{n}_frames_cold"] = n / dt          # includes creating both contexts: no session here
    out["identical"] = all(v for k, v in out.items() if k.startswith("identical_"))
    return out


# --------------------------------------------------------------------------------------------
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (profiling runs)")
    ap.add_argument("--frames-per-step", type=int, default=0)
    ap.add_argument("--fast-float", action="store_true", help="float workloads with bv_opts.fast_float (opt-in build)")
    ap.add_argument("--sharded-frames", type=int, default=SHARDED_FRAMES)
    ap.add_argument("--batch-clips", type=int, default=BATCH_CLIPS)
    ap.add_argument("--contexts", type=int, default=CONTEXTS_PER_GPU,
                    help="contexts side by side on each GPU in the resident run (1: one stream of launches, for ncu captures)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    cx = Ctx(args)
    rank, world, local = cx.rank, cx.world, cx.local

    from pqa2_b200 import _lib as L
    from pqa2_b200.extractor import FeatureExtractor

    wname = args.workload
    if wname == "auto":
        wname = "1080p-float"
    wl = WORKLOADS[wname]

    if args.impl == "reference":
        return reference_arm(args, wl, wname)

    lib = L.load()
    if lib.bv_device_count() <= local:
        raise SystemExit("bench.py: no CUDA device for this rank (the engine has no CPU fallback)")
    if wname == "1080p-float" and args.workload == "auto":
        # the float extractors must exist; otherwise say so loudly instead of measuring something else
        try:
            FeatureExtractor(64, 64, 8, 0, L.FEAT_VMAF_FLOAT, local).close()
        except Exception as e:
            log(f"[bench] float extractors unavailable ({e}); measuring the integer workload 1080p-int instead")
            wname = "1080p-int"
            wl = WORKLOADS[wname]
    cx.init_dist()

    # ---- headline workload
    t0 = time.perf_counter()
    # the same clip on every rank; at N = 1 with chroma planes, which only the file_e2e record (FFmpeg psnr / ssim stats
    # files over Y, Cb, Cr) reads -- libvmaf's psnr=1 is psnr_y, so the headline uploads luma only
    pool = Pool(wl["w"], wl["h"], wl["bpc"], wl["pool"], 100, world == 1 and args.workload == "auto" and not args.no_extras, local)
    if rank == 0:
        log(f"[bench] synthesised {wl['pool']} frame pairs {wl['w']}x{wl['h']} {wl['bpc']}-bit in {time.perf_counter() - t0:.1f}s")
    head = measure_workload(cx, wname, wl, pool, args.steps, args.warmup, True, not args.no_e2e)
    fps_step = args.frames_per_step or wl["frames_per_step"]
    if rank == 0:
        log(f"[bench] {wname}: {head['value']:.0f} fps resident" + (f", {head['e2e']['value']:.0f} e2e" if head["e2e"] else ""))

    extras = not args.no_extras and args.workload == "auto"
    workloads, sharded, batch, file_e2e, twogpu = {}, None, None, None, None
    pools = {wname: pool}

    def get_pool(name, seed, chroma):
        if name not in pools:
            w_ = WORKLOADS[name]
            t0 = time.perf_counter()
            pools[name] = Pool(w_["w"], w_["h"], w_["bpc"], w_["pool"], seed, chroma, local)
            if rank == 0:
                log(f"[bench] synthesised {w_['pool']} frame pairs {w_['w']}x{w_['h']} {w_['bpc']}-bit in {time.perf_counter() - t0:.1f}s")
        return pools[name]

    def guarded(label, fn):
        try:
            return fn()
        except Exception as e:          # noqa: BLE001  (an extra record never costs the headline line)
            log(f"[bench] {label} failed: {type(e).__name__}: {e}")
            if world > 1:
                raise                   # ranks would lose step with each other: fail loudly instead of hanging
            return {"value": None, "error": f"{type(e).__name__}: {e}"}

    if extras:
        steps_x = max(2, min(args.steps, 10))
        if world == 1:
            # the integer v0.6.1 configurations of BASELINE's metric, each with its own roofline and CPU baseline
            r = guarded("1080p-int", lambda: measure_workload(cx, "1080p-int", WORKLOADS["1080p-int"], pool, steps_x,
                                                              args.warmup, False, not args.no_e2e))
            workloads["1080p-int"] = r
            log(f"[bench] 1080p-int: {r.get('value')} fps resident")
            file_e2e = guarded("file_e2e", lambda: measure_file_e2e(cx, pool))
            log(f"[bench] file_e2e: {file_e2e}")
        batch = guarded("batch", lambda: measure_batch(cx, pool))
        if rank == 0:
            log(f"[bench] batch: {batch}")
        if world > 1 and rank == 0 and lib.bv_device_count() >= 2:
            twogpu = guarded("two_gpus_in_one_process", lambda: two_gpus_in_one_process_check(cx, pool))
        cx.barrier()
        pool.free()                     # the 4K pool wants the pinned memory and HBM
        p4 = get_pool("4k-int", 7, False)           # the same clip on every rank
        if world == 1:
            r = guarded("4k-int", lambda: measure_workload(cx, "4k-int", WORKLOADS["4k-int"], p4, steps_x, args.warmup,
                                                           False, not args.no_e2e))
            workloads["4k-int"] = r
            log(f"[bench] 4k-int: {r.get('value')} fps resident")
        sharded = guarded("sharded_4k", lambda: measure_sharded_4k(cx, p4))
        if rank == 0:
            log(f"[bench] sharded_4k: {sharded}")

    if rank != 0:
        if cx.dist is not None:
            cx.dist.destroy_process_group()
        return 0

    base = None
    if not args.no_cpu_baseline and world == 1:
        def cb(w_):
            try:
                return cpu_baseline(w_)
            except Exception as e:          # the baseline is a report, never a reason to lose the GPU numbers
                return {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        base = cb(wl)
        for nm, rec in workloads.items():
            if rec.get("value"):
                rec["cpu_baseline"] = cb(WORKLOADS[nm])
    for nm, rec in workloads.items():
        if "roofline" in rec:
            rec["config"] = bench_config(WORKLOADS[nm], nm, args.frames_per_step or WORKLOADS[nm]["frames_per_step"], world, args.contexts)
            rec["dtype"] = wl_dtype(WORKLOADS[nm])
            rec["unit"] = "frames/s"

    out = {"metric": "vmaf_frames_per_sec", "value": head["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": wl_dtype(wl), "data": "synthetic",
           "config": bench_config(wl, wname, fps_step, world, args.contexts), "clocks": head["clocks"], "e2e": head["e2e"],
           "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "cpu_baseline": base,
           "kernels": head["kernels"],
           "published_reference_fps": "23-26 fps (1080p, libvmaf n_threads=4, decode included; BASELINE.md §1)"}
    if workloads:
        out["workloads"] = workloads
    if sharded is not None:
        out["sharded_4k"] = sharded
        out["e2e_sharded"] = {k: sharded.get(k) for k in ("value", "unit", "frames", "n_gpus", "scaling", "seconds",
                                                          "pooled_vmaf_mean", "identical_to_n1", "resident_value")}
        out["e2e_sharded"]["see"] = "sharded_4k (the full record)"
    if batch is not None:
        out["batch"] = batch
    extra = {}
    if file_e2e is not None:
        extra["file_e2e"] = file_e2e
    if twogpu is not None:
        extra["two_gpus_in_one_process"] = twogpu
    if extra:
        out["extra"] = extra
    print(json.dumps(out), flush=True)
    if cx.dist is not None:
        cx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
