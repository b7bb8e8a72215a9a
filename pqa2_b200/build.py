"""Builds ``pqa2_b200/libb200vmaf.so`` (the C-ABI CUDA library, sm_100a only) in-tree with nvcc.

The built library travels to the GPU box with the repo snapshot (it is git-ignored, not
gpurun-ignored).  No torch extension machinery: the boundary is plain C (include/b200vmaf.h)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200vmaf.so")

# --fmad=false: the bit-exact integer extractors contain a few IEEE float/double steps (VIF gain,
# ADM angle test / gain limit) that must evaluate exactly as the C oracle does.  The float
# extractors follow libvmaf's C float arithmetic operation for operation as well (a fused multiply-add
# in the VIF variance / ADM threshold terms moves VMAF by ~1e-4, and float_ssim by ~1e-6).
EXACT = ["bv_motion.cu", "bv_vif.cu", "bv_adm.cu", "bv_misc.cu", "bv_api.cu", "bv_model.cu", "bv_float.cu"]
FAST = []
# bv_float.cu a second time with -DBV_FAST_FLOAT (contracted multiply-add, folded symmetric taps): bv_opts.fast_float
VARIANTS = {"bv_float_fast.o": ("bv_float.cu", ["-DBV_FAST_FLOAT"])}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# experiments: extra nvcc flags for every source, e.g. BV_EXTRA_NVCC_FLAGS="-DBV_VIF_STAT_FLAT" python -m pqa2_b200.build --force
EXTRA = os.environ.get("BV_EXTRA_NVCC_FLAGS", "").split()
COMMON = EXTRA + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
          "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200vmaf.so cannot be built (there is no CPU fallback)")


def _stale(out: str, deps: list[str]) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "b200vmaf.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    for name in EXACT + FAST:
        src = os.path.join(CSRC, name)
        obj = os.path.join(BUILD, name.replace(".cu", ".o"))
        if force or _stale(obj, [src] + headers):
            flags = list(COMMON) + (["--fmad=false"] if name in EXACT else [])
            if verbose:
                flags += ["-Xptxas", "-v"]
            jobs.append([nvcc, *ARCH, *flags, "-c", src, "-o", obj])

    for oname, (sname, extra) in VARIANTS.items():
        src = os.path.join(CSRC, sname)
        obj = os.path.join(BUILD, oname)
        if force or _stale(obj, [src] + headers):
            flags = list(COMMON) + ["--fmad=false"] + extra
            if verbose:
                flags += ["-Xptxas", "-v"]
            jobs.append([nvcc, *ARCH, *flags, "-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                sys.stderr.write(out)
    objs = [os.path.join(BUILD, n.replace(".cu", ".o")) for n in EXACT + FAST] + [os.path.join(BUILD, n) for n in VARIANTS]
    if force or jobs or _stale(LIB, objs):
        # the shared CUDA runtime (found through the rpath, or already loaded by torch under the same soname): the
        # static one would embed its whole entry-point table in the shipped library
        cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
        run([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "shared", "-Xlinker", f"-rpath={cuda_lib}",
             "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-lpthread", "-ldl", "-lrt"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
