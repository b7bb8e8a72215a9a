"""profiles/r01_traffic.json + per-kernel ncu summaries from `ncu --page raw --csv` dumps of one frame group.

The capture is a window of consecutive launches of a bench run; kernels are named by their position in the
launch order of a frame group (several scales share one template instantiation), anchored on the first kernel
of a group."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = {
    # libvmaf psnr=1 is psnr_y; the motion blur rides in f_vif_subsample_s1 and float_ssim's box decimation in ms_ssim_lpf_s1
    # (fused stagings); ssim_maps follows the MS-SSIM chain
    "1080p-float": ["psnr_sse_y", "f_vif_stat_s0", "f_vif_subsample_s1", "f_motion_sad",
                    "f_vif_stat_s1", "f_vif_subsample_s2", "f_vif_stat_s2", "f_vif_subsample_s3", "f_vif_stat_s3", "f_adm_scale0",
                    "f_adm_scale1", "f_adm_scale2", "f_adm_scale3", "ms_ssim_maps_s0",
                    "ms_ssim_lpf_s1", "ms_ssim_maps_s1", "ms_ssim_lpf_s2", "ms_ssim_maps_s2", "ms_ssim_lpf_s3",
                    "ms_ssim_maps_s3", "ms_ssim_lpf_s4", "ms_ssim_maps_s4", "ssim_maps", "f_reduce"],
    # round 2: pyramid level 1 and the motion blur are produced by vif_stat_s0 (fused); the SAD follows the VIF chain
    "1080p-int": ["vif_stat_s0", "vif_stat_s1", "vif_subsample_s2", "vif_stat_s2", "vif_subsample_s3", "vif_stat_s3",
                  "motion_sad", "adm_scale0", "adm_scale1", "adm_scale2", "adm_scale3", "adm_rows_finish"],
}
ORDER["4k-int"] = ORDER["1080p-int"]
# a kernel that occurs once per frame group, and its name in ORDER
ANCHOR = {"1080p-float": ("f_motion_sad_kernel", "f_motion_sad"), "1080p-int": ("motion_sad_kernel", "motion_sad"),
          "4k-int": ("motion_sad_kernel", "motion_sad")}
FRAMES = {"1080p-float": 32, "1080p-int": 32, "4k-int": 16}
ROUND = "r02"
KEYS = [("us", "gpu__time_duration.sum"), ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("alu_pipe_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        ("fp64_pipe_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("warp_inst", "smsp__inst_executed.sum"),
        ("regs", "launch__registers_per_thread"),
        ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("stall_not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio")]
UNIT = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def load(path, wname, frames):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [d[ix["Kernel Name"]] for d in data]
    anchor, anchor_name = ANCHOR[wname]
    a = next(i for i, n in enumerate(names) if anchor in n and "f_" + anchor not in n)
    order = ORDER[wname]
    out = {}
    for i, d in enumerate(data):
        nm = order[(i - a + order.index(anchor_name)) % len(order)]
        if nm in out:
            continue
        e = {"ncu_kernel": re.sub(r"\(.*", "", names[i]).replace("void <unnamed>::", "").replace("<unnamed>::", ""),
             "grid": d[ix["Grid Size"]], "frames": frames}
        for k, m in KEYS:
            if m not in ix:
                continue
            try:
                v = float(d[ix[m]].replace(",", "")) * UNIT.get(units[ix[m]], 1.0)
            except ValueError:
                continue
            e[k] = round(v, 3)
        e["bytes"] = int(round((e.get("dram_read_MB", 0) + e.get("dram_write_MB", 0)) * 1e6))
        e["source"] = f"ncu --set full --clock-control none, {os.path.basename(path)} (profiles/{ROUND}_ncu_full_{wname}_{VER}.md)"
        e["pipe_note"] = (f"issue slots {e.get('issue_pct')} % active, FMA pipe {e.get('fma_pipe_pct')} %, ALU pipe "
                          f"{e.get('alu_pipe_pct')} %, FP64 pipe {e.get('fp64_pipe_pct')} %: bound by instruction issue, not by HBM")
        out[nm] = e
    return out


VER = sys.argv[1] if len(sys.argv) > 1 else "v3"


def main():
    traffic = {}
    old = os.path.join(ROOT, "profiles", f"{ROUND}_traffic.json")
    if os.path.exists(old):
        traffic = json.load(open(old))
    for wname, path in (("1080p-float", f"gpurun_out/raw_float_{VER}.csv"), ("1080p-int", f"gpurun_out/raw_int_{VER}.csv"),
                        ("4k-int", f"gpurun_out/raw_4k_{VER}.csv")):
        p = os.path.join(ROOT, path)
        if not os.path.exists(p):
            continue
        frames = FRAMES[wname]
        t = load(p, wname, frames)
        traffic[wname] = t
        cols = ["kernel"] + [k for k, _ in KEYS]
        lines = [f"# ncu --set full --clock-control none, one frame group ({frames} frame pairs per launch) of "
                 f"`python bench.py --workload {wname} --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras`", "",
                 " | ".join(cols), " | ".join("---" for _ in cols)]
        for nm in ORDER[wname]:
            if nm in t:
                lines.append(" | ".join([nm] + [str(t[nm].get(k, "")) for k, _ in KEYS]))
        open(os.path.join(ROOT, "profiles", f"{ROUND}_ncu_full_{wname}_{VER}.md"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{ROUND}_traffic.json"), "w"), indent=1)
    print({w: len(t) for w, t in traffic.items()})


if __name__ == "__main__":
    main()
