"""Summarise an `ncu --page raw --csv` dump: one line per kernel with the metrics DESIGN.md argues from."""
import csv
import re
import sys

KEYS = [("us", "gpu__time_duration.sum"), ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("st_long", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("st_short", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        ("st_bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("st_notsel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
        ("bankconf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel | grid | " + " | ".join(k for k, _ in KEYS))
    for d in data:
        nm = re.sub(r"\(.*", "", d[idx["Kernel Name"]]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
        vals = []
        for k, m in KEYS:
            v = d[idx[m]] if m in idx else ""
            try:
                x = float(v.replace(",", ""))
                u = units[idx[m]]
                if k == "us" and u == "ms":
                    x *= 1000
                if k.endswith("_MB") and u == "Kbyte":
                    x /= 1000
                if k.endswith("_MB") and u == "Gbyte":
                    x *= 1000
                if k.endswith("_MB") and u == "byte":
                    x /= 1e6
                v = ("%.0f" % x) if k == "bankconf" or k == "regs" else ("%.2f" % x)
            except ValueError:
                pass
            vals.append(v)
        print(nm + " | " + d[idx["Grid Size"]] + " | " + " | ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
