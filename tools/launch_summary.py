"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): per kernel launches, mean time, share."""
import csv, sys, collections
path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path)) if r]
hi = next(i for i, r in enumerate(rows) if r[0] == "ID")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    key = (r[ix["Kernel Name"]], r[ix["Grid Size"]], r[ix["Block Size"]])
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# ncu launch list summary -- {title}")
print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)")
print("kernel | launches | avg_us | share_of_captured_time | grid | block")
for (nm, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{nm} | {n} | {t / n:.1f} | {t / tot:.3f} | {g} | {b}")
