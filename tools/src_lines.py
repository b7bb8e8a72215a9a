"""Aggregate an `ncu --page source --csv` dump per CUDA source line: executed warp instructions and stall samples."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] in ("Line No", "#", "Address") or (r and "Instructions Executed" in r))
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
print("columns:", hdr[:12], file=sys.stderr)
src_col = [i for i, h in enumerate(hdr) if h == "Source"]
ie = idx["Instructions Executed"]; ss = idx.get("# Samples")
agg = collections.defaultdict(lambda: [0, 0, 0]); text = {}
total = 0; tot_s = 0
cur_line = None
cur_file = ""
for r in rows[:hi]:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]
for r in rows[hi + 1:]:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) <= ie: continue
    try: n = int(r[ie] or 0)
    except ValueError: continue
    s = int(r[ss]) if ss is not None and r[ss].isdigit() else 0
    if not r[0]: continue
    key = cur_file + ':' + r[0]
    agg[key][0] += n; agg[key][1] += s; agg[key][2] += 1
    text.setdefault(key, r[src_col[0]] if src_col else "")
    total += n; tot_s += s
print(f"total warp-instr {total}  samples {tot_s}")
for k, (n, s, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k:>22} {n:>12} {100*n/total:5.1f}% samp {100*s/max(tot_s,1):5.1f}% sass {c:4d} | {text[k][:110]}")
