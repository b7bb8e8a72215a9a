"""Dynamic SASS opcode histogram from an `ncu --page source --csv` dump (unique addresses, weighted by executed count)."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
px = float(sys.argv[2]) if len(sys.argv) > 2 else 0      # pixels per launch, for thread-instr per pixel
seen = {}
for r in rows:
    if len(r) > 8 and r[0] == "" and r[2].startswith("0x"):
        try: seen[r[2]] = (r[3].strip(), int(r[7] or 0), int(r[8] or 0), int(r[6] or 0))
        except ValueError: pass
tot = sum(v[1] for v in seen.values()); tott = sum(v[2] for v in seen.values()); samp = sum(v[3] for v in seen.values())
h = collections.Counter(); hs = collections.Counter()
for txt, n, tn, s in seen.values():
    op = re.sub(r"^@!?U?P\d+\s+", "", txt).split()[0].split(".")[0]
    if txt.split()[0].startswith("@") is False and re.match(r"^(IMAD)\.(MOV|SHL|IADD)", re.sub(r"^@!?U?P\d+\s+", "", txt)): op = "IMAD.mov/shl/iadd"
    h[op] += n; hs[op] += s
print(f"unique SASS {len(seen)}  warp-instr {tot}  thread-instr {tott}  samples {samp}" + (f"  thread-instr/px {tott/px:.1f} warp-instr*32/px {tot*32/px:.1f}" if px else ""))
for op, n in h.most_common(28):
    print(f"{op:>20} {n:>12} {100*n/tot:5.1f}%  samples {100*hs[op]/max(samp,1):5.1f}%")
