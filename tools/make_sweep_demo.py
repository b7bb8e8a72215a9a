import sys, os
sys.path.insert(0, '.')
from pqa2_b200 import synth, yuvio
d = sys.argv[1]; os.makedirs(d, exist_ok=True)
rows = ["name,reference,distorted"]
for k in range(3):
    fr = [synth.frame_pair(70 + k, f, 320, 180, 8) for f in range(6)]
    yuvio.write_y4m(f"{d}/ref{k}.y4m", [list(a) for a, _ in fr], 320, 180)
    yuvio.write_y4m(f"{d}/dis{k}.y4m", [list(b) for _, b in fr], 320, 180)
    rows.append(f"clip{k},ref{k}.y4m,dis{k}.y4m")
open(f"{d}/pairs.csv", "w").write("\n".join(rows) + "\n")
