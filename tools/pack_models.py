"""Converts the reference's libvmaf model files (/root/reference/models/*.json) into the packed,
dense form shipped under pqa2_b200/models/ (the GPU box has no /root/reference).

Run in the build container:  python tools/pack_models.py [/root/reference/models]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pqa2_b200 import model as M  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/models"
for f in sorted(os.listdir(src)):
    if not f.endswith(".json"):
        continue
    m = M.load_model_file(os.path.join(src, f))
    out = os.path.join(M.MODELS_DIR, f[:-5] + ".bvm.json")
    with open(out, "w") as fh:
        json.dump(M.pack(m), fh, separators=(",", ":"))
    print(out, len(m.main.coef), "SV", len(m.bootstrap), "bootstrap models")
