import sys, numpy as np
sys.path.insert(0, '.')
import oracle
from pqa2_b200 import _lib as L, synth
from pqa2_b200.extractor import FeatureExtractor
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 360)
frames = [synth.frame_pair(21, f, w, h, 8, chroma=False) for f in range(2)]
FE = L.FEAT_VMAF_FLOAT | L.FEAT_FLOAT_SSIM | L.FEAT_FLOAT_MS_SSIM
with FeatureExtractor(w, h, 8, 0, FE) as fx:
    for f, (rp, dp) in enumerate(frames):
        fx.submit(f, rp, dp, L.FRAME_FIRST if f == 0 else 0)
    out = fx.fetch()
prev = None
for f, (rp, dp) in enumerate(frames):
    r = oracle.float_features(rp[0], dp[0], 8, prev_ref=prev, ssim=True, ms_ssim=True)
    prev = rp[0]
    o = out[f]
    print("frame", f)
    for s in range(4):
        print(" vif s%d num %.6f vs %.6f (rel %.2e) den %.6f vs %.6f (rel %.2e)" % (s, o.f_vif_num[s], r["vif"]["num"][s],
              o.f_vif_num[s] / r["vif"]["num"][s] - 1, o.f_vif_den[s], r["vif"]["den"][s], o.f_vif_den[s] / r["vif"]["den"][s] - 1))
    for s in range(4):
        print(" adm s%d num %.7f vs %.7f (rel %.2e) den %.7f vs %.7f (rel %.2e)" % (s, o.f_adm_num[s], r["adm"]["num_scale"][s],
              o.f_adm_num[s] / r["adm"]["num_scale"][s] - 1, o.f_adm_den[s], r["adm"]["den_scale"][s], o.f_adm_den[s] / r["adm"]["den_scale"][s] - 1))
        print("     oracle cube sums num", r["adm"]["num_sum"][s], "den", r["adm"]["den_sum"][s])
    print(" adm2 %.9f vs %.9f" % (o.f_adm2, r["adm2"]), " motion %.9f vs %.9f" % (o.f_motion, r["motion"]))
    print(" ssim %.9f vs %.9f  ms_ssim %.9f vs %.9f" % (o.float_ssim, r["float_ssim"], o.float_ms_ssim, r["float_ms_ssim"]))
