"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck): integer + float extractors, psnr, ffssim,
bookend scan, two launch groups with a lead-in frame.  Usage: compute-sanitizer --tool racecheck python tools/sanitize.py"""
import sys
sys.path.insert(0, '.')
from pqa2_b200 import _lib as L, synth, bookend
from pqa2_b200.extractor import FeatureExtractor

w, h = 352, 288
feats = (L.FEAT_VMAF_INT | L.FEAT_VMAF_FLOAT | L.FEAT_PSNR_Y | L.FEAT_PSNR_UV | L.FEAT_FFSSIM | L.FEAT_FLOAT_SSIM |
         L.FEAT_FLOAT_MS_SSIM)
for bpc in (8, 10):
    frames = [synth.frame_pair(3, f, w, h, bpc) for f in range(5)]
    with FeatureExtractor(w, h, bpc, 420, feats, batch_frames=3) as fx:
        for f, (a, b) in enumerate(frames):
            fx.submit(f, a, b, L.FRAME_FIRST if f == 0 else 0)
        out = fx.fetch()
    print(bpc, out[4].adm2, out[4].f_adm2, out[4].float_ms_ssim, out[4].ffssim[0])
# round 2: the fused scale-0 kernel (pyramid level 1 + motion blur from the VIF tile) on a picture whose last tiles
# overhang both edges (redirected blur taps), with a lead-in frame and n_subsample-skipped frames (blur-only tiles)
w2, h2 = 333, 251
frames = [synth.frame_pair(5, f, w2, h2, 8, chroma=False) for f in range(5)]
with FeatureExtractor(w2, h2, 8, 0, L.FEAT_VMAF_INT, batch_frames=4) as fx:
    for f, (a, b) in enumerate(frames):
        fx.submit(f, a, b, (L.FRAME_FIRST | L.FRAME_LEAD_IN) if f == 0 else (L.FRAME_SKIP_SPATIAL if f == 2 else 0))
    out = fx.fetch()
print("fused", out[4].adm2, out[4].vif_scale[1], out[2].motion, out[4].motion)
print(bookend.luma_stats([synth.ref_luma(1, f, w, h) for f in range(3)])[0])
