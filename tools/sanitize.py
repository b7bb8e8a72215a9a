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
print(bookend.luma_stats([synth.ref_luma(1, f, w, h) for f in range(3)])[0])
