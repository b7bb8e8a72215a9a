/*
 * libvmaf_spec.h -- every numeric constant of the restated libvmaf / FFmpeg arithmetic, in ONE place.
 *
 * The arithmetic this engine replaces lives outside the reference repository (Netflix libvmaf reached through
 * `ffmpeg -lavfi libvmaf`, app/vmaf_analyzer.py:411-419), so every table below is a
 * restatement from the published sources.  The CPU oracle (oracle/ *.c, test infrastructure) and the CUDA kernels
 * (pqa2_b200/csrc/ *.cu, the product) both take their constants from here: a recollection error is fixed once, and
 * a difference between the two can only be a difference of CODE, never of tables.  Plain C preprocessor lists so that
 * they initialise `static const` arrays in C and `__constant__` arrays in CUDA alike.
 *
 * Confidence tags follow SURVEY.md Appendix A: (H) high, (M) medium, (L) low -- the L / M items are what a libvmaf log
 * (tests/golden/libvmaf/, tools/make_libvmaf_golden.py) would settle.
 */
#ifndef LIBVMAF_SPEC_H
#define LIBVMAF_SPEC_H

/* ---- integer VIF (integer_vif.c): Q16 Gaussian taps, sigma = N / 5, each table sums to 65536 (H) ---- */
#define SPEC_VIF_Q16_17 489, 935, 1640, 2640, 3896, 5274, 6547, 7455, 7784, 7455, 6547, 5274, 3896, 2640, 1640, 935, 489
#define SPEC_VIF_Q16_9  1244, 3663, 7925, 12590, 14692, 12590, 7925, 3663, 1244
#define SPEC_VIF_Q16_5  3571, 16004, 26386, 16004, 3571
#define SPEC_VIF_Q16_3  10904, 43728, 10904
#define SPEC_VIF_SIGMA_NSQ_Q16 (2 * 65536)          /* sigma_nsq = 2 in Q16 (H) */
#define SPEC_VIF_LOG2_SCALE 2048                    /* log2 table: round(log2f(i) * 2048), i in [32767, 65535] (H) */
#define SPEC_VIF_LOG2_BIAS 17                       /* den_log = sum T / 2048 - sum x - 17 * count (H) */
#define SPEC_VIF_NONLOG_DIV1 16384.0                /* num_non_log / 16384 / 65025 (H) */
#define SPEC_VIF_NONLOG_DIV2 65025.0
#define SPEC_VIF_GAIN_EPS (65536 * 1.0e-10)         /* g = sigma12 / (sigma1_sq + eps) in double (H) */

/* ---- float VIF (vif_options.h): the same Gaussians as float literals (H) ---- */
#define SPEC_VIF_F32_17 0.00745626912f, 0.0142655009f, 0.0250313189f, 0.0402820669f, 0.0594526194f, 0.0804751068f, \
    0.0999041125f, 0.113746084f, 0.118773937f, 0.113746084f, 0.0999041125f, 0.0804751068f, 0.0594526194f, \
    0.0402820669f, 0.0250313189f, 0.0142655009f, 0.00745626912f
#define SPEC_VIF_F32_9  0.0189780835f, 0.0558981746f, 0.120920904f, 0.192116052f, 0.224173605f, 0.192116052f, \
    0.120920904f, 0.0558981746f, 0.0189780835f
#define SPEC_VIF_F32_5  0.054488685f, 0.244201347f, 0.402619958f, 0.244201347f, 0.054488685f
#define SPEC_VIF_F32_3  0.166378498f, 0.667243004f, 0.166378498f
/* vif_tools.c log2f_approx(): degree-8 polynomial of the mantissa, Horner order (H) */
#define SPEC_LOG2_POLY -0.012671635276421f, 0.064841182402670f, -0.157048836463065f, 0.257167726303123f, \
    -0.353800560300520f, 0.480131410397451f, -0.721314327952201f, 1.442694803896991f, 0.0f

/* ---- motion (integer_motion.c / float_motion.c): 5-tap blur (H) ---- */
#define SPEC_MOTION_Q16_5 3571, 16004, 26386, 16004, 3571
#define SPEC_MOTION_F32_5 0.054488685f, 0.244201342f, 0.402619947f, 0.244201342f, 0.054488685f

/* ---- ADM (integer_adm.c / adm_tools.c) ---- */
/* db2 analysis taps: Q15 integers (H) and float (H) */
#define SPEC_DWT_LO_Q15 15826, 27411, 7345, -4240
#define SPEC_DWT_HI_Q15 -4240, -7345, 27411, -15826
#define SPEC_DWT_LO_SUM_Q15 46342
#define SPEC_DWT_LO_F32 0.482962913144690f, 0.836516303737469f, 0.224143868041857f, -0.129409522550921f
#define SPEC_DWT_HI_F32 -0.129409522550921f, -0.224143868041857f, 0.836516303737469f, -0.482962913144690f
/* integer DWT shifts of scales 1..3: vertical {0, 16, 16}, horizontal {15, 16, 15}, half-ulp rounding (L) */
#define SPEC_ADM_DWT_SH_V 0, 16, 16
#define SPEC_ADM_DWT_SH_H 15, 16, 15
/* Watson CSF model of the 9/7 wavelet (H): a, k, f0, g[theta], amplitudes A[scale][theta] */
#define SPEC_DWT79_A 0.495f
#define SPEC_DWT79_K 0.466f
#define SPEC_DWT79_F0 0.401f
#define SPEC_DWT79_G 1.501f, 1.0f, 0.534f, 1.0f
#define SPEC_DWT79_AMP { 0.62171f, 0.67234f, 0.72709f, 0.67234f }, { 0.34537f, 0.41317f, 0.49428f, 0.41317f }, \
    { 0.18004f, 0.22727f, 0.28688f, 0.22727f }, { 0.091401f, 0.11792f, 0.15214f, 0.11792f }
/* scale-0 integer CSF factors for the default 3.0 x 1080 viewing set-up: Q21, Q21, Q23 with shifts 15, 15, 17 (L:
 * literals; within 1e-6 of 1 / Q(0, theta) of the model above but not its rounding, which would give 36452 and 49415 --
 * SURVEY.md Appendix A.4, tests/test_spec_constants.py) */
#define SPEC_ADM_S0_RF 36453, 36453, 49417
#define SPEC_ADM_S0_RF_SHIFT 15, 15, 17
#define SPEC_ADM_S0_RF_ROUND 16384, 16384, 65536
/* csf_f = |csf_a| / 30 and the centre weight |csf_a| / 15: Q12-rounded at scale 0, Q32 at scales 1..3 (M) */
#define SPEC_ADM_ONE_BY_30_Q16 4369
#define SPEC_ADM_ONE_BY_15_Q16 8738
#define SPEC_ADM_ONE_BY_30_Q32 143165577ll
#define SPEC_ADM_ONE_BY_15_Q32 286331153ll
#define SPEC_ADM_BORDER_FACTOR 0.1                  /* centre region: left = (int)(w * 0.1 - 0.5), ... (H) */
#define SPEC_ADM_NUMDEN_LIMIT 1e-10                 /* x (w * h) / (1920 * 1080) (H) */

/* ---- iqa SSIM / MS-SSIM (ssim.c, ms_ssim.c) (H) ---- */
#define SPEC_SSIM_GAUSS11 0.001028f, 0.007599f, 0.036001f, 0.109361f, 0.213006f, 0.266012f, 0.213006f, 0.109361f, \
    0.036001f, 0.007599f, 0.001028f
#define SPEC_MS_SSIM_LPF9 0.026727f, -0.016828f, -0.078201f, 0.266846f, 0.602914f, 0.266846f, -0.078201f, -0.016828f, \
    0.026727f
#define SPEC_MS_SSIM_EXPONENTS 0.0448, 0.2856, 0.3001, 0.2363, 0.1333
#define SPEC_SSIM_K1 0.01f
#define SPEC_SSIM_K2 0.03f

/* ---- FFmpeg vf_ssim.c (8-bit): c1 = (int)(.01^2 * 255^2 * 64 + .5), c2 = (int)(.03^2 * 255^2 * 64 * 63 + .5) (H) ---- */
#define SPEC_FFSSIM_C1 416
#define SPEC_FFSSIM_C2 235963

#endif /* LIBVMAF_SPEC_H */
