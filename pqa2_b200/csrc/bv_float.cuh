// Internal interface of the float extractors (float VIF / ADM / motion, float_ssim, float_ms_ssim)
// used by the vmaf_float_* models (libvmaf float_vif.c, float_adm.c, float_motion.c, float_ssim.c,
// float_ms_ssim.c; reached from the reference at app/vmaf_analyzer.py:417).
#pragma once
#include "bv_common.cuh"
#include "../../include/b200vmaf.h"

#define BV_FRAW_WORDS 64
// fraw[] layout (doubles, per frame)
#define BV_FRAW_MOTION_SAD 0          // sum |blur_i - blur_{i-1}| (float motion)
#define BV_FRAW_VIF 1                 // 4 scales x (num, den)
#define BV_FRAW_ADM 9                 // 4 scales x (num h,v,d cube sums ; den h,v,d cube sums) = 24
#define BV_FRAW_SSIM 33               // sum of the ssim map, count
#define BV_FRAW_MS_SSIM 35            // 5 scales x (l, c, s map sums) = 15 -> 50
#define BV_FRAW_FFSSIM 50             // FFmpeg ssim filter: sum over 8x8 windows, planes Y, U, V

struct BvFloatState;
BvFloatState *bv_float_create(int w, int h, int bpc, unsigned feat, int batch, const bv_opts *opts);
void bv_float_destroy(BvFloatState *);
void bv_float_reset(BvFloatState *);            // forget the motion state (next clip)
void bv_float_launch(BvFloatState *, const BvBatch &b, BvPlane ref_y, BvPlane dis_y, double *d_fraw,
                     const BvLaunch &L);
const char *bv_float_kernel_name(int id);     // ids >= BVK_F_FIRST
// Host finalisation of one frame; returns the BV_FEAT_* bits it filled.
unsigned bv_float_finish(BvFloatState *, const double *fraw, unsigned frame_flags, bv_frame_features *o);
