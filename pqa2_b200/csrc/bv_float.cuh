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

// bv_float.cu is compiled twice into the library: the faithful build (libvmaf's C arithmetic operation for operation)
// and, with -DBV_FAST_FLOAT, the contracted / folded-tap build selected by bv_opts.fast_float.  The second build
// renames its state type and entry points so that both sets link side by side.
#ifdef BV_FAST_FLOAT
#define BvFloatState BvFloatStateFast
#define bv_float_create bv_float_fast_create
#define bv_float_destroy bv_float_fast_destroy
#define bv_float_reset bv_float_fast_reset
#define bv_float_launch bv_float_fast_launch
#define bv_float_kernel_name bv_float_fast_kernel_name
#define bv_float_finish bv_float_fast_finish
#endif
struct BvFloatState;
BvFloatState *bv_float_create(int w, int h, int bpc, unsigned feat, int batch, const bv_opts *opts);
void bv_float_destroy(BvFloatState *);
void bv_float_reset(BvFloatState *);            // forget the motion state (next clip)
void bv_float_launch(BvFloatState *, const BvBatch &b, BvPlane ref_y, BvPlane dis_y, double *d_fraw,
                     const BvLaunch &L);
const char *bv_float_kernel_name(int id);     // ids >= BVK_F_FIRST
// Host finalisation of one frame; returns the BV_FEAT_* bits it filled.
unsigned bv_float_finish(BvFloatState *, const double *fraw, unsigned frame_flags, bv_frame_features *o);
#ifndef BV_FAST_FLOAT
// the fast build's entry points, as bv_api.cu sees them
struct BvFloatStateFast;
BvFloatStateFast *bv_float_fast_create(int w, int h, int bpc, unsigned feat, int batch, const bv_opts *opts);
void bv_float_fast_destroy(BvFloatStateFast *);
void bv_float_fast_reset(BvFloatStateFast *);
void bv_float_fast_launch(BvFloatStateFast *, const BvBatch &b, BvPlane ref_y, BvPlane dis_y, double *d_fraw,
                          const BvLaunch &L);
unsigned bv_float_fast_finish(BvFloatStateFast *, const double *fraw, unsigned frame_flags, bv_frame_features *o);
#endif
