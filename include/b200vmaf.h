/*
 * b200vmaf.h -- C ABI of libb200vmaf.so, the B200-native (sm_100a) VMAF feature engine.
 *
 * This is the drop-in boundary for the ONE hot path of yoseph007/PQA2: today
 * app/vmaf_analyzer.py:411-419 spawns `ffmpeg -lavfi libvmaf=...` (and :1027-1034 /
 * :1057-1064 the `psnr` / `ssim` filters) and reads a JSON log back (:640-641).  The
 * reference has no FFI for this path (it is a process + file boundary), so the entry points
 * below are what a ctypes binding in app/vmaf_analyzer.py would call instead of Popen; each
 * one cites the reference behaviour it replaces.  See INTEGRATION.md for the binding stub.
 *
 * Conventions: plain C, opaque handles, int return (0 = ok, <0 = error; text through
 * bv_last_error), no exceptions across the boundary, no torch / Python types.
 * Threading: one bv_ctx per GPU, driven by one host thread at a time; bv_cancel() may be
 * called from any thread (maps VMAFAnalyzer.terminate_analysis, app/vmaf_analyzer.py:139-151).
 */
#ifndef B200VMAF_H
#define B200VMAF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BV_ABI_VERSION 2   /* 2: bv_host_register / bv_host_unregister, bv_opts.fast_float (took one reserved word) */

/* feature groups (bv_create features_mask) */
#define BV_FEAT_MOTION      0x001u  /* integer_motion / integer_motion2           (libvmaf integer_motion.c) */
#define BV_FEAT_VIF         0x002u  /* integer_vif_scale0..3                      (libvmaf integer_vif.c)    */
#define BV_FEAT_ADM         0x004u  /* integer_adm2, integer_adm_scale0..3        (libvmaf integer_adm.c)    */
#define BV_FEAT_PSNR_Y      0x008u  /* psnr_y  (libvmaf `psnr=1`, reference app/vmaf_analyzer.py:385)        */
#define BV_FEAT_PSNR_UV     0x010u  /* chroma SSE for the FFmpeg `psnr` stats file (app/vmaf_analyzer.py:1032) */
#define BV_FEAT_FFSSIM      0x020u  /* FFmpeg `ssim` filter, all planes           (app/vmaf_analyzer.py:1062) */
#define BV_FEAT_FLOAT_VIF   0x040u  /* vif_scale0..3   for vmaf_float_* models    (libvmaf float_vif)        */
#define BV_FEAT_FLOAT_ADM   0x080u  /* adm2, adm_scale0..3                        (libvmaf float_adm)        */
#define BV_FEAT_FLOAT_MOTION 0x100u /* motion, motion2                            (libvmaf float_motion)     */
#define BV_FEAT_FLOAT_SSIM  0x200u  /* float_ssim      (libvmaf `ssim=1`, app/vmaf_analyzer.py:386)          */
#define BV_FEAT_FLOAT_MS_SSIM 0x400u/* float_ms_ssim                                                         */
#define BV_FEAT_VMAF_INT    (BV_FEAT_MOTION | BV_FEAT_VIF | BV_FEAT_ADM)
#define BV_FEAT_VMAF_FLOAT  (BV_FEAT_FLOAT_MOTION | BV_FEAT_FLOAT_VIF | BV_FEAT_FLOAT_ADM)

/* per-frame flags (bv_submit) */
#define BV_FRAME_LEAD_IN    0x1     /* blur only, no scores: the one-frame overlap of a frame shard          */
#define BV_FRAME_SKIP_SPATIAL 0x2   /* n_subsample: motion only (temporal extractors always run)             */
#define BV_FRAME_FIRST      0x4     /* no previous frame: motion = 0 (libvmaf index == 0)                    */

typedef struct bv_ctx bv_ctx;
typedef struct bv_model bv_model;

/* Feature options.  Defaults = libvmaf defaults; NEG models set both gain limits to 1.0
 * (reference models/vmaf_v0.6.1neg.json:34-51). */
typedef struct bv_opts {
    double vif_enhn_gain_limit;      /* 100.0 */
    double adm_enhn_gain_limit;      /* 100.0 */
    double adm_norm_view_dist;       /* 3.0   */
    int    adm_ref_display_height;   /* 1080  */
    int    batch_frames;             /* frame pairs per kernel launch group (0 = auto) */
    int    fast_float;               /* float extractors only, opt-in: contract multiply-add and fold the symmetric filter
                                        taps (half the FP32 work).  Results leave libvmaf's scalar operation order but stay
                                        inside the float models' tolerance (measured: profiles/r02_fast_float.md).  0 =
                                        faithful order (default). */
    int    reserved[5];
} bv_opts;

#define BV_RAW_WORDS 64
/* Raw per-frame integer accumulators exactly as the kernels produced them (bit-exact parity
 * is asserted on these), plus the derived feature values libvmaf would log. */
typedef struct bv_frame_features {
    int64_t  frame_index;
    uint32_t flags;
    uint32_t valid_mask;             /* BV_FEAT_* actually computed for this frame */
    int64_t  raw[BV_RAW_WORDS];      /* layout: BV_RAW_* below */
    /* derived (host-side scalar finalisation of the raw accumulators) */
    double motion;                   /* integer_motion (motion2 is a host min-filter over frames) */
    double vif_num[4], vif_den[4], vif_scale[4];
    double adm_num[4], adm_den[4], adm_scale[4], adm2;
    double psnr_y, psnr_cb, psnr_cr;
    double ffssim[3];                /* FFmpeg ssim filter Y, U, V */
    /* float extractors (vmaf_float_* models) */
    double f_motion;
    double f_vif_num[4], f_vif_den[4], f_vif_scale[4];
    double f_adm_num[4], f_adm_den[4], f_adm_scale[4], f_adm2;
    double float_ssim, float_ms_ssim;
} bv_frame_features;

/* raw[] layout */
#define BV_RAW_SAD       0           /* u64: sum |blur_i - blur_{i-1}|                     */
#define BV_RAW_VIF       1           /* 4 scales x 7: num_log, den_log, num_non_log,
                                        den_non_log, accum_x, accum_x2, num_accum_x        */
#define BV_RAW_ADM_CM    29          /* 4 scales x 3 (h, v, d): contrast-masked numerator  */
#define BV_RAW_ADM_DEN   41          /* 4 scales x 3 (h, v, d): csf denominator            */
#define BV_RAW_SSE       53          /* 3 planes                                           */

int  bv_abi_version(void);
int  bv_device_count(void);

/* chroma: 420 / 422 / 444 (0 = luma only: chroma planes are never read). */
bv_ctx *bv_create(int device, int width, int height, int bpc, int chroma, unsigned features_mask,
                  const bv_opts *opts);
void bv_destroy(bv_ctx *);
const char *bv_last_error(bv_ctx *);     /* ctx may be NULL: error of the last failed bv_create */

/* cudaHostAlloc / cudaFreeHost wrappers so callers can stage frames in pinned memory. */
int  bv_pinned_alloc(void **p, size_t n);
int  bv_pinned_free(void *p);
/* cudaHostRegister / cudaHostUnregister of caller-owned memory -- e.g. an mmap of a raw .y4m / .yuv clip, so that
 * bv_submit's cudaMemcpyAsync reads the page cache directly instead of a staging copy (the reference lets ffmpeg
 * read the two input files itself, app/vmaf_analyzer.py:415-416).  read_only != 0 asks for a read-only registration
 * (PROT_READ mappings).  Returns BV_ERR_CUDA when the platform refuses; callers then fall back to a pinned ring. */
int  bv_host_register(void *p, size_t n, int read_only);
int  bv_host_unregister(void *p);
/* cudaMalloc / cudaFree / blocking H2D copy on `device`, for callers that keep clips resident in HBM
 * and score them with bv_submit_device (the resident mode of bench.py). */
int  bv_device_alloc(int device, void **p, size_t n);
int  bv_device_free(int device, void *p);
int  bv_device_upload(int device, void *dst, const void *src, size_t n);
size_t bv_sizeof_frame_features(void);   /* ABI check for bindings */

/* Enqueue one ref/dis frame pair (HOST planes; Y, U, V; U/V may be NULL when chroma == 0 or no
 * chroma feature is enabled).  Copies are cudaMemcpy2DAsync on the upload stream; kernels run
 * on the compute stream once a launch group is full (or on bv_flush).  Host buffers must stay
 * valid until bv_wait_uploads() or bv_flush() returns.  Frames must be submitted in order.
 * Replaces: vmaf_read_pictures() inside the ffmpeg child (app/vmaf_analyzer.py:417). */
int  bv_submit(bv_ctx *, int64_t frame_index, const void *const ref_planes[3], const size_t ref_stride[3],
               const void *const dis_planes[3], const size_t dis_stride[3], unsigned frame_flags);
/* Same, planes already in device memory of this ctx's GPU (resident benchmarking, NVDEC). */
int  bv_submit_device(bv_ctx *, int64_t frame_index, const void *const ref_planes[3], const size_t ref_stride[3],
                      const void *const dis_planes[3], const size_t dis_stride[3], unsigned frame_flags);
int  bv_wait_uploads(bv_ctx *);          /* all submitted host buffers may be reused after this */
int  bv_flush(bv_ctx *);                 /* launch the partial group and drain every stream     */
int  bv_kick(bv_ctx *);                  /* launch the partial group now, without draining: lets a caller start the
                                            GPU on the first few frames of a clip instead of waiting for a full group */
int64_t bv_frames_done(bv_ctx *);        /* frames whose features are ready (non-blocking)      */
int  bv_batch_frames(bv_ctx *);          /* frame pairs per launch group of this ctx (opts.batch_frames, or the auto
                                            choice: 32 up to 1440p, 16 at 2160p)                                */
/* Copy out features of `count` frames starting at submission ordinal `first` (0-based, in
 * submission order incl. lead-in frames); blocks until they are ready. */
int  bv_fetch(bv_ctx *, int64_t first, int64_t count, bv_frame_features *out);
int  bv_cancel(bv_ctx *);                /* thread-safe; subsequent calls fail with BV_ERR_CANCELLED */
/* Drain, drop the stored results and the motion state, clear a cancel: the ctx is ready for the next clip of the
 * same geometry (the reference pays a process start per clip, app/vmaf_analyzer.py:446; a batch of clips --
 * BASELINE.json configs[4] -- reuses one ctx per GPU). */
int  bv_reset(bv_ctx *);
int64_t bv_kernel_launches(bv_ctx *);    /* kernels launched by this ctx so far (bench `gpu_launches`) */
/* Per-kernel CUDA-event profiling (events on the compute stream around every launch).  Enable with
 * bv_set_profiling(ctx, 1); ids run 0 .. bv_kernel_slots()-1, bv_kernel_name(id) is NULL for unused
 * ids.  bv_kernel_ms = accumulated device time, bv_kernel_count = launches timed. */
int  bv_set_profiling(bv_ctx *, int enable);
int  bv_kernel_slots(void);
const char *bv_kernel_name(int id);
double bv_kernel_ms(bv_ctx *, int id, int reset);
double bv_kernel_count(bv_ctx *, int id, int reset);
/* Device stopwatch on the compute stream: mark(0) before the first submit of a timed region, mark(1)
 * after the last (launches any partial group first); elapsed blocks until mark(1) has happened. */
int  bv_timer_mark(bv_ctx *, int which);
double bv_timer_elapsed_ms(bv_ctx *);

/* ---- bookend (white-frame) scan: the per-pixel step of the alignment stage that precedes the path ----
 * Replaces the cv2 loop at app/bookend_alignment.py:997-1020 (mean / std / share of pixels above a threshold per
 * frame) and app/reference_analyzer.py:131-141.  Frames are resident luma planes on `device`, frame f at
 * d_luma + f * frame_stride.  out_host: [n_frames][5] = sum y, sum y^2, count(y > thr[0]), (> thr[1]), (> thr[2]).
 * Blocking (one launch + one small D2H). */
int  bv_luma_stats_device(int device, const void *d_luma, size_t pitch, size_t frame_stride, int n_frames,
                          int w, int h, int bpc, const unsigned thr[3], unsigned long long *out_host);

#define BV_ERR_ARG        -1
#define BV_ERR_CUDA       -2
#define BV_ERR_CANCELLED  -3
#define BV_ERR_ORDER      -4
#define BV_ERR_UNSUPPORTED -5

/* ---- SVR fusion (libvmaf predict.c + libsvm svm_predict; reached at app/vmaf_analyzer.py:417) ----
 * sv: dense [n_sv][n_feat] (missing sparse indices = 0); slopes/intercepts: n_feat + 1, index 0 = score.
 * transform: p0,p1,p2 + flags (bit0 p0, bit1 p1, bit2 p2 present; bit3 out_lte_in, bit4 out_gte_in). */
#define BV_MODEL_ENABLE_TRANSFORM 0x1
#define BV_MODEL_DISABLE_CLIP     0x2
bv_model *bv_model_create(int n_feat, int n_sv, const double *sv, const double *coef, double gamma, double rho,
                          const double *slopes, const double *intercepts, const double clip[2], int has_clip,
                          const double transform_p[3], unsigned transform_flags);
void bv_model_free(bv_model *);
/* feat: [n][n_feat] in the model's feature order; out: [n] final scores. */
int  bv_predict(const bv_model *, const double *feat, int64_t n, unsigned flags, double *out);
/* same, evaluated by the svr_predict CUDA kernel on `device` (one thread per (frame, SV) pair).  The model keeps, per
 * device, a mirror of its vectors plus a stream and grow-only scratch of its own (freed by bv_model_free): a call neither
 * allocates nor frees device memory, so it never waits for extractor kernels still running on that device.  Calls on one
 * model are serialised; any thread may call. */
int  bv_predict_device(const bv_model *, int device, const double *feat, int64_t n, unsigned flags, double *out);

#ifdef __cplusplus
}
#endif
#endif /* B200VMAF_H */
