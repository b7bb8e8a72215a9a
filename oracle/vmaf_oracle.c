/*
 * oracle/vmaf_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the arithmetic that the reference application reaches
 * through `ffmpeg -lavfi libvmaf=...` (reference call site: app/vmaf_analyzer.py:373-419)
 * and through the FFmpeg `psnr` / `ssim` filters (app/vmaf_analyzer.py:1027-1034, :1057-1064).
 *
 * PARITY UNPINNED: the arithmetic lives in a third-party dependency that is absent from
 * /root/reference -- Netflix libvmaf (not vendored, not version-pinned; targeted behaviour:
 * libvmaf v3.0.0 == v2.3.x feature arithmetic: src/feature/integer_motion.c, integer_vif.c,
 * integer_adm.c, integer_psnr.c, src/predict.c, src/svm.cpp) and FFmpeg libavfilter
 * (vf_psnr.c, vf_ssim.c).  Neither library nor any golden vector of theirs exists in this
 * image, so this file restates their published algorithms from the description in
 * SURVEY.md Appendix A.  What IS pinned (tests/test_oracle.py, tests/test_cpu_boundary.py): SVR
 * known answers derived from the reference's own models/*.json, identical-pair / static-clip
 * invariants, filter table sums, monotonicity in distortion strength, agreement between the
 * fixed-point and the fp32 restatements, and self-generated regression vectors (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may call into this file.  The product path (pqa2_b200/) never does.
 *
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off: no FMA contraction, so the few
 * float/double steps evaluate exactly as written).
 */
#include "../include/libvmaf_spec.h"
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define MAXI(a, b) ((a) > (b) ? (a) : (b))
#define MINI(a, b) ((a) < (b) ? (a) : (b))

/* ------------------------------------------------------------------------------------------
 * Pixel access.  Pictures are planar; bpc 8 -> uint8 samples, bpc > 8 -> uint16 LE samples
 * (SURVEY.md Appendix A.1).  stride is in BYTES.
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t px(const void *p, int bpc, ptrdiff_t stride, int i, int j)
{
    if (bpc == 8)
        return ((const uint8_t *)p)[(ptrdiff_t)i * stride + j];
    return *(const uint16_t *)((const uint8_t *)p + (ptrdiff_t)i * stride + 2 * (ptrdiff_t)j);
}

/* ==========================================================================================
 * Integer motion (libvmaf integer_motion.c).
 * 5-tap Q16 blur, vertical then horizontal, asymmetric mirror borders
 * (idx < 0 -> -idx ; idx >= n -> 2n - idx - 1); SAD against the previous blurred frame.
 * ======================================================================================== */
static const uint16_t motion_filter[5] = { SPEC_MOTION_Q16_5 };

static inline int mirror_asym(int i, int n)
{
    if (i < 0) return -i;
    if (i >= n) return 2 * n - i - 1;
    return i;
}

ORC_API void orc_motion_blur(const void *src, int bpc, int w, int h, ptrdiff_t stride,
                             uint16_t *dst /* w*h, tight */)
{
    uint16_t *tmp = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    const uint32_t add_v = 1u << (bpc - 1);
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            uint32_t acc = 0;
            for (int k = 0; k < 5; ++k)
                acc += motion_filter[k] * px(src, bpc, stride, mirror_asym(i - 2 + k, h), j);
            tmp[(size_t)i * w + j] = (uint16_t)((acc + add_v) >> bpc);
        }
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            uint32_t acc = 0;
            for (int k = 0; k < 5; ++k)
                acc += motion_filter[k] * (uint32_t)tmp[(size_t)i * w + mirror_asym(j - 2 + k, w)];
            dst[(size_t)i * w + j] = (uint16_t)((acc + 32768u) >> 16);
        }
    free(tmp);
}

ORC_API uint64_t orc_motion_sad(const uint16_t *a, const uint16_t *b, int w, int h)
{
    uint64_t sad = 0;
    for (size_t n = 0; n < (size_t)w * h; ++n)
        sad += (uint64_t)abs((int)a[n] - (int)b[n]);
    return sad;
}

/* libvmaf normalize_and_scale_sad(): (float)(sad / 256.) / (w * h) */
ORC_API double orc_motion_score(uint64_t sad, int w, int h)
{
    return (double)((float)(sad / 256.) / (float)((unsigned)w * (unsigned)h));
}

/* ==========================================================================================
 * Integer VIF (libvmaf integer_vif.c).
 * ======================================================================================== */
static const uint16_t vif_filter[4][17] = { { SPEC_VIF_Q16_17 }, { SPEC_VIF_Q16_9 }, { SPEC_VIF_Q16_5 }, { SPEC_VIF_Q16_3 } };
static const int vif_filter_width[4] = { 17, 9, 5, 3 };

static uint16_t vif_log2_table[65536];
static int vif_log2_ready = 0;

static void vif_log2_init(void)
{
    if (vif_log2_ready) return;
    for (unsigned i = 32767; i < 65536; ++i)
        vif_log2_table[i] = (uint16_t)round(log2f((float)i) * 2048);
    vif_log2_ready = 1;
}

ORC_API const uint16_t *orc_vif_log2_table(void) { vif_log2_init(); return vif_log2_table; }
ORC_API const uint16_t *orc_vif_filter(int scale) { return vif_filter[scale]; }

static inline int reflect101(int i, int n)
{
    if (i < 0) return -i;
    if (i >= n) return 2 * (n - 1) - i;
    return i;
}

static inline uint16_t best16_from32(uint32_t v, int *x)
{
    int k = 16 - __builtin_clz(v);
    v >>= k;
    *x = -k;
    return (uint16_t)v;
}

static inline uint16_t best16_from64(uint64_t v, int *x)
{
    int k = __builtin_clzll(v);
    if (k > 48) {
        k -= 48;
        v <<= k;
        *x = k;
    } else if (k < 47) {
        k = 48 - k;
        v >>= k;
        *x = -k;
    } else {
        *x = 0;
        if (v >> 16) {
            v >>= 1;
            *x = -1;
        }
    }
    return (uint16_t)v;
}

/* Accumulator order in acc[7]:
 *  0 num_log  1 den_log  2 num_non_log  3 den_non_log  4 accum_x  5 accum_x2  6 num_accum_x */
enum { V_NUM_LOG, V_DEN_LOG, V_NUM_NONLOG, V_DEN_NONLOG, V_X, V_X2, V_CNT };

/* The two filter passes run a whole row at a time -- tap loop outside, column loop inside, accumulators in row buffers --
 * so that gcc vectorises them (integer sums: the order of the taps cannot change a bit); each hot loop is compiled for
 * AVX2 and for the baseline ISA, the loader picks one (target_clones).  Same arithmetic as the per-pixel form of
 * libvmaf's integer_vif.c, ~4x the frames/s: this file is also the CPU baseline bench.py reports. */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define ORC_SIMD __attribute__((target_clones("avx2", "default")))
#else
#define ORC_SIMD
#endif

/* a[j] = sum_k f[k] * row_k[j] for the two pictures, and the three second-moment products.  Blocks of ORC_BLK pixels keep
 * the five accumulators in vector registers across the tap loop (a row-long accumulator array would go through L2 once
 * per tap). */
#define ORC_BLK 32
#define ORC_VPASS(T)                                                                            \
    for (int j0 = 0; j0 < w; j0 += ORC_BLK) {                                                   \
        const int nb = w - j0 < ORC_BLK ? w - j0 : ORC_BLK;                                     \
        uint32_t m1[ORC_BLK] = { 0 }, m2[ORC_BLK] = { 0 };                                      \
        uint64_t sxx[ORC_BLK] = { 0 }, syy[ORC_BLK] = { 0 }, sxy[ORC_BLK] = { 0 };              \
        for (int k = 0; k < fw; ++k) {                                                          \
            const T *xr = (const T *)rows_r[k] + j0, *yr = (const T *)rows_d[k] + j0;           \
            const uint32_t fk = f[k];                                                           \
            if (nb == ORC_BLK) {                                                                \
                for (int j = 0; j < ORC_BLK; ++j) {                                             \
                    const uint32_t x = xr[j], y = yr[j];                                        \
                    const uint32_t fx = fk * x, fy = fk * y;                                    \
                    m1[j] += fx; m2[j] += fy;                                                   \
                    sxx[j] += (uint64_t)fx * x; syy[j] += (uint64_t)fy * y; sxy[j] += (uint64_t)fx * y; \
                }                                                                               \
            } else {                                                                            \
                for (int j = 0; j < nb; ++j) {                                                  \
                    const uint32_t x = xr[j], y = yr[j];                                        \
                    const uint32_t fx = fk * x, fy = fk * y;                                    \
                    m1[j] += fx; m2[j] += fy;                                                   \
                    sxx[j] += (uint64_t)fx * x; syy[j] += (uint64_t)fy * y; sxy[j] += (uint64_t)fx * y; \
                }                                                                               \
            }                                                                                   \
        }                                                                                       \
        memcpy(a_mu1 + j0, m1, sizeof(uint32_t) * nb); memcpy(a_mu2 + j0, m2, sizeof(uint32_t) * nb); \
        memcpy(a_xx + j0, sxx, sizeof(uint64_t) * nb); memcpy(a_yy + j0, syy, sizeof(uint64_t) * nb); \
        memcpy(a_xy + j0, sxy, sizeof(uint64_t) * nb);                                          \
    }

ORC_SIMD void orc_vif_vpass(const void *const *rows_r, const void *const *rows_d, int bytes_per_sample, int w, int fw,
                            const uint16_t *f, uint32_t *a_mu1, uint32_t *a_mu2, uint64_t *a_xx, uint64_t *a_yy,
                            uint64_t *a_xy)
{
    if (bytes_per_sample == 1) { ORC_VPASS(uint8_t) } else { ORC_VPASS(uint16_t) }
}

/* horizontal pass over rows that carry their reflect-101 border (r samples on either side): no index arithmetic */
ORC_SIMD void orc_vif_hpass(const uint16_t *p_mu1, const uint16_t *p_mu2, const uint32_t *p_xx, const uint32_t *p_yy,
                            const uint32_t *p_xy, int w, int fw, const uint16_t *f, uint32_t *h_mu1, uint32_t *h_mu2,
                            uint64_t *h_xx, uint64_t *h_yy, uint64_t *h_xy)
{
    for (int j0 = 0; j0 < w; j0 += ORC_BLK) {
        const int nb = w - j0 < ORC_BLK ? w - j0 : ORC_BLK;
        uint32_t m1[ORC_BLK] = { 0 }, m2[ORC_BLK] = { 0 };
        uint64_t sxx[ORC_BLK] = { 0 }, syy[ORC_BLK] = { 0 }, sxy[ORC_BLK] = { 0 };
        for (int k = 0; k < fw; ++k) {
            const uint32_t fk = f[k];
            const uint16_t *q1 = p_mu1 + j0 + k, *q2 = p_mu2 + j0 + k;
            const uint32_t *qxx = p_xx + j0 + k, *qyy = p_yy + j0 + k, *qxy = p_xy + j0 + k;
            if (nb == ORC_BLK) {
                for (int j = 0; j < ORC_BLK; ++j) {
                    m1[j] += fk * (uint32_t)q1[j]; m2[j] += fk * (uint32_t)q2[j];
                    sxx[j] += (uint64_t)fk * qxx[j]; syy[j] += (uint64_t)fk * qyy[j]; sxy[j] += (uint64_t)fk * qxy[j];
                }
            } else {
                for (int j = 0; j < nb; ++j) {
                    m1[j] += fk * (uint32_t)q1[j]; m2[j] += fk * (uint32_t)q2[j];
                    sxx[j] += (uint64_t)fk * qxx[j]; syy[j] += (uint64_t)fk * qyy[j]; sxy[j] += (uint64_t)fk * qxy[j];
                }
            }
        }
        memcpy(h_mu1 + j0, m1, sizeof(uint32_t) * nb); memcpy(h_mu2 + j0, m2, sizeof(uint32_t) * nb);
        memcpy(h_xx + j0, sxx, sizeof(uint64_t) * nb); memcpy(h_yy + j0, syy, sizeof(uint64_t) * nb);
        memcpy(h_xy + j0, sxy, sizeof(uint64_t) * nb);
    }
}

#define ORC_REFLECT_BORDER(p, w, r)                                                             \
    for (int m = 1; m <= (r); ++m) { (p)[(r) - m] = (p)[(r) + m]; (p)[(r) + (w) - 1 + m] = (p)[(r) + (w) - 1 - m]; }

/* One scale of the statistic.  img is either the source picture (scale 0, bpc samples) or a
 * u16 pyramid level (scale > 0; passed with bpc = 16 semantics: shift 16).                 */
static void vif_statistic(const void *ref, const void *dis, int src_bpc, ptrdiff_t stride,
                          int w, int h, int scale, int pic_bpc, double egl, int64_t acc[7])
{
    const int fw = vif_filter_width[scale], r = fw / 2;
    const uint16_t *f = vif_filter[scale];
    int sh_v, sh_v_sq;
    uint32_t rnd_v;
    uint64_t rnd_v_sq;
    if (scale == 0) {
        sh_v = pic_bpc; rnd_v = 1u << (pic_bpc - 1);
        sh_v_sq = (pic_bpc - 8) * 2; rnd_v_sq = (pic_bpc == 8) ? 0 : (1ull << (sh_v_sq - 1));
    } else {
        sh_v = 16; rnd_v = 32768; sh_v_sq = 16; rnd_v_sq = 32768;
    }
    const int32_t sigma_nsq = 65536 << 1;
    const int bps = src_bpc == 8 ? 1 : 2;
    const size_t pw = (size_t)w + 2 * r;                 /* padded row: r border samples on either side */

    uint16_t *t_mu1 = malloc(sizeof(uint16_t) * pw), *t_mu2 = malloc(sizeof(uint16_t) * pw);
    uint32_t *t_xx = malloc(sizeof(uint32_t) * pw), *t_yy = malloc(sizeof(uint32_t) * pw),
             *t_xy = malloc(sizeof(uint32_t) * pw);
    uint32_t *a_mu1 = malloc(sizeof(uint32_t) * w), *a_mu2 = malloc(sizeof(uint32_t) * w);
    uint64_t *a_xx = malloc(sizeof(uint64_t) * w), *a_yy = malloc(sizeof(uint64_t) * w),
             *a_xy = malloc(sizeof(uint64_t) * w);
    memset(acc, 0, sizeof(int64_t) * 7);

    for (int i = 0; i < h; ++i) {
        /* vertical pass over the (reflect-101 padded) column */
        const void *rows_r[17], *rows_d[17];
        for (int k = 0; k < fw; ++k) {
            const int ii = reflect101(i - r + k, h);
            rows_r[k] = (const uint8_t *)ref + (ptrdiff_t)ii * stride;
            rows_d[k] = (const uint8_t *)dis + (ptrdiff_t)ii * stride;
        }
        orc_vif_vpass(rows_r, rows_d, bps, w, fw, f, a_mu1, a_mu2, a_xx, a_yy, a_xy);
        for (int j = 0; j < w; ++j) {
            t_mu1[r + j] = (uint16_t)((a_mu1[j] + rnd_v) >> sh_v);
            t_mu2[r + j] = (uint16_t)((a_mu2[j] + rnd_v) >> sh_v);
            t_xx[r + j] = (uint32_t)((a_xx[j] + rnd_v_sq) >> sh_v_sq);
            t_yy[r + j] = (uint32_t)((a_yy[j] + rnd_v_sq) >> sh_v_sq);
            t_xy[r + j] = (uint32_t)((a_xy[j] + rnd_v_sq) >> sh_v_sq);
        }
        ORC_REFLECT_BORDER(t_mu1, w, r) ORC_REFLECT_BORDER(t_mu2, w, r)
        ORC_REFLECT_BORDER(t_xx, w, r) ORC_REFLECT_BORDER(t_yy, w, r) ORC_REFLECT_BORDER(t_xy, w, r)
        /* horizontal pass (the accumulator buffers are free again) + statistic */
        orc_vif_hpass(t_mu1, t_mu2, t_xx, t_yy, t_xy, w, fw, f, a_mu1, a_mu2, a_xx, a_yy, a_xy);
        for (int j = 0; j < w; ++j) {
            const uint32_t h_mu1 = a_mu1[j], h_mu2 = a_mu2[j];
            uint32_t mu1_sq = (uint32_t)((((uint64_t)h_mu1 * h_mu1) + 2147483648ull) >> 32);
            uint32_t mu2_sq = (uint32_t)((((uint64_t)h_mu2 * h_mu2) + 2147483648ull) >> 32);
            uint32_t mu1_mu2 = (uint32_t)((((uint64_t)h_mu1 * h_mu2) + 2147483648ull) >> 32);
            uint32_t xx = (uint32_t)((a_xx[j] + 32768) >> 16);
            uint32_t yy = (uint32_t)((a_yy[j] + 32768) >> 16);
            uint32_t xy = (uint32_t)((a_xy[j] + 32768) >> 16);
            int32_t sigma1_sq = (int32_t)(xx - mu1_sq);
            int32_t sigma2_sq = (int32_t)(yy - mu2_sq);
            int32_t sigma12 = (int32_t)(xy - mu1_mu2);
            sigma2_sq = MAXI(sigma2_sq, 0);

            if (sigma1_sq >= sigma_nsq) {
                int x;
                uint16_t d16 = best16_from32((uint32_t)(sigma_nsq + sigma1_sq), &x);
                acc[V_X] += x;
                acc[V_CNT] += 1;
                acc[V_DEN_LOG] += vif_log2_table[d16];
                if (sigma12 > 0 && sigma2_sq > 0) {
                    const double eps = 65536 * 1.0e-10;
                    double g = sigma12 / (sigma1_sq + eps);
                    int32_t sv_sq = (int32_t)(sigma2_sq - g * sigma12);
                    sv_sq = MAXI(sv_sq, 0);
                    g = g < egl ? g : egl;
                    int x1, x2;
                    uint32_t numer1 = (uint32_t)(sv_sq + sigma_nsq);
                    int64_t numer1_tmp = (int64_t)(g * g * sigma1_sq) + numer1;
                    uint16_t n16 = best16_from64((uint64_t)numer1_tmp, &x1);
                    uint16_t m16 = best16_from64((uint64_t)numer1, &x2);
                    acc[V_X2] += (x2 - x1);
                    acc[V_NUM_LOG] += (int64_t)vif_log2_table[n16] - (int64_t)vif_log2_table[m16];
                }
            } else {
                acc[V_NUM_NONLOG] += sigma2_sq;
                acc[V_DEN_NONLOG] += 1;
            }
        }
    }
    free(t_mu1); free(t_mu2); free(t_xx); free(t_yy); free(t_xy);
    free(a_mu1); free(a_mu2); free(a_xx); free(a_yy); free(a_xy);
}

#define ORC_VSUB(T)                                                                             \
    for (int k = 0; k < fw; ++k) {                                                              \
        const T *xr = (const T *)rows_r[k], *yr = (const T *)rows_d[k];                         \
        const uint32_t fk = f[k];                                                               \
        for (int j = 0; j < w; ++j) { ar[j] += fk * (uint32_t)xr[j]; ad[j] += fk * (uint32_t)yr[j]; } \
    }

ORC_SIMD void orc_vif_vsub(const void *const *rows_r, const void *const *rows_d, int bytes_per_sample, int w, int fw,
                           const uint16_t *f, uint32_t *ar, uint32_t *ad)
{
    memset(ar, 0, sizeof(uint32_t) * w); memset(ad, 0, sizeof(uint32_t) * w);
    if (bytes_per_sample == 1) { ORC_VSUB(uint8_t) } else { ORC_VSUB(uint16_t) }
}

/* Filter the level with the NEXT scale's table (V then H) and keep even rows/cols. */
static void vif_subsample(const void *ref, const void *dis, int src_bpc, ptrdiff_t stride,
                          int w, int h, int next_scale, int pic_bpc,
                          uint16_t *oref, uint16_t *odis /* (w/2)*(h/2) tight */)
{
    const int fw = vif_filter_width[next_scale], r = fw / 2;
    const uint16_t *f = vif_filter[next_scale];
    int sh_v; uint32_t rnd_v;
    if (next_scale == 1) { sh_v = pic_bpc; rnd_v = 1u << (pic_bpc - 1); }
    else { sh_v = 16; rnd_v = 32768; }
    const int bps = src_bpc == 8 ? 1 : 2;
    const size_t pw = (size_t)w + 2 * r;
    uint16_t *tr = malloc(sizeof(uint16_t) * pw), *td = malloc(sizeof(uint16_t) * pw);
    uint32_t *ar = malloc(sizeof(uint32_t) * w), *ad = malloc(sizeof(uint32_t) * w);
    const int ow = w / 2, oh = h / 2;
    for (int oi = 0; oi < oh; ++oi) {
        const int i = 2 * oi;
        const void *rows_r[17], *rows_d[17];
        for (int k = 0; k < fw; ++k) {
            const int ii = reflect101(i - r + k, h);
            rows_r[k] = (const uint8_t *)ref + (ptrdiff_t)ii * stride;
            rows_d[k] = (const uint8_t *)dis + (ptrdiff_t)ii * stride;
        }
        orc_vif_vsub(rows_r, rows_d, bps, w, fw, f, ar, ad);
        for (int j = 0; j < w; ++j) {
            tr[r + j] = (uint16_t)((ar[j] + rnd_v) >> sh_v);
            td[r + j] = (uint16_t)((ad[j] + rnd_v) >> sh_v);
        }
        ORC_REFLECT_BORDER(tr, w, r) ORC_REFLECT_BORDER(td, w, r)
        for (int oj = 0; oj < ow; ++oj) {
            const int j = 2 * oj;
            uint32_t sr = 0, sd = 0;
            for (int k = 0; k < fw; ++k) {
                sr += f[k] * (uint32_t)tr[j + k];
                sd += f[k] * (uint32_t)td[j + k];
            }
            oref[(size_t)oi * ow + oj] = (uint16_t)((sr + 32768) >> 16);
            odis[(size_t)oi * ow + oj] = (uint16_t)((sd + 32768) >> 16);
        }
    }
    free(tr); free(td); free(ar); free(ad);
}

/* Derived per-scale num/den: stored through float like libvmaf's VifScore {float num, den}. */
ORC_API void orc_vif_finish(const int64_t acc[7], double *num, double *den)
{
    float n = (float)(acc[V_NUM_LOG] / 2048.0 + (double)acc[V_X2] +
                      ((double)acc[V_DEN_NONLOG] - ((double)acc[V_NUM_NONLOG] / 16384.0) / 65025.0));
    float d = (float)(acc[V_DEN_LOG] / 2048.0 - ((double)acc[V_X] + (double)(acc[V_CNT] * 17)) +
                      (double)acc[V_DEN_NONLOG]);
    *num = n;
    *den = d;
}

ORC_API int orc_vif(const void *ref, const void *dis, int bpc, int w, int h, ptrdiff_t stride,
                    double egl, int64_t acc[4][7], double num[4], double den[4], double score[4])
{
    vif_log2_init();
    if (w < 32 || h < 32) return -1;
    uint16_t *lr[2] = { NULL, NULL }, *ld[2] = { NULL, NULL };
    const void *cr = ref, *cd = dis;
    int cbpc = bpc; ptrdiff_t cstride = stride;
    int cw = w, ch = h;
    for (int scale = 0; scale < 4; ++scale) {
        if (scale > 0) {
            int ow = cw / 2, oh = ch / 2;
            uint16_t *nr = malloc(sizeof(uint16_t) * (size_t)ow * oh);
            uint16_t *nd = malloc(sizeof(uint16_t) * (size_t)ow * oh);
            vif_subsample(cr, cd, cbpc, cstride, cw, ch, scale, bpc, nr, nd);
            free(lr[0]); free(ld[0]);
            lr[0] = nr; ld[0] = nd;
            cr = nr; cd = nd; cbpc = 16; cstride = (ptrdiff_t)ow * 2; cw = ow; ch = oh;
        }
        vif_statistic(cr, cd, cbpc, cstride, cw, ch, scale, bpc, egl, acc[scale]);
        orc_vif_finish(acc[scale], &num[scale], &den[scale]);
        score[scale] = (double)((float)num[scale] / (float)den[scale]);
    }
    free(lr[0]); free(ld[0]);
    return 0;
}

/* ==========================================================================================
 * Integer ADM (libvmaf integer_adm.c).
 * ======================================================================================== */
static const int32_t dwt_lo[4] = { SPEC_DWT_LO_Q15 };
static const int32_t dwt_hi[4] = { SPEC_DWT_HI_Q15 };
static const int32_t dwt_lo_sum = SPEC_DWT_LO_SUM_Q15;

#define ADM_BORDER_FACTOR SPEC_ADM_BORDER_FACTOR

static int32_t adm_div_lookup[65537];
static int adm_div_ready = 0;
static void adm_div_init(void)
{
    if (adm_div_ready) return;
    adm_div_lookup[32768] = 0;
    for (int i = 1; i <= 32768; ++i) {
        int32_t recip = (int32_t)(1073741824 / i);
        adm_div_lookup[32768 + i] = recip;
        adm_div_lookup[32768 - i] = 0 - recip;
    }
    adm_div_ready = 1;
}

/* Watson DWT quantisation step (float evaluation as in libvmaf adm_tools.h dwt_quant_step). */
struct dwt_model_params { float a, k, f0, g[4]; };
static const struct dwt_model_params dwt_7_9_Y = { SPEC_DWT79_A, SPEC_DWT79_K, SPEC_DWT79_F0, { SPEC_DWT79_G } };
static const float dwt_7_9_amp[4][4] = { SPEC_DWT79_AMP };
static float dwt_quant_step(int lambda, int theta, double view_dist, int display_h)
{
    float r = view_dist * display_h * M_PI / 180.0;
    float temp = log10(pow(2.0, lambda + 1) * dwt_7_9_Y.f0 * dwt_7_9_Y.g[theta] / r);
    float Q = 2.0 * dwt_7_9_Y.a * pow(10.0, dwt_7_9_Y.k * temp * temp) / dwt_7_9_amp[lambda][theta];
    return Q;
}

ORC_API void orc_adm_rfactor(int scale, double view_dist, int display_h, float rf[3])
{
    float f1 = dwt_quant_step(scale, 1, view_dist, display_h);
    float f2 = dwt_quant_step(scale, 2, view_dist, display_h);
    rf[0] = 1.0f / f1; rf[1] = 1.0f / f1; rf[2] = 1.0f / f2;
}

typedef struct { int32_t *a, *v, *h, *d; } bands_t;  /* i32 storage for every scale */

static void bands_alloc(bands_t *b, size_t n)
{
    b->a = malloc(sizeof(int32_t) * n); b->v = malloc(sizeof(int32_t) * n);
    b->h = malloc(sizeof(int32_t) * n); b->d = malloc(sizeof(int32_t) * n);
}
static void bands_free(bands_t *b) { free(b->a); free(b->v); free(b->h); free(b->d); }

/* scale 0: source picture -> int16-valued bands (stored in i32) */
static void adm_dwt_s0(const void *src, int bpc, ptrdiff_t stride, int w, int h, bands_t *dst)
{
    const int ow = (w + 1) / 2, oh = (h + 1) / 2;
    const int32_t add_v = 1 << (bpc - 1);
    int16_t *tlo = malloc(sizeof(int16_t) * w), *thi = malloc(sizeof(int16_t) * w);
    for (int i = 0; i < oh; ++i) {
        int iy[4];
        for (int k = 0; k < 4; ++k) iy[k] = mirror_asym(2 * i - 1 + k, h);
        for (int j = 0; j < w; ++j) {
            int32_t s[4];
            for (int k = 0; k < 4; ++k) s[k] = (int32_t)px(src, bpc, stride, iy[k], j);
            int32_t acc = 0;
            for (int k = 0; k < 4; ++k) acc += dwt_lo[k] * s[k];
            acc -= dwt_lo_sum * add_v;      /* (0..N) -> (-N/2..N/2) */
            tlo[j] = (int16_t)((acc + add_v) >> bpc);
            acc = 0;
            for (int k = 0; k < 4; ++k) acc += dwt_hi[k] * s[k];
            thi[j] = (int16_t)((acc + add_v) >> bpc);
        }
        for (int j = 0; j < ow; ++j) {
            int jx[4];
            for (int k = 0; k < 4; ++k) jx[k] = mirror_asym(2 * j - 1 + k, w);
            int32_t acc;
            acc = 0; for (int k = 0; k < 4; ++k) acc += dwt_lo[k] * (int32_t)tlo[jx[k]];
            dst->a[(size_t)i * ow + j] = (int16_t)((acc + 32768) >> 16);
            acc = 0; for (int k = 0; k < 4; ++k) acc += dwt_hi[k] * (int32_t)tlo[jx[k]];
            dst->v[(size_t)i * ow + j] = (int16_t)((acc + 32768) >> 16);
            acc = 0; for (int k = 0; k < 4; ++k) acc += dwt_lo[k] * (int32_t)thi[jx[k]];
            dst->h[(size_t)i * ow + j] = (int16_t)((acc + 32768) >> 16);
            acc = 0; for (int k = 0; k < 4; ++k) acc += dwt_hi[k] * (int32_t)thi[jx[k]];
            dst->d[(size_t)i * ow + j] = (int16_t)((acc + 32768) >> 16);
        }
    }
    free(tlo); free(thi);
}

/* scales 1..3: previous band_a (i32) -> i32 bands */
static void adm_dwt_s123(const int32_t *src, int w, int h, int scale, bands_t *dst)
{
    static const int64_t rnd_v[3] = { 0, 32768, 32768 };
    static const int64_t rnd_h[3] = { 16384, 32768, 16384 };
    static const int sh_v[3] = { SPEC_ADM_DWT_SH_V };
    static const int sh_h[3] = { SPEC_ADM_DWT_SH_H };
    const int ow = (w + 1) / 2, oh = (h + 1) / 2, s = scale - 1;
    int32_t *tlo = malloc(sizeof(int32_t) * w), *thi = malloc(sizeof(int32_t) * w);
    for (int i = 0; i < oh; ++i) {
        int iy[4];
        for (int k = 0; k < 4; ++k) iy[k] = mirror_asym(2 * i - 1 + k, h);
        for (int j = 0; j < w; ++j) {
            int64_t acc = 0;
            for (int k = 0; k < 4; ++k) acc += (int64_t)dwt_lo[k] * src[(size_t)iy[k] * w + j];
            tlo[j] = (int32_t)((acc + rnd_v[s]) >> sh_v[s]);
            acc = 0;
            for (int k = 0; k < 4; ++k) acc += (int64_t)dwt_hi[k] * src[(size_t)iy[k] * w + j];
            thi[j] = (int32_t)((acc + rnd_v[s]) >> sh_v[s]);
        }
        for (int j = 0; j < ow; ++j) {
            int jx[4];
            for (int k = 0; k < 4; ++k) jx[k] = mirror_asym(2 * j - 1 + k, w);
            int64_t acc;
            acc = 0; for (int k = 0; k < 4; ++k) acc += (int64_t)dwt_lo[k] * tlo[jx[k]];
            dst->a[(size_t)i * ow + j] = (int32_t)((acc + rnd_h[s]) >> sh_h[s]);
            acc = 0; for (int k = 0; k < 4; ++k) acc += (int64_t)dwt_hi[k] * tlo[jx[k]];
            dst->v[(size_t)i * ow + j] = (int32_t)((acc + rnd_h[s]) >> sh_h[s]);
            acc = 0; for (int k = 0; k < 4; ++k) acc += (int64_t)dwt_lo[k] * thi[jx[k]];
            dst->h[(size_t)i * ow + j] = (int32_t)((acc + rnd_h[s]) >> sh_h[s]);
            acc = 0; for (int k = 0; k < 4; ++k) acc += (int64_t)dwt_hi[k] * thi[jx[k]];
            dst->d[(size_t)i * ow + j] = (int32_t)((acc + rnd_h[s]) >> sh_h[s]);
        }
    }
    free(tlo); free(thi);
}

static inline uint16_t best15_from32(uint32_t v, int *x)
{
    int k = 17 - __builtin_clz(v);
    v = (v + (1u << (k - 1))) >> k;
    *x = k;
    return (uint16_t)v;
}

/* k = clamp(t/o, 0, 1) in Q15 via the reciprocal table, then restored value r = k*o. */
static inline int32_t adm_k_s0(int32_t o, int32_t t)
{
    int32_t tmp = (o == 0) ? 32768 : (int32_t)((((int64_t)adm_div_lookup[o + 32768] * t) + 16384) >> 15);
    return tmp < 0 ? 0 : (tmp > 32768 ? 32768 : tmp);
}
static inline int64_t adm_k_s123(int32_t o, int32_t t)
{
    if (o == 0) return 32768;
    int sh = 0;
    uint32_t ao = (uint32_t)abs(o);
    int sign = o < 0 ? -1 : 1;
    uint16_t msb = ao < 32768 ? (uint16_t)ao : best15_from32(ao, &sh);
    int64_t tmp = (((int64_t)adm_div_lookup[msb + 32768] * t) * sign + ((int64_t)1 << (14 + sh))) >> (15 + sh);
    return tmp < 0 ? 0 : (tmp > 32768 ? 32768 : tmp);
}

static inline int adm_angle_flag(int64_t ot_dp, int64_t o_mag_sq, int64_t t_mag_sq, float cos_1deg_sq)
{
    return (((float)ot_dp / 4096.0) >= 0.0f) &&
           (((float)ot_dp / 4096.0) * ((float)ot_dp / 4096.0) >=
            cos_1deg_sq * ((float)o_mag_sq / 4096.0) * ((float)t_mag_sq / 4096.0));
}

/* Decouple one pixel of one scale; outputs restored r[3] and additive a[3] in (h, v, d) order */
static inline void adm_decouple_px(int scale, const int32_t o[3], const int32_t t[3], double egl,
                                   float cos_1deg_sq, int32_t r[3], int32_t a[3])
{
    int64_t ot_dp = (int64_t)o[0] * t[0] + (int64_t)o[1] * t[1];
    int64_t o_mag = (int64_t)o[0] * o[0] + (int64_t)o[1] * o[1];
    int64_t t_mag = (int64_t)t[0] * t[0] + (int64_t)t[1] * t[1];
    int flag = adm_angle_flag(ot_dp, o_mag, t_mag, cos_1deg_sq);
    for (int b = 0; b < 3; ++b) {
        int64_t k = scale == 0 ? adm_k_s0(o[b], t[b]) : adm_k_s123(o[b], t[b]);
        int32_t rst = (int32_t)(((k * o[b]) + 16384) >> 15);
        if (scale == 0) rst = (int16_t)rst;
        const float rst_f = ((float)k / 32768) * ((float)o[b] / 64);
        if (flag && rst_f > 0.) { double v = rst * egl; double tt = t[b]; rst = (int32_t)(v < tt ? v : tt); }
        if (flag && rst_f < 0.) { double v = rst * egl; double tt = t[b]; rst = (int32_t)(v > tt ? v : tt); }
        if (scale == 0) rst = (int16_t)rst;
        r[b] = rst;
        a[b] = t[b] - rst;
        if (scale == 0) a[b] = (int16_t)a[b];
    }
}

static inline int32_t shl32(int32_t v, int s) { return (int32_t)((uint32_t)v << s); } /* wraps like x86 */

/* One scale: decouple + csf + contrast masking numerator, and the csf denominator.
 * cm[3], dn[3] are the row-shifted integer accumulators for (h, v, d). */
static void adm_scale(const bands_t *ref, const bands_t *dis, int w, int h, int scale, double egl,
                      double view_dist, int display_h, int64_t cm[3], uint64_t dn[3],
                      float *num_scale, float *den_scale)
{
    const float cos_1deg_sq = cos(1.0 * M_PI / 180.0) * cos(1.0 * M_PI / 180.0);
    float rf[3];
    orc_adm_rfactor(scale, view_dist, display_h, rf);

    /* fixed-point csf factors */
    uint32_t i_rf[3];
    if (scale == 0) {
        if (fabs(view_dist * display_h - 3.0 * 1080) < 1.0e-8) {
            static const uint32_t s0_rf[3] = { SPEC_ADM_S0_RF };
            i_rf[0] = s0_rf[0]; i_rf[1] = s0_rf[1]; i_rf[2] = s0_rf[2];
        } else {
            i_rf[0] = (uint16_t)(rf[0] * pow(2, 21));
            i_rf[1] = (uint16_t)(rf[1] * pow(2, 21));
            i_rf[2] = (uint16_t)(rf[2] * pow(2, 23));
        }
    } else {
        for (int b = 0; b < 3; ++b) i_rf[b] = (uint32_t)(rf[b] * pow(2, 32));
    }

    /* region of interest (cm / den), and the one-tap-grown region for decouple + csf */
    int left = w * ADM_BORDER_FACTOR - 0.5;
    int top = h * ADM_BORDER_FACTOR - 0.5;
    int right = w - left;
    int bottom = h - top;
    int gl = MAXI(left - 1, 0), gt = MAXI(top - 1, 0), gr = MINI(right + 1, w), gb = MINI(bottom + 1, h);

    const size_t n = (size_t)w * h;
    int32_t *R[3], *CA[3], *CF[3];
    for (int b = 0; b < 3; ++b) {
        R[b] = calloc(n, sizeof(int32_t)); CA[b] = calloc(n, sizeof(int32_t)); CF[b] = calloc(n, sizeof(int32_t));
    }
    const int32_t *OB[3] = { ref->h, ref->v, ref->d };
    const int32_t *TB[3] = { dis->h, dis->v, dis->d };

    /* scale-0 csf constants */
    static const int s0_shift[3] = { SPEC_ADM_S0_RF_SHIFT };
    static const int32_t s0_add[3] = { SPEC_ADM_S0_RF_ROUND };

    for (int i = gt; i < gb; ++i)
        for (int j = gl; j < gr; ++j) {
            size_t p = (size_t)i * w + j;
            int32_t o[3] = { OB[0][p], OB[1][p], OB[2][p] };
            int32_t t[3] = { TB[0][p], TB[1][p], TB[2][p] };
            int32_t r[3], a[3];
            adm_decouple_px(scale, o, t, egl, cos_1deg_sq, r, a);
            for (int b = 0; b < 3; ++b) {
                R[b][p] = r[b];
                if (scale == 0) {
                    int32_t dv = (int32_t)i_rf[b] * a[b];
                    int16_t ca = (int16_t)((dv + s0_add[b]) >> s0_shift[b]);
                    CA[b][p] = ca;
                    CF[b][p] = (int16_t)(((SPEC_ADM_ONE_BY_30_Q16 * abs((int32_t)ca)) + 2048) >> 12);
                } else {
                    int32_t ca = (int32_t)((((int64_t)i_rf[b] * (int64_t)a[b]) + (1ll << 27)) >> 28);
                    CA[b][p] = ca;
                    CF[b][p] = (int32_t)((((int64_t)SPEC_ADM_ONE_BY_30_Q32 * abs(ca)) + (1ll << 31)) >> 32);
                }
            }
        }

    /* ---- contrast-masked numerator ---- */
    int sh_sub[3], sh_sq[3], sh_cub[3];
    int64_t add_sq[3], add_cub[3];
    int sh_inner;
    if (scale == 0) {
        sh_sub[0] = 10; sh_sub[1] = 10; sh_sub[2] = 12;
        sh_sq[0] = 29; sh_sq[1] = 29; sh_sq[2] = 30;
        sh_cub[0] = sh_cub[1] = (int)(uint32_t)ceil(log2(w) - 4);
        sh_cub[2] = (int)(uint32_t)ceil(log2(w) - 3);
    } else {
        for (int b = 0; b < 3; ++b) { sh_sub[b] = 0; sh_sq[b] = 30; sh_cub[b] = (int)(uint32_t)ceil(log2(w)); }
    }
    for (int b = 0; b < 3; ++b) {
        add_sq[b] = (int64_t)1 << (sh_sq[b] - 1);
        add_cub[b] = (int64_t)(uint32_t)pow(2, (sh_cub[b] - 1));
    }
    sh_inner = (int)(uint32_t)ceil(log2(h));
    const int64_t add_inner = (int64_t)(uint32_t)pow(2, (sh_inner - 1));

    cm[0] = cm[1] = cm[2] = 0;
    for (int i = top; i < bottom; ++i) {
        int64_t inner[3] = { 0, 0, 0 };
        for (int j = left; j < right; ++j) {
            size_t p = (size_t)i * w + j;
            /* threshold: 3x3 neighbourhood of csf_f (centre replaced by |csf_a|/15), 3 bands */
            int32_t thr = 0;
            for (int b = 0; b < 3; ++b) {
                int32_t sum = 0;
                for (int di = -1; di <= 1; ++di)
                    for (int dj = -1; dj <= 1; ++dj) {
                        int ii = mirror_asym(i + di, h), jj = mirror_asym(j + dj, w);
                        size_t q = (size_t)ii * w + jj;
                        if (di == 0 && dj == 0) {
                            if (scale == 0)
                                sum += (int16_t)(((SPEC_ADM_ONE_BY_15_Q16 * abs(CA[b][q])) + 2048) >> 12);
                            else
                                sum += (int32_t)((((int64_t)SPEC_ADM_ONE_BY_15_Q32 * abs(CA[b][q])) + (1ll << 31)) >> 32);
                        } else {
                            sum += CF[b][q];
                        }
                    }
                thr += sum;
            }
            for (int b = 0; b < 3; ++b) {
                int32_t x;
                if (scale == 0)
                    x = R[b][p] * (int32_t)i_rf[b];
                else
                    x = (int32_t)((((int64_t)R[b][p] * i_rf[b]) + (1ll << 27)) >> 28);
                x = abs(x) - shl32(thr, sh_sub[b]);
                x = x < 0 ? 0 : x;
                int32_t x_sq = (int32_t)((((int64_t)x * x) + add_sq[b]) >> sh_sq[b]);
                int64_t val = (((int64_t)x_sq * x) + add_cub[b]) >> sh_cub[b];
                inner[b] += val;
            }
        }
        for (int b = 0; b < 3; ++b) cm[b] += (inner[b] + add_inner) >> sh_inner;
    }
    {
        float f_acc[3];
        if (scale == 0) {
            f_acc[0] = (float)(cm[0] / pow(2, (52 - sh_cub[0] - sh_inner)));
            f_acc[1] = (float)(cm[1] / pow(2, (52 - sh_cub[1] - sh_inner)));
            f_acc[2] = (float)(cm[2] / pow(2, (57 - sh_cub[2] - sh_inner)));
        } else {
            static const int fs[3] = { 45, 39, 36 };
            float final_shift = pow(2, (fs[scale - 1] - sh_cub[0] - sh_inner));
            for (int b = 0; b < 3; ++b) f_acc[b] = (float)(cm[b] / final_shift);
        }
        float area_term = powf((bottom - top) * (right - left) / 32.0f, 1.0f / 3.0f);
        float nh = powf(f_acc[0], 1.0f / 3.0f) + area_term;
        float nv = powf(f_acc[1], 1.0f / 3.0f) + area_term;
        float nd = powf(f_acc[2], 1.0f / 3.0f) + area_term;
        *num_scale = nh + nv + nd;
    }

    /* ---- csf denominator on the reference bands ---- */
    dn[0] = dn[1] = dn[2] = 0;
    if (scale == 0) {
        int32_t sh_acc = (int32_t)ceil(log2((bottom - top) * (right - left)) - 20);
        sh_acc = sh_acc > 0 ? sh_acc : 0;
        uint64_t add_acc = sh_acc > 0 ? (1ull << (sh_acc - 1)) : 0;
        for (int i = top; i < bottom; ++i) {
            uint64_t inner[3] = { 0, 0, 0 };
            for (int j = left; j < right; ++j) {
                size_t p = (size_t)i * w + j;
                for (int b = 0; b < 3; ++b) {
                    uint16_t v = (uint16_t)abs(OB[b][p]);
                    inner[b] += ((uint64_t)v * v) * v;
                }
            }
            for (int b = 0; b < 3; ++b) dn[b] += (inner[b] + add_acc) >> sh_acc;
        }
        double shift_csf = pow(2, (18 - sh_acc));
        float area_term = powf((bottom - top) * (right - left) / 32.0f, 1.0f / 3.0f);
        float s = 0;
        float part[3];
        for (int b = 0; b < 3; ++b) {
            double csf = (double)(dn[b] / shift_csf) * pow(rf[b], 3);
            part[b] = powf(csf, 1.0f / 3.0f) + area_term;
        }
        s = part[0] + part[1] + part[2];
        *den_scale = s;
    } else {
        static const int sh_sqd[3] = { 31, 30, 31 };
        static const int conv[3] = { 32, 27, 23 };
        const uint64_t add_sqd = 1ull << (sh_sqd[scale - 1] - 1);
        uint32_t sh_c = (uint32_t)ceil(log2(right - left));
        uint64_t add_c = (uint64_t)(uint32_t)pow(2, (sh_c - 1));
        uint32_t sh_acc = (uint32_t)ceil(log2(bottom - top));
        uint64_t add_acc = (uint64_t)(uint32_t)pow(2, (sh_acc - 1));
        for (int i = top; i < bottom; ++i) {
            uint64_t inner[3] = { 0, 0, 0 };
            for (int j = left; j < right; ++j) {
                size_t p = (size_t)i * w + j;
                for (int b = 0; b < 3; ++b) {
                    uint32_t v = (uint32_t)abs(OB[b][p]);
                    uint64_t val = ((((((uint64_t)v * v) + add_sqd) >> sh_sqd[scale - 1]) * v) + add_c) >> sh_c;
                    inner[b] += val;
                }
            }
            for (int b = 0; b < 3; ++b) dn[b] += (inner[b] + add_acc) >> sh_acc;
        }
        double shift_csf = pow(2, (conv[scale - 1] - (int)sh_acc - (int)sh_c));
        float area_term = powf((bottom - top) * (right - left) / 32.0f, 1.0f / 3.0f);
        float part[3];
        for (int b = 0; b < 3; ++b) {
            double csf = (double)(dn[b] / shift_csf) * pow(rf[b], 3);
            part[b] = powf(csf, 1.0f / 3.0f) + area_term;
        }
        *den_scale = part[0] + part[1] + part[2];
    }

    for (int b = 0; b < 3; ++b) { free(R[b]); free(CA[b]); free(CF[b]); }
}

/* Full integer ADM for one frame pair.  Returns raw accumulators and the derived scores. */
ORC_API int orc_adm(const void *ref, const void *dis, int bpc, int w, int h, ptrdiff_t stride,
                    double egl, double view_dist, int display_h,
                    int64_t cm_acc[4][3], uint64_t den_acc[4][3],
                    double num_scale[4], double den_scale[4], double *adm2,
                    int32_t *dbg_ref_bands /* optional: scale-0 a,v,h,d of ref, 4*ow*oh */)
{
    adm_div_init();
    if (w < 32 || h < 32) return -1;
    const double numden_limit = 1e-10 * (w * h) / (1920.0 * 1080.0);
    bands_t rb[2], db[2];
    int cw = w, ch = h;
    double num = 0, den = 0;
    int cur = 0;
    for (int scale = 0; scale < 4; ++scale) {
        const int ow = (cw + 1) / 2, oh = (ch + 1) / 2;
        bands_alloc(&rb[cur], (size_t)ow * oh);
        bands_alloc(&db[cur], (size_t)ow * oh);
        if (scale == 0) {
            adm_dwt_s0(ref, bpc, stride, cw, ch, &rb[cur]);
            adm_dwt_s0(dis, bpc, stride, cw, ch, &db[cur]);
            if (dbg_ref_bands) {
                size_t n = (size_t)ow * oh;
                memcpy(dbg_ref_bands, rb[cur].a, n * 4); memcpy(dbg_ref_bands + n, rb[cur].v, n * 4);
                memcpy(dbg_ref_bands + 2 * n, rb[cur].h, n * 4); memcpy(dbg_ref_bands + 3 * n, rb[cur].d, n * 4);
            }
        } else {
            adm_dwt_s123(rb[cur ^ 1].a, cw, ch, scale, &rb[cur]);
            adm_dwt_s123(db[cur ^ 1].a, cw, ch, scale, &db[cur]);
            bands_free(&rb[cur ^ 1]); bands_free(&db[cur ^ 1]);
        }
        cw = ow; ch = oh;
        float ns, ds;
        adm_scale(&rb[cur], &db[cur], cw, ch, scale, egl, view_dist, display_h,
                  cm_acc[scale], den_acc[scale], &ns, &ds);
        num += ns; den += ds;
        num_scale[scale] = ns; den_scale[scale] = ds;
        cur ^= 1;
    }
    bands_free(&rb[cur ^ 1]); bands_free(&db[cur ^ 1]);
    num = num < numden_limit ? 0 : num;
    den = den < numden_limit ? 0 : den;
    *adm2 = (den == 0.0) ? 1.0 : num / den;
    return 0;
}

/* ==========================================================================================
 * PSNR (libvmaf integer_psnr.c and FFmpeg vf_psnr.c share the SSE)
 * ======================================================================================== */
ORC_API uint64_t orc_sse(const void *a, const void *b, int bpc, int w, int h, ptrdiff_t stride)
{
    uint64_t sse = 0;
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            int64_t d = (int64_t)px(a, bpc, stride, i, j) - (int64_t)px(b, bpc, stride, i, j);
            sse += (uint64_t)(d * d);
        }
    return sse;
}

/* libvmaf psnr: min(10*log10(peak^2 / mse), 6*bpc + 12) with mse = sse / (w*h) */
ORC_API double orc_psnr_from_sse(uint64_t sse, int bpc, int w, int h)
{
    const double peak = (double)((1 << bpc) - 1);
    const double psnr_max = 6.0 * bpc + 12.0;
    const double mse = (double)sse / ((double)w * h);
    if (sse == 0) return psnr_max;
    double v = 10.0 * log10(peak * peak / mse);
    return v < psnr_max ? v : psnr_max;
}

/* ==========================================================================================
 * FFmpeg `ssim` filter: x264-style integer SSIM on 4x4 block sums, 8x8 overlapped windows
 * (libavfilter/vf_ssim.c).  One plane.
 * ======================================================================================== */
ORC_API double orc_ffssim_plane(const void *a, const void *b, int bpc, int w, int h, ptrdiff_t stride)
{
    const int W4 = w >> 2, H4 = h >> 2;
    if (W4 < 2 || H4 < 2) return 1.0;
    const int maxv = (1 << bpc) - 1;
    int64_t (*sums)[4] = malloc(sizeof(int64_t[4]) * (size_t)W4 * H4);
    for (int by = 0; by < H4; ++by)
        for (int bx = 0; bx < W4; ++bx) {
            int64_t s1 = 0, s2 = 0, ss = 0, s12 = 0;
            for (int y = 0; y < 4; ++y)
                for (int x = 0; x < 4; ++x) {
                    int64_t p = px(a, bpc, stride, by * 4 + y, bx * 4 + x);
                    int64_t q = px(b, bpc, stride, by * 4 + y, bx * 4 + x);
                    s1 += p; s2 += q; ss += p * p + q * q; s12 += p * q;
                }
            int64_t *d = sums[(size_t)by * W4 + bx];
            d[0] = s1; d[1] = s2; d[2] = ss; d[3] = s12;
        }
    double total = 0.0;
    for (int by = 0; by < H4 - 1; ++by) {
        float row = 0.0f;   /* vf_ssim accumulates each row's windows in float (ssim_endn) */
        double rowd = 0.0;
        for (int bx = 0; bx < W4 - 1; ++bx) {
            int64_t s1 = 0, s2 = 0, ss = 0, s12 = 0;
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    const int64_t *d = sums[(size_t)(by + dy) * W4 + bx + dx];
                    s1 += d[0]; s2 += d[1]; ss += d[2]; s12 += d[3];
                }
            if (bpc == 8) {
                const int c1 = SPEC_FFSSIM_C1, c2 = SPEC_FFSSIM_C2;
                int fs1 = (int)s1, fs2 = (int)s2, fss = (int)ss, fs12 = (int)s12;
                int vars = fss * 64 - fs1 * fs1 - fs2 * fs2;
                int covar = fs12 * 64 - fs1 * fs2;
                float v = (float)(2 * fs1 * fs2 + c1) * (float)(2 * covar + c2) /
                          ((float)(fs1 * fs1 + fs2 * fs2 + c1) * (float)(vars + c2));
                row += v;
                (void)rowd;
            } else {
                /* ssim_end1x(): int64 block sums, integer c1/c2 from max = 2^depth - 1 */
                const int64_t c1 = (int64_t)(.01 * .01 * maxv * maxv * 64 + .5);
                const int64_t c2 = (int64_t)(.03 * .03 * maxv * maxv * 64 * 63 + .5);
                int64_t vars = ss * 64 - s1 * s1 - s2 * s2;
                int64_t covar = s12 * 64 - s1 * s2;
                float v = (float)(2 * s1 * s2 + c1) * (float)(2 * covar + c2) /
                          ((float)(s1 * s1 + s2 * s2 + c1) * (float)(vars + c2));
                row += v;
            }
        }
        total += row;
    }
    free(sums);
    return total / ((double)(H4 - 1) * (W4 - 1));
}

/* ==========================================================================================
 * SVR prediction (libsvm svm_predict for nu-SVR/RBF + libvmaf predict.c normalisation;
 * SURVEY.md Appendix A.6).  sv is dense [n_sv][n_feat] (missing sparse indices = 0).
 * ======================================================================================== */
ORC_API double orc_svr_predict(const double *feat, int n_feat, const double *slopes /* n_feat+1 */,
                               const double *intercepts /* n_feat+1 */, const double *sv,
                               const double *coef, int n_sv, double gamma, double rho)
{
    double x[64];
    for (int i = 0; i < n_feat; ++i) x[i] = slopes[i + 1] * feat[i] + intercepts[i + 1];
    double sum = 0;
    for (int k = 0; k < n_sv; ++k) {
        double d2 = 0;
        for (int i = 0; i < n_feat; ++i) {
            double d = x[i] - sv[(size_t)k * n_feat + i];
            d2 += d * d;
        }
        sum += coef[k] * exp(-gamma * d2);
    }
    sum -= rho;
    return (sum - intercepts[0]) / slopes[0];
}
