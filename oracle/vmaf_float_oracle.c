/*
 * oracle/vmaf_float_oracle.c -- CPU ORACLE for the float extractors (test infrastructure, NOT the
 * product; see the header of vmaf_oracle.c for who may call it).
 *
 * Restates the float32 feature extractors the `vmaf_float_*` models use and libvmaf's `ssim=1` /
 * `ms_ssim=1` options (reference call site: app/vmaf_analyzer.py:373-419, option `ssim=1` at :386):
 *   libvmaf src/feature/vif.c + vif_tools.c      (float_vif)      SURVEY.md Appendix A.5
 *   libvmaf src/feature/adm.c + adm_tools.c      (float_adm)      SURVEY.md Appendix A.5
 *   libvmaf src/feature/motion.c                 (float_motion)   SURVEY.md Appendix A.5
 *   libvmaf src/feature/ssim.c, ms_ssim.c, iqa/  (float_ssim, float_ms_ssim)  SURVEY.md Appendix A.9
 *
 * PARITY UNPINNED: libvmaf is absent from /root/reference (not vendored, not pinned); this file is
 * written from the published algorithm.  The GPU kernels are held to it within the north star's
 * float tolerance (1e-4 per-frame VMAF, 1e-5 pooled), not bit-exactly.
 */
#include "../include/libvmaf_spec.h"
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline int mirror(int i, int n)          /* -i -> i ; n+i -> n-1-i */
{
    if (i < 0) return -i;
    if (i >= n) return 2 * n - i - 1;
    return i;
}

/* picture_copy(): luma -> float, (v / 2^(bpc-8)) + offset */
ORC_API void orc_f_picture_copy(const void *src, int bpc, int w, int h, ptrdiff_t stride, float offset, float *dst)
{
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            float v;
            if (bpc == 8) v = (float)((const uint8_t *)src)[(ptrdiff_t)i * stride + j];
            else v = (float)*(const uint16_t *)((const uint8_t *)src + (ptrdiff_t)i * stride + 2 * j) / (float)(1 << (bpc - 8));
            dst[(size_t)i * w + j] = v + offset;
        }
}

/* ============================== float VIF ============================================== */
static const float vif_f17[17] = { SPEC_VIF_F32_17 };
static const float vif_f9[9] = { SPEC_VIF_F32_9 };
static const float vif_f5[5] = { SPEC_VIF_F32_5 };
static const float vif_f3[3] = { SPEC_VIF_F32_3 };
static const float *const vif_ftab[4] = { vif_f17, vif_f9, vif_f5, vif_f3 };
static const int vif_fw[4] = { 17, 9, 5, 3 };

ORC_API const float *orc_f_vif_filter(int scale) { return vif_ftab[scale]; }

/* vif_tools.c vif_filter1d_s(): vertical then horizontal, mirrored borders, full-size output.  A row at a time, tap loop
 * outside and column loop inside, so that gcc vectorises ACROSS pixels: every pixel still sees acc = 0, then
 * acc += f[k] * v[k] for k = 0 .. fw-1 with each product and each sum rounded to float -- the scalar C order, bit for
 * bit (-ffp-contract=off).  The two inner loops are built for AVX2 and for the baseline ISA; the loader picks. */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define ORC_SIMD __attribute__((target_clones("avx2", "default")))
#else
#define ORC_SIMD
#endif

#define ORC_FBLK 32          /* pixels per block: the accumulators stay in vector registers across the tap loop */

ORC_SIMD void orc_f_fir_rows(const float *const *rows, const float *f, int fw, int w, float *restrict out)
{
    for (int j0 = 0; j0 < w; j0 += ORC_FBLK) {
        const int nb = w - j0 < ORC_FBLK ? w - j0 : ORC_FBLK;
        float acc[ORC_FBLK] = { 0.0f };
        for (int k = 0; k < fw; ++k) {
            const float *restrict v = rows[k] + j0;
            const float fk = f[k];
            if (nb == ORC_FBLK) { for (int j = 0; j < ORC_FBLK; ++j) acc[j] += fk * v[j]; }
            else { for (int j = 0; j < nb; ++j) acc[j] += fk * v[j]; }
        }
        memcpy(out + j0, acc, sizeof(float) * nb);
    }
}

ORC_SIMD void orc_f_fir_shift(const float *restrict padded, const float *f, int fw, int w, float *restrict out)
{
    for (int j0 = 0; j0 < w; j0 += ORC_FBLK) {
        const int nb = w - j0 < ORC_FBLK ? w - j0 : ORC_FBLK;
        float acc[ORC_FBLK] = { 0.0f };
        for (int k = 0; k < fw; ++k) {
            const float *restrict v = padded + j0 + k;
            const float fk = f[k];
            if (nb == ORC_FBLK) { for (int j = 0; j < ORC_FBLK; ++j) acc[j] += fk * v[j]; }
            else { for (int j = 0; j < nb; ++j) acc[j] += fk * v[j]; }
        }
        memcpy(out + j0, acc, sizeof(float) * nb);
    }
}

static void vif_filter1d(const float *f, int fw, const float *src, float *dst, float *tmp /* w + fw */, int w, int h)
{
    const int r = fw / 2;
    for (int i = 0; i < h; ++i) {
        const float *rows[17];
        for (int k = 0; k < fw; ++k) rows[k] = src + (size_t)mirror(i - r + k, h) * w;
        orc_f_fir_rows(rows, f, fw, w, tmp + r);
        for (int m = 1; m <= r; ++m) {                  /* mirror(): -m -> m ; w - 1 + m -> w - m */
            tmp[r - m] = tmp[r + m];
            tmp[r + w - 1 + m] = tmp[r + w - m];
        }
        orc_f_fir_shift(tmp, f, fw, w, dst + (size_t)i * w);
    }
}

/* vif_tools.c log2f_approx() (VIF_OPT_FAST_LOG2): exponent + degree-8 polynomial of the mantissa */
static const float log2_poly[9] = { SPEC_LOG2_POLY };
ORC_API float orc_f_log2_approx(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    int e = (int)((u & 0x7F800000u) >> 23) - 127;
    uint32_t m = (u & 0x007FFFFFu) | 0x3F800000u;
    float r;
    memcpy(&r, &m, 4);
    float t = r - 1.0f, v = 0;
    for (int i = 0; i < 9; ++i) v = v * t + log2_poly[i];
    return (float)e + v;
}

static void vif_statistic(const float *mu1, const float *mu2, const float *xx, const float *yy, const float *xy,
                          int w, int h, double egl, double *num, double *den)
{
    const float sigma_nsq = 2.0f, eps = 1.0e-10f, sigma_max_inv = 4.0f / (255.0f * 255.0f);
    const float gl = (float)egl;
    double an = 0, ad = 0;
    for (int i = 0; i < h; ++i) {
        float rn = 0, rd = 0;
        for (int j = 0; j < w; ++j) {
            size_t p = (size_t)i * w + j;
            float m1 = mu1[p], m2 = mu2[p];
            float s1 = xx[p] - m1 * m1, s2 = yy[p] - m2 * m2, s12 = xy[p] - m1 * m2;
            s1 = s1 < 0 ? 0 : s1;
            s2 = s2 < 0 ? 0 : s2;
            float g = s12 / (s1 + eps);
            float sv = s2 - g * s12;
            if (s1 < eps) { g = 0; sv = s2; s1 = 0; }
            if (s2 < eps) { g = 0; sv = 0; }
            if (g < 0) { sv = s2; g = 0; }
            sv = sv < eps ? eps : sv;
            g = g < gl ? g : gl;
            float nv = orc_f_log2_approx(1.0f + (g * g * s1) / (sv + sigma_nsq));
            float dv = orc_f_log2_approx(1.0f + s1 / sigma_nsq);
            if (s12 < 0) nv = 0;
            if (s1 < sigma_nsq) { nv = 1.0f - s2 * sigma_max_inv; dv = 1.0f; }
            rn += nv; rd += dv;
        }
        an += rn; ad += rd;
    }
    *num = an; *den = ad;
}

/* ref/dis: float pictures (luma - 128), tight pitch.  num/den: 4 scales. */
ORC_API int orc_f_vif(const float *ref, const float *dis, int w, int h, double egl, double num[4], double den[4])
{
    size_t n = (size_t)w * h;
    float *cr = malloc(4 * n), *cd = malloc(4 * n), *mu1 = malloc(4 * n), *mu2 = malloc(4 * n);
    float *a = malloc(4 * n), *b = malloc(4 * n), *c = malloc(4 * n), *t = malloc(4 * n), *tmp = malloc(4 * ((size_t)w + 17));
    memcpy(cr, ref, 4 * n); memcpy(cd, dis, 4 * n);
    for (int s = 0; s < 4; ++s) {
        const float *f = vif_ftab[s];
        int fw = vif_fw[s];
        if (s > 0) {
            vif_filter1d(f, fw, cr, mu1, tmp, w, h);
            vif_filter1d(f, fw, cd, mu2, tmp, w, h);
            int ow = w / 2, oh = h / 2;
            for (int i = 0; i < oh; ++i)
                for (int j = 0; j < ow; ++j) {
                    cr[(size_t)i * ow + j] = mu1[(size_t)(2 * i) * w + 2 * j];
                    cd[(size_t)i * ow + j] = mu2[(size_t)(2 * i) * w + 2 * j];
                }
            w = ow; h = oh;
        }
        size_t m = (size_t)w * h;
        vif_filter1d(f, fw, cr, mu1, tmp, w, h);
        vif_filter1d(f, fw, cd, mu2, tmp, w, h);
        for (size_t p = 0; p < m; ++p) t[p] = cr[p] * cr[p];
        vif_filter1d(f, fw, t, a, tmp, w, h);
        for (size_t p = 0; p < m; ++p) t[p] = cd[p] * cd[p];
        vif_filter1d(f, fw, t, b, tmp, w, h);
        for (size_t p = 0; p < m; ++p) t[p] = cr[p] * cd[p];
        vif_filter1d(f, fw, t, c, tmp, w, h);
        vif_statistic(mu1, mu2, a, b, c, w, h, egl, &num[s], &den[s]);
    }
    free(cr); free(cd); free(mu1); free(mu2); free(a); free(b); free(c); free(t); free(tmp);
    return 0;
}

/* ============================== float motion =========================================== */
static const float motion_f5[5] = { SPEC_MOTION_F32_5 };

ORC_API void orc_f_motion_blur(const float *src, int w, int h, float *dst)
{
    float *tmp = malloc(4 * (size_t)w * h);
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            float acc = 0;
            for (int k = 0; k < 5; ++k) acc += motion_f5[k] * src[(size_t)mirror(i - 2 + k, h) * w + j];
            tmp[(size_t)i * w + j] = acc;
        }
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            float acc = 0;
            for (int k = 0; k < 5; ++k) acc += motion_f5[k] * tmp[(size_t)i * w + mirror(j - 2 + k, w)];
            dst[(size_t)i * w + j] = acc;
        }
    free(tmp);
}

/* motion.c vmaf_image_sad_c(): float row sums, float total, / (w*h) */
ORC_API double orc_f_motion_sad(const float *a, const float *b, int w, int h)
{
    float acc = 0;
    for (int i = 0; i < h; ++i) {
        float line = 0;
        for (int j = 0; j < w; ++j) line += fabsf(a[(size_t)i * w + j] - b[(size_t)i * w + j]);
        acc += line;
    }
    return (double)(float)(acc / (w * h));
}

/* ============================== float ADM ============================================== */
static const float dwt_lo[4] = { SPEC_DWT_LO_F32 };
static const float dwt_hi[4] = { SPEC_DWT_HI_F32 };

typedef struct { float *a, *v, *h, *d; } fbands;

static void fbands_alloc(fbands *b, size_t n) { b->a = malloc(4 * n); b->v = malloc(4 * n); b->h = malloc(4 * n); b->d = malloc(4 * n); }
static void fbands_free(fbands *b) { free(b->a); free(b->v); free(b->h); free(b->d); }

static void adm_dwt2(const float *src, int w, int h, fbands *dst)
{
    const int ow = (w + 1) / 2, oh = (h + 1) / 2;
    float *tlo = malloc(4 * (size_t)w), *thi = malloc(4 * (size_t)w);
    for (int i = 0; i < oh; ++i) {
        int iy[4];
        for (int k = 0; k < 4; ++k) iy[k] = mirror(2 * i - 1 + k, h);
        for (int j = 0; j < w; ++j) {
            float s0 = src[(size_t)iy[0] * w + j], s1 = src[(size_t)iy[1] * w + j];
            float s2 = src[(size_t)iy[2] * w + j], s3 = src[(size_t)iy[3] * w + j];
            float acc = 0;
            acc += dwt_lo[0] * s0; acc += dwt_lo[1] * s1; acc += dwt_lo[2] * s2; acc += dwt_lo[3] * s3;
            tlo[j] = acc;
            acc = 0;
            acc += dwt_hi[0] * s0; acc += dwt_hi[1] * s1; acc += dwt_hi[2] * s2; acc += dwt_hi[3] * s3;
            thi[j] = acc;
        }
        for (int j = 0; j < ow; ++j) {
            int jx[4];
            for (int k = 0; k < 4; ++k) jx[k] = mirror(2 * j - 1 + k, w);
            float acc;
            acc = 0; for (int k = 0; k < 4; ++k) acc += dwt_lo[k] * tlo[jx[k]];
            dst->a[(size_t)i * ow + j] = acc;
            acc = 0; for (int k = 0; k < 4; ++k) acc += dwt_hi[k] * tlo[jx[k]];
            dst->v[(size_t)i * ow + j] = acc;
            acc = 0; for (int k = 0; k < 4; ++k) acc += dwt_lo[k] * thi[jx[k]];
            dst->h[(size_t)i * ow + j] = acc;
            acc = 0; for (int k = 0; k < 4; ++k) acc += dwt_hi[k] * thi[jx[k]];
            dst->d[(size_t)i * ow + j] = acc;
        }
    }
    free(tlo); free(thi);
}

void orc_adm_rfactor(int scale, double view_dist, int display_h, float rf[3]);   /* vmaf_oracle.c */

/* One scale after the DWT: decouple, csf, contrast masking, denominator.  Outputs the cube sums
 * (h, v, d) and the float scale scores. */
static void f_adm_scale(const fbands *ref, const fbands *dis, int w, int h, int scale, double egl, double view_dist,
                        int display_h, double num_sum[3], double den_sum[3], float *num_scale, float *den_scale)
{
    const float cos_1deg_sq = cos(1.0 * M_PI / 180.0) * cos(1.0 * M_PI / 180.0);
    const float eps = 1e-30f, one_by_30 = 0.0333333351f, one_by_15 = 0.0666666701f;
    float rf[3];
    orc_adm_rfactor(scale, view_dist, display_h, rf);
    int left = w * 0.1 - 0.5, top = h * 0.1 - 0.5;
    int right = w - left, bottom = h - top;
    int gl = left - 1 < 0 ? 0 : left - 1, gt = top - 1 < 0 ? 0 : top - 1;
    int gr = right + 1 > w ? w : right + 1, gb = bottom + 1 > h ? h : bottom + 1;
    const size_t n = (size_t)w * h;
    float *R[3], *CA[3], *CF[3];
    for (int b = 0; b < 3; ++b) { R[b] = calloc(n, 4); CA[b] = calloc(n, 4); CF[b] = calloc(n, 4); }
    const float *O[3] = { ref->h, ref->v, ref->d }, *T[3] = { dis->h, dis->v, dis->d };
    for (int i = gt; i < gb; ++i)
        for (int j = gl; j < gr; ++j) {
            size_t p = (size_t)i * w + j;
            float oh = O[0][p], ov = O[1][p], th = T[0][p], tv = T[1][p];
            float ot_dp = oh * th + ov * tv, o_mag = oh * oh + ov * ov, t_mag = th * th + tv * tv;
            int flag = (ot_dp >= 0.0f) && (ot_dp * ot_dp >= cos_1deg_sq * o_mag * t_mag);
            for (int b = 0; b < 3; ++b) {
                float o = O[b][p], t = T[b][p];
                float k = t / (o + eps);
                k = k < 0.0f ? 0.0f : (k > 1.0f ? 1.0f : k);
                float rst = k * o;
                if (flag) {
                    if (rst > 0.) { double v = rst * egl; rst = (float)(v < t ? v : t); }
                    else if (rst < 0.) { double v = rst * egl; rst = (float)(v > t ? v : t); }
                }
                R[b][p] = rst;
                float ca = rf[b] * (t - rst);
                CA[b][p] = ca;
                CF[b][p] = one_by_30 * fabsf(ca);
            }
        }
    float acc_n[3] = { 0, 0, 0 }, acc_d[3] = { 0, 0, 0 };
    for (int i = top; i < bottom; ++i) {
        float in_n[3] = { 0, 0, 0 }, in_d[3] = { 0, 0, 0 };
        for (int j = left; j < right; ++j) {
            size_t p = (size_t)i * w + j;
            float thr = 0;
            for (int b = 0; b < 3; ++b) {
                float sum = 0;
                for (int di = -1; di <= 1; ++di)
                    for (int dj = -1; dj <= 1; ++dj) {
                        size_t q = (size_t)mirror(i + di, h) * w + mirror(j + dj, w);
                        if (di == 0 && dj == 0) sum += one_by_15 * fabsf(CA[b][q]);
                        else sum += CF[b][q];
                    }
                thr += sum;
            }
            for (int b = 0; b < 3; ++b) {
                float x = fabsf(R[b][p] * rf[b]) - thr;
                x = x < 0.0f ? 0.0f : x;
                in_n[b] += x * x * x;
                float v = fabsf(O[b][p]) * rf[b];
                in_d[b] += v * v * v;
            }
        }
        for (int b = 0; b < 3; ++b) { acc_n[b] += in_n[b]; acc_d[b] += in_d[b]; }
    }
    const float area = powf((bottom - top) * (right - left) / 32.0f, 1.0f / 3.0f);
    float ns = 0, ds = 0;
    for (int b = 0; b < 3; ++b) {
        num_sum[b] = acc_n[b]; den_sum[b] = acc_d[b];
        ns += powf(acc_n[b], 1.0f / 3.0f) + area;
        ds += powf(acc_d[b], 1.0f / 3.0f) + area;
    }
    *num_scale = ns; *den_scale = ds;
    for (int b = 0; b < 3; ++b) { free(R[b]); free(CA[b]); free(CF[b]); }
}

ORC_API int orc_f_adm(const float *ref, const float *dis, int w, int h, double egl, double view_dist, int display_h,
                      double num_sum[4][3], double den_sum[4][3], double num_scale[4], double den_scale[4], double *adm2)
{
    const double limit = 1e-10 * (w * h) / (1920.0 * 1080.0);
    float *cr = malloc(4 * (size_t)w * h), *cd = malloc(4 * (size_t)w * h);
    memcpy(cr, ref, 4 * (size_t)w * h); memcpy(cd, dis, 4 * (size_t)w * h);
    double num = 0, den = 0;
    for (int s = 0; s < 4; ++s) {
        fbands rb, db;
        int ow = (w + 1) / 2, oh = (h + 1) / 2;
        fbands_alloc(&rb, (size_t)ow * oh); fbands_alloc(&db, (size_t)ow * oh);
        adm_dwt2(cr, w, h, &rb); adm_dwt2(cd, w, h, &db);
        w = ow; h = oh;
        float ns, ds;
        f_adm_scale(&rb, &db, w, h, s, egl, view_dist, display_h, num_sum[s], den_sum[s], &ns, &ds);
        num_scale[s] = ns; den_scale[s] = ds;
        num += ns; den += ds;
        memcpy(cr, rb.a, 4 * (size_t)w * h); memcpy(cd, db.a, 4 * (size_t)w * h);
        fbands_free(&rb); fbands_free(&db);
    }
    free(cr); free(cd);
    num = num < limit ? 0 : num;
    den = den < limit ? 0 : den;
    *adm2 = den == 0.0 ? 1.0 : num / den;
    return 0;
}

/* ============================== SSIM / MS-SSIM (iqa) =================================== */
static const float g_gauss11[11] = { SPEC_SSIM_GAUSS11 };
static const float g_lpf9[9] = { SPEC_MS_SSIM_LPF9 };

static inline int sym(int i, int n)            /* KBND_SYMMETRIC: -1 -> 0 ; n -> n-1 */
{
    if (i < 0) return -1 - i;
    if (i >= n) return 2 * n - i - 1;
    return i;
}

/* _iqa_decimate with an f x f box kernel (even f: taps x-f/2 .. x+f/2-1), symmetric borders */
static float *decimate_box(const float *img, int w, int h, int f, int *ow, int *oh)
{
    int dw = w / f + (w & 1), dh = h / f + (h & 1);
    float *dst = malloc(4 * (size_t)dw * dh);
    const float kv = 1.0f / (float)(f * f);
    const int c = f / 2;
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            float sum = 0;
            for (int v = 0; v < f; ++v)
                for (int u = 0; u < f; ++u)
                    sum += img[(size_t)sym(y * f - c + v, h) * w + sym(x * f - c + u, w)] * kv;
            dst[(size_t)y * dw + x] = sum;
        }
    *ow = dw; *oh = dh;
    return dst;
}

/* _iqa_decimate by 2 with the separable 9-tap low-pass (horizontal then vertical), symmetric borders.  The horizontal pass
 * reads a row copy that carries its 4-sample symmetric border (no index arithmetic per tap); the vertical pass is the
 * row-wise FIR helper.  Per output still sum = 0, sum += sample * g[u] for u = 0 .. 8 in float. */
static float *decimate_lpf2(const float *img, int w, int h, int *ow, int *oh)
{
    int dw = w / 2 + (w & 1), dh = h / 2 + (h & 1);
    float *tmp = malloc(4 * (size_t)dw * h);
    float *row = malloc(4 * ((size_t)w + 16));
    for (int y = 0; y < h; ++y) {
        const float *src = img + (size_t)y * w;
        memcpy(row + 4, src, 4 * (size_t)w);
        for (int m = 1; m <= 4; ++m) row[4 - m] = src[sym(-m, w)];
        for (int m = 0; m < 8; ++m) row[4 + w + m] = src[sym(w + m, w)];
        for (int x = 0; x < dw; ++x) {
            const float *q = row + 2 * x;
            float sum = 0;
            for (int u = 0; u < 9; ++u) sum += q[u] * g_lpf9[u];
            tmp[(size_t)y * dw + x] = sum;
        }
    }
    free(row);
    float *dst = malloc(4 * (size_t)dw * dh);
    for (int y = 0; y < dh; ++y) {
        const float *rows[9];
        for (int v = 0; v < 9; ++v) rows[v] = tmp + (size_t)sym(2 * y - 4 + v, h) * dw;
        orc_f_fir_rows(rows, g_lpf9, 9, dw, dst + (size_t)y * dw);
    }
    free(tmp);
    *ow = dw; *oh = dh;
    return dst;
}

/* valid 11x11 separable Gaussian (horizontal then vertical); the row-wise FIR helpers of the VIF section: every output is
 * still s = 0, s += sample * g[u] for u = 0 .. 10 in float (multiplication commutes, so the operand order is immaterial) */
static void gauss_valid(const float *img, int w, int h, float *dst)
{
    int vw = w - 10, vh = h - 10;
    float *tmp = malloc(4 * (size_t)vw * h);
    for (int y = 0; y < h; ++y) orc_f_fir_shift(img + (size_t)y * w, g_gauss11, 11, vw, tmp + (size_t)y * vw);
    for (int y = 0; y < vh; ++y) {
        const float *rows[11];
        for (int v = 0; v < 11; ++v) rows[v] = tmp + (size_t)(y + v) * vw;
        orc_f_fir_rows(rows, g_gauss11, 11, vw, dst + (size_t)y * vw);
    }
    free(tmp);
}

/* _iqa_ssim(): sums of the ssim, l, c, s maps over the valid region.  sums[4] = ssim, l, c, s */
static void ssim_maps(const float *ref, const float *cmp, int w, int h, double sums[4], int *count)
{
    const float C1 = (SPEC_SSIM_K1 * 255.0f) * (SPEC_SSIM_K1 * 255.0f), C2 = (SPEC_SSIM_K2 * 255.0f) * (SPEC_SSIM_K2 * 255.0f), C3 = C2 / 2.0f;
    const int vw = w - 10, vh = h - 10;
    const size_t n = (size_t)w * h, vn = (size_t)vw * vh;
    float *t = malloc(4 * n), *mu1 = malloc(4 * vn), *mu2 = malloc(4 * vn), *s1 = malloc(4 * vn), *s2 = malloc(4 * vn),
          *s12 = malloc(4 * vn);
    gauss_valid(ref, w, h, mu1);
    gauss_valid(cmp, w, h, mu2);
    for (size_t p = 0; p < n; ++p) t[p] = ref[p] * ref[p];
    gauss_valid(t, w, h, s1);
    for (size_t p = 0; p < n; ++p) t[p] = cmp[p] * cmp[p];
    gauss_valid(t, w, h, s2);
    for (size_t p = 0; p < n; ++p) t[p] = ref[p] * cmp[p];
    gauss_valid(t, w, h, s12);
    double a = 0, l = 0, c = 0, s = 0;
    for (size_t p = 0; p < vn; ++p) {
        float v1 = s1[p] - mu1[p] * mu1[p], v2 = s2[p] - mu2[p] * mu2[p], cv = s12[p] - mu1[p] * mu2[p];
        v1 = v1 < 0.0f ? 0.0f : v1;
        v2 = v2 < 0.0f ? 0.0f : v2;
        double sr = sqrt((double)v1 * v2);
        double lv = (2.0 * mu1[p] * mu2[p] + C1) / ((double)mu1[p] * mu1[p] + (double)mu2[p] * mu2[p] + C1);
        double cc = (2.0 * sr + C2) / ((double)v1 + v2 + C2);
        double sv = ((double)cv + C3) / (sr + C3);
        a += lv * cc * sv; l += lv; c += cc; s += sv;
    }
    sums[0] = a; sums[1] = l; sums[2] = c; sums[3] = s;
    *count = (int)vn;
    free(t); free(mu1); free(mu2); free(s1); free(s2); free(s12);
}

/* float_ssim: ref/cmp are luma as float in [0, 255] */
ORC_API double orc_f_ssim(const float *ref, const float *cmp, int w, int h)
{
    int mn = w < h ? w : h;
    int f = (int)lroundf((float)mn / 256.0f);
    if (f < 1) f = 1;
    float *r = (float *)ref, *c = (float *)cmp;
    int cw = w, ch = h;
    if (f > 1) { r = decimate_box(ref, w, h, f, &cw, &ch); c = decimate_box(cmp, w, h, f, &cw, &ch); }
    double sums[4];
    int cnt;
    ssim_maps(r, c, cw, ch, sums, &cnt);
    if (f > 1) { free(r); free(c); }
    return sums[0] / cnt;
}

/* float_ms_ssim: 5 scales; per scale the MEANS of l, c, s; score = prod l^a * c^b * s^g.
 * lcs[15] (optional) receives the per-scale means. */
ORC_API double orc_f_ms_ssim(const float *ref, const float *cmp, int w, int h, double *lcs)
{
    static const double alphas[5] = { 0.0, 0.0, 0.0, 0.0, 0.1333 };
    static const double betas[5] = { SPEC_MS_SSIM_EXPONENTS };
    float *r = malloc(4 * (size_t)w * h), *c = malloc(4 * (size_t)w * h);
    memcpy(r, ref, 4 * (size_t)w * h); memcpy(c, cmp, 4 * (size_t)w * h);
    double score = 1.0;
    for (int s = 0; s < 5; ++s) {
        if (s > 0) {
            int nw, nh;
            float *nr = decimate_lpf2(r, w, h, &nw, &nh), *nc = decimate_lpf2(c, w, h, &nw, &nh);
            free(r); free(c);
            r = nr; c = nc; w = nw; h = nh;
        }
        double sums[4];
        int cnt;
        ssim_maps(r, c, w, h, sums, &cnt);
        double l = sums[1] / cnt, cc = sums[2] / cnt, sv = sums[3] / cnt;
        if (lcs) { lcs[3 * s] = l; lcs[3 * s + 1] = cc; lcs[3 * s + 2] = sv; }
        score *= pow(l, alphas[s]) * pow(cc, betas[s]) * pow(sv, betas[s]);
    }
    free(r); free(c);
    return score;
}
