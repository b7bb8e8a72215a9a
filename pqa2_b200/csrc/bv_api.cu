// libb200vmaf C ABI (include/b200vmaf.h): contexts, the frame-group pipeline, scalar finalisation.
//
// Replaces the process boundary of the reference (app/vmaf_analyzer.py:411-455 spawns ffmpeg and
// :640-641 reads libvmaf's JSON log back).  One bv_ctx drives one GPU:
//
//   upload stream : cudaMemcpy2DAsync of the submitted (pinned) YUV planes into a ring of frame groups
//   compute stream: per group ONE launch of every kernel with blockIdx.z = frame (small pyramid
//                   levels of a single frame cannot fill 148 SMs; a group of B frames can)
//   results       : 64 raw 64-bit accumulators (+ float sums) per frame, D2H into pinned memory,
//                   finalised by scalar host code in bv_fetch()
//
// There is no CPU fallback: every entry point that computes anything fails with BV_ERR_CUDA when no
// device is usable.
#include "bv_common.cuh"
#include "bv_float.cuh"
#include "../../include/b200vmaf.h"

#include <math.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

int bv_vif_fuse_mask();         // bv_vif.cu

namespace {

thread_local std::string g_create_error;

constexpr int N_SLOTS = 3;          // frame groups in flight (upload / compute / readback)

const char *const k_kernel_names[BV_MAX_KERNELS] = {
    "motion_blur", "motion_sad",
    "vif_stat_s0", "vif_subsample_s1", "vif_stat_s1", "vif_subsample_s2", "vif_stat_s2", "vif_subsample_s3", "vif_stat_s3",
    "adm_scale0", "adm_scale1", "adm_scale2", "adm_scale3", "adm_rows_finish",
    "psnr_sse_y", "psnr_sse_u", "psnr_sse_v",
    "ffssim_y", "ffssim_u", "ffssim_v",
};

struct Group {
    int n = 0;
    int64_t first_ordinal = 0;
    int64_t frame_index[BV_MAX_BATCH];
    unsigned flags[BV_MAX_BATCH];
    // device copies of host-submitted frames: [frame][clip 0 ref / 1 dis][plane]
    uint8_t *d_planes[3][2] = { { nullptr, nullptr }, { nullptr, nullptr }, { nullptr, nullptr } };
    // per-frame plane pointers for this group's launch (device staging or caller's device memory)
    const uint8_t *p[BV_MAX_BATCH][2][3];
    size_t pitch[2][3];
    bool pitch_set = false;
    bool uses_staging = false;
    unsigned long long *d_raw = nullptr, *h_raw = nullptr;     // [B][BV_RAW_WORDS]
    double *d_fraw = nullptr, *h_fraw = nullptr;               // [B][BV_FRAW_WORDS]
    cudaEvent_t uploaded = nullptr, done = nullptr;
    BvProf prof;
    bool in_flight = false;
};

}  // namespace

struct bv_ctx {
    int device = 0, w = 0, h = 0, bpc = 8, chroma = 420;
    int cw = 0, ch = 0;                 // chroma plane dims
    unsigned feat = 0;
    bv_opts opts;
    int B = BV_MAX_BATCH;
    cudaStream_t up = nullptr, comp = nullptr;
    Group groups[N_SLOTS];
    int cur = 0;                        // group being filled
    int64_t submitted = 0;
    std::vector<bv_frame_features> results;
    std::atomic<int64_t> ready{ 0 };
    std::atomic<int> cancelled{ 0 };
    std::string err;
    long long nlaunch = 0;
    bool profiling = false;
    double k_ms[BV_MAX_KERNELS] = {};
    double k_count[BV_MAX_KERNELS] = {};
    cudaEvent_t t_ev[2] = { nullptr, nullptr };      // bv_timer_mark
    bool t_set[2] = { false, false };
    size_t staging_pitch[3] = { 0, 0, 0 };
    size_t staging_frame_bytes[3] = { 0, 0, 0 };

    // motion state: two blur arrays of B frames, used alternately by consecutive groups
    uint16_t *blur[2] = { nullptr, nullptr };
    int blur_cur = 0, blur_prev_n = 0;
    size_t blur_elems = 0;
    // vif
    BvVifLevels vif_lv;
    uint16_t *d_log2 = nullptr;
    uint8_t *d_log2c = nullptr;         // compressed log2 table (see alloc_ctx)
    // adm
    BvAdmBuffers adm;
    BvAdmScaleParams adm_sp[4];
    int *d_div = nullptr;
    // float extractors
    BvFloatState *fl = nullptr;
    BvFloatStateFast *flf = nullptr;    // opts.fast_float: the contracted / folded-tap build of the same kernels
};

namespace {

int fail(bv_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                             \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) return fail(c, BV_ERR_CUDA, #call, e_);       \
    } while (0)

size_t sample_bytes(int bpc) { return bpc > 8 ? 2 : 1; }

void plane_dims(const bv_ctx *c, int plane, int *pw, int *ph)
{
    if (plane == 0) { *pw = c->w; *ph = c->h; }
    else { *pw = c->cw; *ph = c->ch; }
}

bool needs_chroma(const bv_ctx *c) { return (c->feat & (BV_FEAT_PSNR_UV | BV_FEAT_FFSSIM)) && c->chroma != 0 && c->chroma != 400; }

int alloc_ctx(bv_ctx *c)
{
    CK(cudaSetDevice(c->device));
    CK(cudaStreamCreateWithFlags(&c->up, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->comp, cudaStreamNonBlocking));
    const int B = c->B;
    for (int s = 0; s < N_SLOTS; ++s) {
        Group &g = c->groups[s];
        CK(cudaMalloc(&g.d_raw, sizeof(unsigned long long) * B * BV_RAW_WORDS));
        CK(cudaHostAlloc(&g.h_raw, sizeof(unsigned long long) * B * BV_RAW_WORDS, cudaHostAllocDefault));
        CK(cudaMalloc(&g.d_fraw, sizeof(double) * B * BV_FRAW_WORDS));
        CK(cudaHostAlloc(&g.h_fraw, sizeof(double) * B * BV_FRAW_WORDS, cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&g.uploaded, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
        memset(&g.prof, 0, sizeof g.prof);
    }
    for (int k = 0; k < 2; ++k) CK(cudaEventCreate(&c->t_ev[k]));
    const size_t npx = (size_t)c->w * c->h;
    if (c->feat & BV_FEAT_MOTION) {
        c->blur_elems = (npx + 7) & ~(size_t)7;      // frames stay 16-byte aligned for the uint4 SAD loads
        for (int k = 0; k < 2; ++k) CK(cudaMalloc(&c->blur[k], sizeof(uint16_t) * c->blur_elems * B));
    }
    if (c->feat & BV_FEAT_VIF) {
        std::vector<uint16_t> tab(65536, 0);
        for (unsigned i = 32767; i < 65536; ++i) tab[i] = (uint16_t)round(log2f((float)i) * 2048);
        CK(cudaMalloc(&c->d_log2, sizeof(uint16_t) * 65536));
        CK(cudaMemcpy(c->d_log2, tab.data(), sizeof(uint16_t) * 65536, cudaMemcpyHostToDevice));
        // Compressed copy for shared memory (the table is monotone with slope < 0.1 per entry): 512 u16 bases,
        // one per 64 entries, followed by 32768 4-bit deltas -> 17 KB instead of 64 KB.  Exact by construction.
        std::vector<uint8_t> packed(BV_LOG2C_BYTES, 0);
        uint16_t *base = reinterpret_cast<uint16_t *>(packed.data());
        for (unsigned j = 0; j < 32768; ++j) {
            const unsigned i = 32768 + j, b = tab[i & ~63u], d = tab[i] - b;
            if (d > 15u) return fail(c, BV_ERR_UNSUPPORTED, "log2 table does not compress to 4-bit deltas");
            base[j >> 6] = (uint16_t)b;
            packed[1024 + (j >> 1)] |= (uint8_t)(d << ((j & 1) * 4));
        }
        CK(cudaMalloc(&c->d_log2c, BV_LOG2C_BYTES));
        CK(cudaMemcpy(c->d_log2c, packed.data(), BV_LOG2C_BYTES, cudaMemcpyHostToDevice));
        int lw = c->w, lh = c->h;
        c->vif_lv.w[0] = lw; c->vif_lv.h[0] = lh;
        c->vif_lv.ref[0] = c->vif_lv.dis[0] = nullptr; c->vif_lv.frame_elems[0] = 0;
        for (int s = 1; s < 4; ++s) {
            lw /= 2; lh /= 2;
            c->vif_lv.w[s] = lw; c->vif_lv.h[s] = lh;
            c->vif_lv.frame_elems[s] = (size_t)lw * lh;
            CK(cudaMalloc(&c->vif_lv.ref[s], sizeof(uint16_t) * (size_t)lw * lh * B));
            CK(cudaMalloc(&c->vif_lv.dis[s], sizeof(uint16_t) * (size_t)lw * lh * B));
        }
    }
    if (c->feat & BV_FEAT_ADM) {
        std::vector<int> div(65537, 0);
        for (int i = 1; i <= 32768; ++i) {
            const int recip = (int)(1073741824 / i);
            div[32768 + i] = recip;
            div[32768 - i] = 0 - recip;
        }
        CK(cudaMalloc(&c->d_div, sizeof(int) * 65537));
        CK(cudaMemcpy(c->d_div, div.data(), sizeof(int) * 65537, cudaMemcpyHostToDevice));
        bv_adm_make_params(c->w, c->h, c->opts.adm_norm_view_dist, c->opts.adm_ref_display_height, c->adm_sp);
        size_t rows = 0;
        for (int s = 0; s < 4; ++s) {
            const size_t n = (size_t)c->adm_sp[s].w * c->adm_sp[s].h;
            c->adm.band_plane_elems[s] = n;
            c->adm.bands[s] = nullptr;
            if (s < 3) CK(cudaMalloc(&c->adm.bands[s], (s == 0 ? 2 : 4) * n * 2 * BV_MAX_BATCH));
            c->adm.rows_scale_offset[s] = rows;
            rows += (size_t)c->adm_sp[s].h * 6;
        }
        c->adm.rows_frame_stride = rows;
        CK(cudaMalloc(&c->adm.rows, sizeof(unsigned long long) * rows * B));
        c->adm.div_lookup = c->d_div;
    }
    if (c->feat & (BV_FEAT_VMAF_FLOAT | BV_FEAT_FLOAT_SSIM | BV_FEAT_FLOAT_MS_SSIM)) {
        if (c->opts.fast_float) c->flf = bv_float_fast_create(c->w, c->h, c->bpc, c->feat, B, &c->opts);
        else c->fl = bv_float_create(c->w, c->h, c->bpc, c->feat, B, &c->opts);
        if (!c->fl && !c->flf) return fail(c, BV_ERR_CUDA, "bv_float_create failed (out of device memory?)");
    }
    return 0;
}

int ensure_staging(bv_ctx *c, Group &g)
{
    if (g.d_planes[0][0]) return 0;
    const int np = needs_chroma(c) ? 3 : 1;
    for (int p = 0; p < np; ++p) {
        int pw, ph;
        plane_dims(c, p, &pw, &ph);
        const size_t pitch = (((size_t)pw * sample_bytes(c->bpc)) + 255) & ~(size_t)255;
        c->staging_pitch[p] = pitch;
        c->staging_frame_bytes[p] = pitch * ph;
        for (int clip = 0; clip < 2; ++clip)
            CK(cudaMalloc(&g.d_planes[p][clip], pitch * ph * c->B));
    }
    return 0;
}

// ---- scalar finalisation (host) ---------------------------------------------------------------
void finish_frame(bv_ctx *c, const Group &g, int f, bv_frame_features &o)
{
    memset(&o, 0, sizeof o);
    o.frame_index = g.frame_index[f];
    o.flags = g.flags[f];
    const unsigned long long *raw = g.h_raw + (size_t)f * BV_RAW_WORDS;
    for (int k = 0; k < BV_RAW_WORDS; ++k) o.raw[k] = (int64_t)raw[k];
    const bool spatial = !(g.flags[f] & (BV_FRAME_LEAD_IN | BV_FRAME_SKIP_SPATIAL));
    const bool lead = g.flags[f] & BV_FRAME_LEAD_IN;
    unsigned valid = 0;
    if ((c->feat & BV_FEAT_MOTION) && !lead) {
        // libvmaf integer_motion.c normalize_and_scale_sad(): (float)(sad / 256.) / (w * h)
        o.motion = (double)((float)((double)raw[BV_RAW_SAD] / 256.) / (float)((unsigned)c->w * (unsigned)c->h));
        valid |= BV_FEAT_MOTION;
    }
    if ((c->feat & BV_FEAT_VIF) && spatial) {
        for (int s = 0; s < 4; ++s) {
            const int64_t *a = o.raw + BV_RAW_VIF + 7 * s;
            // libvmaf integer_vif.c: the scale sums are stored through `float num, den`
            const float n = (float)(a[0] / 2048.0 + (double)a[5] + ((double)a[3] - ((double)a[2] / 16384.0) / 65025.0));
            const float d = (float)(a[1] / 2048.0 - ((double)a[4] + (double)(a[6] * 17)) + (double)a[3]);
            o.vif_num[s] = n; o.vif_den[s] = d;
            o.vif_scale[s] = (double)(n / d);
        }
        valid |= BV_FEAT_VIF;
    }
    if ((c->feat & BV_FEAT_ADM) && spatial) {
        double num = 0, den = 0;
        for (int s = 0; s < 4; ++s) {
            float ns, ds;
            bv_adm_finish_scale(c->adm_sp[s], s, c->opts.adm_norm_view_dist, c->opts.adm_ref_display_height,
                                o.raw + BV_RAW_ADM_CM + 3 * s, (const uint64_t *)(o.raw + BV_RAW_ADM_DEN + 3 * s),
                                &ns, &ds);
            o.adm_num[s] = ns; o.adm_den[s] = ds;
            o.adm_scale[s] = (double)ns / (double)ds;
            num += ns; den += ds;
        }
        const double limit = 1e-10 * ((double)c->w * c->h) / (1920.0 * 1080.0);
        num = num < limit ? 0 : num;
        den = den < limit ? 0 : den;
        o.adm2 = den == 0.0 ? 1.0 : num / den;
        valid |= BV_FEAT_ADM;
    }
    if (spatial) {
        const double peak = (double)((1 << c->bpc) - 1), pmax = 6.0 * c->bpc + 12.0;
        double *dst[3] = { &o.psnr_y, &o.psnr_cb, &o.psnr_cr };
        for (int p = 0; p < 3; ++p) {
            const unsigned need = p == 0 ? BV_FEAT_PSNR_Y : BV_FEAT_PSNR_UV;
            if (!(c->feat & need) || (p > 0 && !needs_chroma(c))) continue;
            int pw, ph;
            plane_dims(c, p, &pw, &ph);
            const uint64_t sse = raw[BV_RAW_SSE + p];
            if (sse == 0) *dst[p] = pmax;
            else {
                const double mse = (double)sse / ((double)pw * ph);
                const double v = 10.0 * log10(peak * peak / mse);
                *dst[p] = v < pmax ? v : pmax;
            }
            valid |= need;
        }
    }
    if ((c->feat & BV_FEAT_FFSSIM) && spatial) {
        // vf_ssim.c ssim_plane(): mean over the (W/4 - 1) x (H/4 - 1) overlapped 8x8 windows
        const int np = needs_chroma(c) ? 3 : 1;
        for (int p = 0; p < np; ++p) {
            int pw, ph;
            plane_dims(c, p, &pw, &ph);
            const int W4 = pw >> 2, H4 = ph >> 2;
            o.ffssim[p] = (W4 < 2 || H4 < 2) ? 1.0
                : g.h_fraw[(size_t)f * BV_FRAW_WORDS + BV_FRAW_FFSSIM + p] / ((double)(H4 - 1) * (W4 - 1));
        }
        valid |= BV_FEAT_FFSSIM;
    }
    if (c->fl) valid |= bv_float_finish(c->fl, g.h_fraw + (size_t)f * BV_FRAW_WORDS, g.flags[f], &o);
    if (c->flf) valid |= bv_float_fast_finish(c->flf, g.h_fraw + (size_t)f * BV_FRAW_WORDS, g.flags[f], &o);
    o.valid_mask = valid;
}

int harvest(bv_ctx *c, Group &g)
{
    if (!g.in_flight) return 0;
    CK(cudaEventSynchronize(g.done));
    if (g.prof.on) {
        for (int k = 0; k < BV_MAX_KERNELS; ++k) {
            if (!g.prof.used[k]) continue;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, g.prof.ev[k][0], g.prof.ev[k][1]) == cudaSuccess) {
                c->k_ms[k] += ms;
                c->k_count[k] += 1.0;
            }
            g.prof.used[k] = false;
        }
        g.prof.on = false;
    }
    if ((int64_t)c->results.size() < g.first_ordinal + g.n) c->results.resize(g.first_ordinal + g.n);
    for (int f = 0; f < g.n; ++f) finish_frame(c, g, f, c->results[g.first_ordinal + f]);
    c->ready.store(g.first_ordinal + g.n);
    g.in_flight = false;
    g.n = 0;
    g.pitch_set = false;
    g.uses_staging = false;
    return 0;
}

BvPlane group_plane(const Group &g, int clip, int plane)
{
    BvPlane pl;
    for (int f = 0; f < BV_MAX_BATCH; ++f) pl.p[f] = f < g.n ? g.p[f][clip][plane] : nullptr;
    pl.pitch = g.pitch[clip][plane];
    return pl;
}

int launch_group(bv_ctx *c, Group &g)
{
    if (g.n == 0) return 0;
    CK(cudaSetDevice(c->device));
    if (g.uses_staging) {
        CK(cudaEventRecord(g.uploaded, c->up));
        CK(cudaStreamWaitEvent(c->comp, g.uploaded, 0));
    }
    cudaStream_t st = c->comp;
    BvBatch b;
    b.n = g.n;
    for (int f = 0; f < BV_MAX_BATCH; ++f) b.flags[f] = f < g.n ? g.flags[f] : 0u;
    CK(cudaMemsetAsync(g.d_raw, 0, sizeof(unsigned long long) * g.n * BV_RAW_WORDS, st));
    CK(cudaMemsetAsync(g.d_fraw, 0, sizeof(double) * g.n * BV_FRAW_WORDS, st));
    const BvPlane ry = group_plane(g, 0, 0), dy = group_plane(g, 1, 0);
    if (c->profiling && !g.prof.have_events) {
        for (int k = 0; k < BV_MAX_KERNELS; ++k) { CK(cudaEventCreate(&g.prof.ev[k][0])); CK(cudaEventCreate(&g.prof.ev[k][1])); }
        g.prof.have_events = true;
    }
    g.prof.on = c->profiling;
    const BvLaunch L = { st, &g.prof, &c->nlaunch };

    {
        // integer motion + VIF: with both enabled the blur rides on the VIF scale-0 kernel's staged tile; the SAD follows
        uint16_t *cur = (c->feat & BV_FEAT_MOTION) ? c->blur[c->blur_cur] : nullptr;
        const bool fused_blur = cur && (c->feat & BV_FEAT_VIF) && (bv_vif_fuse_mask() & 2);
        if (cur && !fused_blur) bv_launch_motion_blur(b, ry, c->bpc, c->w, c->h, cur, c->blur_elems, L);
        if (c->feat & BV_FEAT_VIF)
            bv_launch_vif(b, ry, dy, c->bpc, c->vif_lv, c->d_log2, c->d_log2c, c->opts.vif_enhn_gain_limit, g.d_raw, L,
                          fused_blur ? cur : nullptr, c->blur_elems);
        if (cur) {
            const uint16_t *prev_last = c->blur_prev_n > 0
                ? c->blur[c->blur_cur ^ 1] + (size_t)(c->blur_prev_n - 1) * c->blur_elems : cur;
            bv_launch_motion_sad(b, cur, prev_last, c->blur_elems, c->w, c->h, g.d_raw, L);
            c->blur_prev_n = g.n;
            c->blur_cur ^= 1;
        }
    }
    if (c->feat & BV_FEAT_ADM) {
        CK(cudaMemsetAsync(c->adm.rows, 0, sizeof(unsigned long long) * c->adm.rows_frame_stride * g.n, st));
        bv_launch_adm(b, ry, dy, c->bpc, c->adm, c->adm_sp, c->opts.adm_enhn_gain_limit, g.d_raw, L);
    }
    if (c->feat & BV_FEAT_PSNR_Y)
        bv_launch_sse(b, ry, dy, c->bpc, c->w, c->h, 0, g.d_raw, L);
    if ((c->feat & BV_FEAT_PSNR_UV) && needs_chroma(c))
        for (int p = 1; p < 3; ++p)
            bv_launch_sse(b, group_plane(g, 0, p), group_plane(g, 1, p), c->bpc, c->cw, c->ch, p, g.d_raw, L);
    if (c->feat & BV_FEAT_FFSSIM) {
        const int np = needs_chroma(c) ? 3 : 1;
        for (int p = 0; p < np; ++p) {
            int pw, ph;
            plane_dims(c, p, &pw, &ph);
            bv_launch_ffssim(b, group_plane(g, 0, p), group_plane(g, 1, p), c->bpc, pw, ph, p, g.d_fraw,
                             BV_FRAW_FFSSIM + p, BV_FRAW_WORDS, L);
        }
    }
    if (c->fl) bv_float_launch(c->fl, b, ry, dy, g.d_fraw, L);
    if (c->flf) bv_float_fast_launch(c->flf, b, ry, dy, g.d_fraw, L);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(g.h_raw, g.d_raw, sizeof(unsigned long long) * g.n * BV_RAW_WORDS, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(g.h_fraw, g.d_fraw, sizeof(double) * g.n * BV_FRAW_WORDS, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(g.done, st));
    g.in_flight = true;
    c->cur = (c->cur + 1) % N_SLOTS;
    return 0;
}

int submit_common(bv_ctx *c, int64_t frame_index, const void *const rp[3], const size_t rs[3],
                  const void *const dp[3], const size_t ds[3], unsigned flags, bool from_host)
{
    if (!c) return BV_ERR_ARG;
    if (c->cancelled.load()) return fail(c, BV_ERR_CANCELLED, "cancelled");
    if (!rp || !dp || !rs || !ds || !rp[0] || !dp[0]) return fail(c, BV_ERR_ARG, "bv_submit: null luma plane");
    const int np = needs_chroma(c) ? 3 : 1;
    for (int p = 1; p < np; ++p)
        if (!rp[p] || !dp[p]) return fail(c, BV_ERR_ARG, "bv_submit: chroma feature enabled but chroma plane is null");
    CK(cudaSetDevice(c->device));
    Group *g = &c->groups[c->cur];
    if (g->in_flight) { int rc = harvest(c, *g); if (rc) return rc; }
    if (g->n == 0) {
        g->first_ordinal = c->submitted;
        g->uses_staging = false;
        g->pitch_set = false;
    }
    const int f = g->n;
    if (from_host) {
        int rc = ensure_staging(c, *g);
        if (rc) return rc;
        g->uses_staging = true;
        for (int clip = 0; clip < 2; ++clip) {
            const void *const *src = clip == 0 ? rp : dp;
            const size_t *stride = clip == 0 ? rs : ds;
            for (int p = 0; p < np; ++p) {
                int pw, ph;
                plane_dims(c, p, &pw, &ph);
                uint8_t *dst = g->d_planes[p][clip] + (size_t)f * c->staging_frame_bytes[p];
                const size_t row_bytes = (size_t)pw * sample_bytes(c->bpc);
                // A 2-D copy of 1080 rows of 1920 B runs the DMA engine at well under half of PCIe
                // speed; tightly packed host planes go over as ONE linear copy into a tight device pitch
                // (the kernels take any pitch; 16-byte row alignment only matters for the vector paths).
                const bool tight = stride[p] == row_bytes && (row_bytes % 16) == 0;
                const size_t dpitch = tight ? row_bytes : c->staging_pitch[p];
                if (g->pitch_set && g->pitch[clip][p] != dpitch)
                    return fail(c, BV_ERR_ARG, "bv_submit: plane strides must not change within a frame group");
                if (tight) CK(cudaMemcpyAsync(dst, src[p], row_bytes * ph, cudaMemcpyHostToDevice, c->up));
                else CK(cudaMemcpy2DAsync(dst, dpitch, src[p], stride[p], row_bytes, ph, cudaMemcpyHostToDevice, c->up));
                g->p[f][clip][p] = dst;
                g->pitch[clip][p] = dpitch;
            }
        }
    } else {
        for (int clip = 0; clip < 2; ++clip) {
            const void *const *src = clip == 0 ? rp : dp;
            const size_t *stride = clip == 0 ? rs : ds;
            for (int p = 0; p < np; ++p) {
                if (g->pitch_set && g->pitch[clip][p] != stride[p])
                    return fail(c, BV_ERR_ARG, "bv_submit_device: plane pitch must be constant within a frame group");
                g->p[f][clip][p] = static_cast<const uint8_t *>(src[p]);
                g->pitch[clip][p] = stride[p];
            }
        }
    }
    g->pitch_set = true;
    g->frame_index[f] = frame_index;
    g->flags[f] = flags;
    g->n = f + 1;
    c->submitted += 1;
    if (g->n == c->B) return launch_group(c, *g);
    return 0;
}

}  // namespace

// ================================================================================================
extern "C" {

int bv_abi_version(void) { return BV_ABI_VERSION; }

int bv_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

bv_ctx *bv_create(int device, int width, int height, int bpc, int chroma, unsigned features_mask, const bv_opts *opts)
{
    g_create_error.clear();
    if (width < 32 || height < 32 || width > 16384 || height > 16384) { fail(nullptr, BV_ERR_ARG, "bv_create: width/height must be in [32, 16384]"); return nullptr; }
    if (bpc != 8 && bpc != 10 && bpc != 12 && bpc != 16) { fail(nullptr, BV_ERR_ARG, "bv_create: bpc must be 8, 10, 12 or 16"); return nullptr; }
    if (chroma != 0 && chroma != 400 && chroma != 420 && chroma != 422 && chroma != 444) { fail(nullptr, BV_ERR_ARG, "bv_create: chroma must be 0, 400, 420, 422 or 444"); return nullptr; }
    if (features_mask == 0) { fail(nullptr, BV_ERR_ARG, "bv_create: empty features_mask"); return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { fail(nullptr, BV_ERR_CUDA, "bv_create: no CUDA device (this engine has no CPU fallback)", e); cudaGetLastError(); return nullptr; }
    if (device < 0 || device >= ndev) { fail(nullptr, BV_ERR_ARG, "bv_create: device out of range"); return nullptr; }
    bv_ctx *c = new bv_ctx();
    c->device = device; c->w = width; c->h = height; c->bpc = bpc; c->chroma = chroma; c->feat = features_mask;
    c->cw = chroma == 420 || chroma == 422 ? (width + 1) / 2 : width;
    c->ch = chroma == 420 ? (height + 1) / 2 : height;
    bv_opts d;
    memset(&d, 0, sizeof d);
    d.vif_enhn_gain_limit = 100.0; d.adm_enhn_gain_limit = 100.0; d.adm_norm_view_dist = 3.0; d.adm_ref_display_height = 1080;
    if (opts) {
        d = *opts;
        if (!(d.vif_enhn_gain_limit >= 1.0)) d.vif_enhn_gain_limit = 100.0;
        if (!(d.adm_enhn_gain_limit >= 1.0)) d.adm_enhn_gain_limit = 100.0;
        if (!(d.adm_norm_view_dist > 0.0)) d.adm_norm_view_dist = 3.0;
        if (d.adm_ref_display_height <= 0) d.adm_ref_display_height = 1080;
    }
    c->opts = d;
    if (d.batch_frames > 0) c->B = d.batch_frames > BV_MAX_BATCH ? BV_MAX_BATCH : d.batch_frames;
    else {
        // auto: twice the work of a launch group of 32 1080p frames, capped at 32 frames (enough tiles to fill 148
        // SMs at every pyramid level, launch gaps amortised).  Larger pictures reach it with fewer frames -- 2160p: 16
        // -- which bounds the per-context footprint and the pipeline fill/drain time (one group).
        const double px = (double)width * height;
        int b = (int)(2.0 * BV_MAX_BATCH * (1920.0 * 1080.0) / px + 0.5);
        c->B = b < 4 ? 4 : (b > BV_MAX_BATCH ? BV_MAX_BATCH : b);
    }
    if (alloc_ctx(c) != 0) {
        g_create_error = c->err;
        bv_destroy(c);
        return nullptr;
    }
    return c;
}

void bv_destroy(bv_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->comp) cudaStreamSynchronize(c->comp);
    if (c->up) cudaStreamSynchronize(c->up);
    for (int s = 0; s < N_SLOTS; ++s) {
        Group &g = c->groups[s];
        for (int p = 0; p < 3; ++p) for (int k = 0; k < 2; ++k) if (g.d_planes[p][k]) cudaFree(g.d_planes[p][k]);
        if (g.d_raw) cudaFree(g.d_raw);
        if (g.h_raw) cudaFreeHost(g.h_raw);
        if (g.d_fraw) cudaFree(g.d_fraw);
        if (g.h_fraw) cudaFreeHost(g.h_fraw);
        if (g.uploaded) cudaEventDestroy(g.uploaded);
        if (g.done) cudaEventDestroy(g.done);
        if (g.prof.have_events)
            for (int k = 0; k < BV_MAX_KERNELS; ++k) { cudaEventDestroy(g.prof.ev[k][0]); cudaEventDestroy(g.prof.ev[k][1]); }
    }
    for (int k = 0; k < 2; ++k) if (c->blur[k]) cudaFree(c->blur[k]);
    if (c->d_log2) cudaFree(c->d_log2);
    if (c->d_log2c) cudaFree(c->d_log2c);
    if (c->feat & BV_FEAT_VIF) for (int s = 1; s < 4; ++s) { if (c->vif_lv.ref[s]) cudaFree(c->vif_lv.ref[s]); if (c->vif_lv.dis[s]) cudaFree(c->vif_lv.dis[s]); }
    if (c->feat & BV_FEAT_ADM) { for (int s = 0; s < 3; ++s) if (c->adm.bands[s]) cudaFree(c->adm.bands[s]); if (c->adm.rows) cudaFree(c->adm.rows); }
    if (c->d_div) cudaFree(c->d_div);
    if (c->fl) bv_float_destroy(c->fl);
    if (c->flf) bv_float_fast_destroy(c->flf);
    for (int k = 0; k < 2; ++k) if (c->t_ev[k]) cudaEventDestroy(c->t_ev[k]);
    if (c->up) cudaStreamDestroy(c->up);
    if (c->comp) cudaStreamDestroy(c->comp);
    delete c;
}

const char *bv_last_error(bv_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int bv_pinned_alloc(void **p, size_t n)
{
    if (!p) return BV_ERR_ARG;
    cudaError_t e = cudaHostAlloc(p, n, cudaHostAllocDefault);
    if (e != cudaSuccess) { fail(nullptr, BV_ERR_CUDA, "cudaHostAlloc", e); cudaGetLastError(); return BV_ERR_CUDA; }
    return 0;
}
int bv_pinned_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? 0 : BV_ERR_CUDA; }

int bv_host_register(void *p, size_t n, int read_only)
{
    if (!p || n == 0) return BV_ERR_ARG;
    cudaError_t e = cudaHostRegister(p, n, read_only ? cudaHostRegisterReadOnly : cudaHostRegisterDefault);
    if (e != cudaSuccess) { fail(nullptr, BV_ERR_CUDA, "cudaHostRegister", e); cudaGetLastError(); return BV_ERR_CUDA; }
    return 0;
}
int bv_host_unregister(void *p)
{
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return BV_ERR_CUDA; }
    return 0;
}

int bv_device_alloc(int device, void **p, size_t n)
{
    if (!p) return BV_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(p, n) != cudaSuccess) {
        fail(nullptr, BV_ERR_CUDA, "cudaMalloc", cudaGetLastError());
        return BV_ERR_CUDA;
    }
    return 0;
}
int bv_device_free(int device, void *p) { cudaSetDevice(device); return cudaFree(p) == cudaSuccess ? 0 : BV_ERR_CUDA; }
int bv_device_upload(int device, void *dst, const void *src, size_t n)
{
    cudaSetDevice(device);
    return cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : BV_ERR_CUDA;
}

int bv_submit(bv_ctx *c, int64_t frame_index, const void *const ref_planes[3], const size_t ref_stride[3],
              const void *const dis_planes[3], const size_t dis_stride[3], unsigned frame_flags)
{
    return submit_common(c, frame_index, ref_planes, ref_stride, dis_planes, dis_stride, frame_flags, true);
}

int bv_submit_device(bv_ctx *c, int64_t frame_index, const void *const ref_planes[3], const size_t ref_stride[3],
                     const void *const dis_planes[3], const size_t dis_stride[3], unsigned frame_flags)
{
    return submit_common(c, frame_index, ref_planes, ref_stride, dis_planes, dis_stride, frame_flags, false);
}

int bv_wait_uploads(bv_ctx *c)
{
    if (!c) return BV_ERR_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->up));
    return 0;
}

int bv_flush(bv_ctx *c)
{
    if (!c) return BV_ERR_ARG;
    if (c->cancelled.load()) return fail(c, BV_ERR_CANCELLED, "cancelled");
    CK(cudaSetDevice(c->device));
    Group &g = c->groups[c->cur];
    if (!g.in_flight && g.n > 0) { int rc = launch_group(c, g); if (rc) return rc; }
    // harvest in submission order
    for (int k = 0; k < N_SLOTS; ++k) {
        int best = -1;
        for (int s = 0; s < N_SLOTS; ++s)
            if (c->groups[s].in_flight && (best < 0 || c->groups[s].first_ordinal < c->groups[best].first_ordinal)) best = s;
        if (best < 0) break;
        int rc = harvest(c, c->groups[best]);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(c->up));
    CK(cudaStreamSynchronize(c->comp));
    return 0;
}

int bv_kick(bv_ctx *c)
{
    if (!c) return BV_ERR_ARG;
    if (c->cancelled.load()) return fail(c, BV_ERR_CANCELLED, "cancelled");
    CK(cudaSetDevice(c->device));
    Group &g = c->groups[c->cur];
    if (!g.in_flight && g.n > 0) return launch_group(c, g);
    return 0;
}

int64_t bv_frames_done(bv_ctx *c) { return c ? c->ready.load() : 0; }

int bv_batch_frames(bv_ctx *c) { return c ? c->B : 0; }

int bv_reset(bv_ctx *c)
{
    if (!c) return BV_ERR_ARG;
    c->cancelled.store(0);
    int rc = bv_flush(c);
    if (rc) return rc;
    c->results.clear();
    c->submitted = 0;
    c->ready.store(0);
    c->blur_prev_n = 0;
    if (c->fl) bv_float_reset(c->fl);
    if (c->flf) bv_float_fast_reset(c->flf);
    return 0;
}

int bv_fetch(bv_ctx *c, int64_t first, int64_t count, bv_frame_features *out)
{
    if (!c || !out || first < 0 || count < 0) return BV_ERR_ARG;
    if (first + count > c->submitted) return fail(c, BV_ERR_ORDER, "bv_fetch: frames not submitted yet");
    if (first + count > c->ready.load()) {
        int rc = bv_flush(c);
        if (rc) return rc;
    }
    if (first + count > (int64_t)c->results.size()) return fail(c, BV_ERR_ORDER, "bv_fetch: results missing");
    memcpy(out, c->results.data() + first, sizeof(bv_frame_features) * (size_t)count);
    return 0;
}

int bv_cancel(bv_ctx *c)
{
    if (!c) return BV_ERR_ARG;
    c->cancelled.store(1);
    return 0;
}

int64_t bv_kernel_launches(bv_ctx *c) { return c ? c->nlaunch : 0; }

int bv_set_profiling(bv_ctx *c, int enable)
{
    if (!c) return BV_ERR_ARG;
    c->profiling = enable != 0;
    return 0;
}

int bv_kernel_slots(void) { return BV_MAX_KERNELS; }

const char *bv_kernel_name(int id)
{
    if (id < 0 || id >= BV_MAX_KERNELS) return nullptr;
    if (id >= BVK_F_FIRST) return bv_float_kernel_name(id);
    return k_kernel_names[id];
}

double bv_kernel_ms(bv_ctx *c, int id, int reset)
{
    if (!c || id < 0 || id >= BV_MAX_KERNELS) return -1.0;
    const double v = c->k_ms[id];
    if (reset) c->k_ms[id] = 0.0;
    return v;
}

double bv_kernel_count(bv_ctx *c, int id, int reset)
{
    if (!c || id < 0 || id >= BV_MAX_KERNELS) return -1.0;
    const double v = c->k_count[id];
    if (reset) c->k_count[id] = 0.0;
    return v;
}

// Device-side stopwatch on the compute stream: mark(0) before the first submit of a timed region,
// mark(1) after the last; elapsed = time between the two marks as the GPU saw them.
int bv_timer_mark(bv_ctx *c, int which)
{
    if (!c || which < 0 || which > 1) return BV_ERR_ARG;
    CK(cudaSetDevice(c->device));
    if (which == 1) {
        // everything submitted so far must be in the stream before the end mark
        Group &g = c->groups[c->cur];
        if (!g.in_flight && g.n > 0) { int rc = launch_group(c, g); if (rc) return rc; }
    }
    CK(cudaEventRecord(c->t_ev[which], c->comp));
    c->t_set[which] = true;
    return 0;
}

double bv_timer_elapsed_ms(bv_ctx *c)
{
    if (!c || !c->t_set[0] || !c->t_set[1]) return -1.0;
    cudaSetDevice(c->device);
    if (cudaEventSynchronize(c->t_ev[1]) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->t_ev[0], c->t_ev[1]) != cudaSuccess) return -1.0;
    return (double)ms;
}

size_t bv_sizeof_frame_features(void) { return sizeof(bv_frame_features); }

}  // extern "C"
