// adm_scale<S> / adm_rows_finish -- integer ADM at 4 scales (replaces libvmaf integer_adm.c, reached
// from the reference at app/vmaf_analyzer.py:417; algorithm: SURVEY.md Appendix A.4).
//
// One fused kernel per scale: db2 DWT (vertical then horizontal, MIRROR borders, libvmaf's
// per-scale shifts) of the ref and dis tile -> band_a written for the next scale; h/v/d bands stay
// on chip -> decouple (Q15 reciprocal LUT, 1-degree angle test, enhancement-gain limit) -> CSF
// weighting -> 3x3 contrast-masking threshold -> cubed numerator, and the cubed CSF denominator of
// the reference bands.  libvmaf rounds its accumulators once per image ROW, so the kernel produces
// exact per-row 64-bit sums (warp shuffle -> shared -> one global atomic per row/band/CTA) and
// adm_rows_finish applies the row shift and adds the rows.  All arithmetic is integer except the
// angle test and the gain limit, which follow the C promotion rules in IEEE double (no FMA).
//
// Persistent CTAs (2 per SM) loop over (frame, tile) items; tile = 64x16 band pixels (+1 halo for the 3x3
// threshold) = 136x38 staged input samples per picture, prefetched into registers one tile ahead.
#include "bv_common.cuh"
#include "../../include/b200vmaf.h"
#include "../../include/libvmaf_spec.h"
#include <math.h>
#include <type_traits>

namespace {

constexpr int AT_W = 64, AT_H = 16, AT_THREADS = 256;
constexpr int AP_W = AT_W + 2, AP_H = AT_H + 2;          // band positions incl. the contrast-masking halo
constexpr int AN_C = 2 * AT_W + 8, AN_R = 2 * AT_H + 6;  // staged input samples (columns start at 2*tx0 - 4: 4-sample aligned)
constexpr int AN_P = AN_C;                               // shared-memory pitch (elements)
constexpr int AN_G = AN_C / 4;                           // 4-sample groups per staged row
constexpr int A_RING = 2 * AP_W + 2 * AT_H;              // halo-only positions

__constant__ int c_dwt_lo[4] = { SPEC_DWT_LO_Q15 };
__constant__ int c_dwt_hi[4] = { SPEC_DWT_HI_Q15 };
constexpr int DWT_LO_SUM = SPEC_DWT_LO_SUM_Q15;

struct AdmArgs {
    BvPlane ref, dis;                // input of this scale (picture or previous band_a)
    void *a_ref, *a_dis;             // band_a outputs, frame f at + f * a_frame_elems (unused at scale 3)
    size_t a_frame_elems;
    BvAdmScaleParams sp;
    int bpc;
    const int *div_lookup;
    double egl;
    float cos_1deg_sq;
    unsigned long long *rows;        // [frame][rows_frame_stride]; this scale at + rows_offset: [row][6]
    size_t rows_frame_stride, rows_offset;
    int vec_ok;                      // planes aligned for 4-sample vector loads
};

template <int SCALE> struct AdmTypes;
template <> struct AdmTypes<0> { using Stage = uint16_t; using V = short4; using Out = int16_t; };
template <> struct AdmTypes<1> { using Stage = int16_t;  using V = int4;   using Out = int32_t; };
template <> struct AdmTypes<2> { using Stage = int32_t;  using V = int4;   using Out = int32_t; };
template <> struct AdmTypes<3> { using Stage = int32_t;  using V = int4;   using Out = int32_t; };

template <int SCALE> struct AdmShifts {
    // DWT shifts of scales 1..3 (index SCALE-1); scale 0 uses the picture depth
    static constexpr int sh_v = SCALE == 1 ? 0 : 16;
    static constexpr long long rnd_v = SCALE == 1 ? 0 : 32768;
    static constexpr int sh_h = SCALE == 2 ? 16 : 15;
    static constexpr long long rnd_h = SCALE == 2 ? 32768 : 16384;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// libvmaf's div_lookup[d] = 2^30 / d (truncated) for d in [1, 32768], computed instead of loaded at scales 1-3: there the
// index is a normalised 15-bit value, the 256 KB table does not stay in L1 and each lookup was a dependent L2 round trip
// (long_scoreboard was the top stall), while the
// FP64 pipe is idle in this kernel.  Exact: 1/d is correctly rounded (error < 2^-53 relative), scaling by 2^30 is exact,
// and for d not a power of two 2^30/d is at least 2^-15 away from an integer, so truncation cannot cross one
// (tests/test_cpu_boundary.py checks all 32768 values against the integer division).
__device__ __forceinline__ int adm_recip_q30(int d) { return __double2int_rz(__dmul_rn(__drcp_rn((double)d), 1073741824.0)); }

// Decouple + CSF of one band position.  o/t: reference / distorted (h, v, d).
// Returns |csf-weighted restored| per band, the 3-band sums of csf_f (neighbour weight 1/30) and of
// the centre weight (1/15).
template <int SCALE>
__device__ __forceinline__ void adm_decouple_csf(const int (&o)[3], const int (&t)[3], const AdmArgs &a,
                                                 int (&xabs)[3], int &cfsum, int &ccsum)
{
    const long long ot_dp = (long long)o[0] * t[0] + (long long)o[1] * t[1];
    const long long o_mag = (long long)o[0] * o[0] + (long long)o[1] * o[1];
    const long long t_mag = (long long)t[0] * t[0] + (long long)t[1] * t[1];
    const double fa = (double)__ll2float_rn(ot_dp) * (1.0 / 4096.0);
    const double fo = (double)__ll2float_rn(o_mag) * (1.0 / 4096.0);
    const double ft = (double)__ll2float_rn(t_mag) * (1.0 / 4096.0);
    const bool flag = (fa >= 0.0) &&
                      (__dmul_rn(fa, fa) >= __dmul_rn(__dmul_rn((double)a.cos_1deg_sq, fo), ft));
    unsigned cf_acc = 0, cc_acc = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const int ob = o[b], tb = t[b];
        int k;
        if (ob == 0) {
            k = 32768;
        } else if (SCALE == 0) {
            // scale 0: band values are small and cluster around 0, the table lines stay in L1 (the computed form measured 2 % slower here)
            const int tmp = (int)(((long long)__ldg(a.div_lookup + ob + 32768) * tb + 16384) >> 15);
            k = clampi(tmp, 0, 32768);
        } else {
            const unsigned ao = (unsigned)abs(ob);
            unsigned msb = ao;
            int sh = 0;
            if (ao >= 32768u) {
                sh = 17 - __clz(ao);
                msb = (ao + (1u << (sh - 1))) >> sh;
            }
            long long prod = (long long)adm_recip_q30((int)msb) * tb;
            if (ob < 0) prod = -prod;
            const long long tmp = (prod + (1ll << (14 + sh))) >> (15 + sh);
            k = (int)(tmp < 0 ? 0 : (tmp > 32768 ? 32768 : tmp));
        }
        int rst = (int)(((long long)k * ob + 16384) >> 15);
        if (SCALE == 0) rst = (short)rst;
        if (flag && k > 0) {
            const double v = __dmul_rn((double)rst, a.egl), tt = (double)tb;
            if (ob > 0) rst = __double2int_rz(v < tt ? v : tt);
            else if (ob < 0) rst = __double2int_rz(v > tt ? v : tt);
        }
        if (SCALE == 0) rst = (short)rst;
        int ad = tb - rst;
        if (SCALE == 0) ad = (short)ad;
        int ca, x;
        if (SCALE == 0) {
            const int sh = b == 2 ? 17 : 15;
            const int dv = (int)(a.sp.rf[b] * (unsigned)ad);
            ca = (short)((dv + (1 << (sh - 1))) >> sh);
            cf_acc += (unsigned)(int)(short)((SPEC_ADM_ONE_BY_30_Q16 * abs(ca) + 2048) >> 12);
            cc_acc += (unsigned)(int)(short)((SPEC_ADM_ONE_BY_15_Q16 * abs(ca) + 2048) >> 12);
            x = (int)((unsigned)rst * a.sp.rf[b]);
        } else {
            ca = (int)(((long long)a.sp.rf[b] * (long long)ad + (1ll << 27)) >> 28);
            cf_acc += (unsigned)(int)((SPEC_ADM_ONE_BY_30_Q32 * abs(ca) + (1ll << 31)) >> 32);
            cc_acc += (unsigned)(int)((SPEC_ADM_ONE_BY_15_Q32 * abs(ca) + (1ll << 31)) >> 32);
            x = (int)(((long long)rst * (long long)a.sp.rf[b] + (1ll << 27)) >> 28);
        }
        xabs[b] = abs(x);
    }
    cfsum = (int)cf_acc;
    ccsum = (int)cc_acc;
}

template <typename Stage> struct StageStore;
template <> struct StageStore<uint16_t> {
    static __device__ __forceinline__ void st(uint16_t *p, const unsigned (&u)[4]) { *reinterpret_cast<uint2 *>(p) = make_uint2(u[0] | (u[1] << 16), u[2] | (u[3] << 16)); }
};
template <> struct StageStore<int16_t> {
    static __device__ __forceinline__ void st(int16_t *p, const int (&u)[4])
    {
        *reinterpret_cast<uint2 *>(p) = make_uint2(((unsigned)u[0] & 0xffffu) | ((unsigned)u[1] << 16), ((unsigned)u[2] & 0xffffu) | ((unsigned)u[3] << 16));
    }
};
template <> struct StageStore<int32_t> {
    static __device__ __forceinline__ void st(int32_t *p, const int (&u)[4]) { *reinterpret_cast<int4 *>(p) = make_int4(u[0], u[1], u[2], u[3]); }
};

// Persistent CTAs over (frame, tile) work items with register prefetch of the next tile's samples.
template <int SCALE, typename TIn>
__global__ void __launch_bounds__(AT_THREADS, SCALE == 0 ? 3 : 2)
adm_scale_kernel(BvBatch batch, AdmArgs a, BvDiv tiles_x, BvDiv tiles_per_frame, int total_tiles)
{
    using Stage = typename AdmTypes<SCALE>::Stage;
    using VT = typename AdmTypes<SCALE>::V;
    using Out = typename AdmTypes<SCALE>::Out;
    using V4 = typename Px4<TIn>::V;
    using Raw = typename std::conditional<std::is_signed<TIn>::value, int, unsigned>::type;
    constexpr int NGRP = AN_R * AN_G, NPF = (NGRP + AT_THREADS - 1) / AT_THREADS;

    extern __shared__ __align__(16) unsigned char smem[];
    VT *s_v = reinterpret_cast<VT *>(smem);                                          // [AP_H][AN_P]
    int *s_cf = reinterpret_cast<int *>(smem + sizeof(VT) * AP_H * AN_P);            // [AP_H][AP_W]
    int *s_x = s_cf + AP_H * AP_W;                                                   // [3][AT_H*AT_W]
    int *s_cc = s_x + 3 * AT_H * AT_W;                                               // [AT_H*AT_W]
    unsigned long long *s_row = reinterpret_cast<unsigned long long *>(s_cc + AT_H * AT_W);   // [AT_H][2 half rows][6]
    Stage *s_in = reinterpret_cast<Stage *>(s_row + AT_H * 12);                      // [2][AN_R][AN_P]

    __shared__ int4 s_rk[AP_H], s_ck[AP_W];          // staged row / column of the 4 DWT taps of a band row / column
    __shared__ int2 s_rinfo[AP_H], s_cinfo[AP_W];    // (mirrored band index, region flags: 1 valid, 2 in image, 4 decouple region, 8 core)

    const int in_w = a.sp.in_w, in_h = a.sp.in_h, ow = a.sp.w, oh = a.sp.h;
    const int tid = threadIdx.x;
    V4 pre_r[NPF], pre_d[NPF];

    auto prefetch = [&](int t) {
        const int f = t / tiles_per_frame, rem = t - f * tiles_per_frame;
        if (batch.flags[f] & (BV_FRAME_LEAD_IN | BV_FRAME_SKIP_SPATIAL)) return;
        const int by = rem / tiles_x, bx = rem - by * tiles_x;
        const int cx0 = 2 * bx * AT_W - 4, ry0 = 2 * by * AT_H - 3;
        const uint8_t *pr = a.ref.p[f], *pd = a.dis.p[f];
#pragma unroll
        for (int k = 0; k < NPF; ++k) {
            const int g = tid + k * AT_THREADS;
            if (g < NGRP) {
                const int r = g / AN_G, gc = g - r * AN_G;
                const int gy = bv_mirror(clampi(ry0 + r, -(in_h - 1), 2 * in_h - 1), in_h);
                pre_r[k] = load_px4<TIn>(pr + (size_t)gy * a.ref.pitch, cx0 + 4 * gc, in_w, 2 * in_w - 1, a.vec_ok);
                pre_d[k] = load_px4<TIn>(pd + (size_t)gy * a.dis.pitch, cx0 + 4 * gc, in_w, 2 * in_w - 1, a.vec_ok);
            }
        }
    };

    // per-(tile row, half row, band) sums: zeroed here once, then by the thread that consumes a slot at the end of a
    // tile (the same thread reads and clears, so no other thread's clear can overtake the read)
    if (tid < AT_H * 12) s_row[tid] = 0ull;
    int t = blockIdx.x;
    if (t < total_tiles) prefetch(t);
    for (; t < total_tiles; t += gridDim.x) {
    const int f = t / tiles_per_frame, rem = t - f * tiles_per_frame;
    const bool skip = batch.flags[f] & (BV_FRAME_LEAD_IN | BV_FRAME_SKIP_SPATIAL);          // CTA-uniform
    const int tx0 = (rem % tiles_x) * AT_W, ty0 = (rem / tiles_x) * AT_H;
    const int cx0 = 2 * tx0 - 4, ry0 = 2 * ty0 - 3;

    // Per-tile index tables: the MIRROR / clamp arithmetic of the two DWT passes and the region tests depend only on
    // the band row or the band column, so they are evaluated once per row / column here instead of once per tap of
    // every position (ncu: that arithmetic was ~25 % of the executed instructions).
    const int left = a.sp.left, top = a.sp.top, right = a.sp.right, bottom = a.sp.bottom;
    if (tid < AP_H) {
        const int r = tid, bi_raw = ty0 - 1 + r;
        const int bi = bv_mirror(clampi(bi_raw, -1, oh), oh);
        int4 rk;
        rk.x = clampi(bv_mirror(2 * bi - 1, in_h) - ry0, 0, AN_R - 1);
        rk.y = clampi(bv_mirror(2 * bi, in_h) - ry0, 0, AN_R - 1);
        rk.z = clampi(bv_mirror(2 * bi + 1, in_h) - ry0, 0, AN_R - 1);
        rk.w = clampi(bv_mirror(2 * bi + 2, in_h) - ry0, 0, AN_R - 1);
        s_rk[r] = rk;
        const int gt = max(top - 1, 0), gb = min(bottom + 1, oh);
        s_rinfo[r] = make_int2(bi, (bi_raw >= -1 && bi_raw <= oh ? 1 : 0) | (bi_raw < oh ? 2 : 0) |
                                   (bi >= gt && bi < gb ? 4 : 0) | (bi >= top && bi < bottom ? 8 : 0));
    } else if (tid >= 32 && tid < 32 + AP_W) {
        const int c = tid - 32, bj_raw = tx0 - 1 + c;
        const int bj = bv_mirror(clampi(bj_raw, -1, ow), ow);
        int4 ck;
        ck.x = clampi(bv_mirror(2 * bj - 1, in_w) - cx0, 0, AN_C - 1);
        ck.y = clampi(bv_mirror(2 * bj, in_w) - cx0, 0, AN_C - 1);
        ck.z = clampi(bv_mirror(2 * bj + 1, in_w) - cx0, 0, AN_C - 1);
        ck.w = clampi(bv_mirror(2 * bj + 2, in_w) - cx0, 0, AN_C - 1);
        s_ck[c] = ck;
        const int gl = max(left - 1, 0), gr = min(right + 1, ow);
        s_cinfo[c] = make_int2(bj, (bj_raw >= -1 && bj_raw <= ow ? 1 : 0) | (bj_raw < ow ? 2 : 0) |
                                   (bj >= gl && bj < gr ? 4 : 0) | (bj >= left && bj < right ? 8 : 0));
    }

    // ---- phase A: registers -> shared (MIRROR was resolved by the loads) ----
    if (!skip) {
#pragma unroll
        for (int k = 0; k < NPF; ++k) {
            const int g = tid + k * AT_THREADS;
            if (g < NGRP) {
                const int r = g / AN_G, gc = g - r * AN_G;
                Raw ur[4], ud[4];
                Px4<TIn>::raw(pre_r[k], ur);
                Px4<TIn>::raw(pre_d[k], ud);
                StageStore<Stage>::st(s_in + r * AN_P + 4 * gc, ur);
                StageStore<Stage>::st(s_in + (AN_R + r) * AN_P + 4 * gc, ud);
            }
        }
    }
    __syncthreads();
    if (t + (int)gridDim.x < total_tiles) prefetch(t + gridDim.x);
    if (skip) continue;

    // ---- phase B: vertical DWT pass: lo/hi of ref and dis for every band row of the halo tile ----
    for (int idx = tid; idx < AP_H * AN_C; idx += AT_THREADS) {
        const int r = idx / AN_C, c = idx - r * AN_C;
        const int4 rk4 = s_rk[r];
        const int rk[4] = { rk4.x, rk4.y, rk4.z, rk4.w };
        VT out;
        if (SCALE == 0) {
            const int add_v = 1 << (a.bpc - 1);
            int lo_r = 0, hi_r = 0, lo_d = 0, hi_d = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xr = (int)s_in[rk[k] * AN_P + c], xd = (int)s_in[(AN_R + rk[k]) * AN_P + c];
                lo_r += c_dwt_lo[k] * xr; hi_r += c_dwt_hi[k] * xr;
                lo_d += c_dwt_lo[k] * xd; hi_d += c_dwt_hi[k] * xd;
            }
            lo_r -= DWT_LO_SUM * add_v; lo_d -= DWT_LO_SUM * add_v;
            out.x = (short)((lo_r + add_v) >> a.bpc); out.y = (short)((hi_r + add_v) >> a.bpc);
            out.z = (short)((lo_d + add_v) >> a.bpc); out.w = (short)((hi_d + add_v) >> a.bpc);
        } else {
            long long lo_r = 0, hi_r = 0, lo_d = 0, hi_d = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long xr = (long long)s_in[rk[k] * AN_P + c], xd = (long long)s_in[(AN_R + rk[k]) * AN_P + c];
                lo_r += c_dwt_lo[k] * xr; hi_r += c_dwt_hi[k] * xr;
                lo_d += c_dwt_lo[k] * xd; hi_d += c_dwt_hi[k] * xd;
            }
            constexpr int sh = AdmShifts<SCALE>::sh_v;
            constexpr long long rnd = AdmShifts<SCALE>::rnd_v;
            out.x = (int)((lo_r + rnd) >> sh); out.y = (int)((hi_r + rnd) >> sh);
            out.z = (int)((lo_d + rnd) >> sh); out.w = (int)((hi_d + rnd) >> sh);
        }
        s_v[r * AN_P + c] = out;
    }
    __syncthreads();

    // ---- phase C: horizontal DWT pass + decouple + CSF for every position (interior first) ----
#pragma unroll 1
    for (int p = tid; p < AT_H * AT_W + A_RING; p += AT_THREADS) {
        const bool interior = p < AT_H * AT_W;         // warp-uniform (AT_H*AT_W is a multiple of 32)
        int r, c;
        if (interior) {
            r = p / AT_W + 1; c = p % AT_W + 1;
        } else {
            const int q = p - AT_H * AT_W;
            if (q < AP_W) { r = 0; c = q; }
            else if (q < 2 * AP_W) { r = AP_H - 1; c = q - AP_W; }
            else { r = 1 + ((q - 2 * AP_W) >> 1); c = ((q - 2 * AP_W) & 1) ? AP_W - 1 : 0; }
        }
        const int2 ri = s_rinfo[r], ci = s_cinfo[c];
        const int bi = ri.x, bj = ci.x, fl = ri.y & ci.y;
        const bool in_img = fl & 2;                                // interior positions only
        const bool in_g = (fl & 5) == 5;
        const bool core = interior && (fl & 10) == 10;

        unsigned long long dsum[3] = { 0ull, 0ull, 0ull };
        int cfsum = 0;
        if (in_g || (interior && in_img && SCALE < 3)) {
            const int4 ck = s_ck[c];
            const VT *vrow = s_v + r * AN_P;
            const VT tv[4] = { vrow[ck.x], vrow[ck.y], vrow[ck.z], vrow[ck.w] };
            if (SCALE < 3 && interior && in_img) {
                Out ar, ad;
                if (SCALE == 0) {
                    int s_r = 0, s_d = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) { s_r += c_dwt_lo[k] * (int)tv[k].x; s_d += c_dwt_lo[k] * (int)tv[k].z; }
                    ar = (Out)((s_r + 32768) >> 16); ad = (Out)((s_d + 32768) >> 16);
                } else {
                    long long s_r = 0, s_d = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) { s_r += (long long)c_dwt_lo[k] * tv[k].x; s_d += (long long)c_dwt_lo[k] * tv[k].z; }
                    ar = (Out)((s_r + AdmShifts<SCALE>::rnd_h) >> AdmShifts<SCALE>::sh_h);
                    ad = (Out)((s_d + AdmShifts<SCALE>::rnd_h) >> AdmShifts<SCALE>::sh_h);
                }
                const size_t off = (size_t)f * a.a_frame_elems + (size_t)bi * ow + bj;
                static_cast<Out *>(a.a_ref)[off] = ar;
                static_cast<Out *>(a.a_dis)[off] = ad;
            }
            if (in_g) {
                int o[3], t[3];       // (h, v, d): h = lo_H(hi_V), v = hi_H(lo_V), d = hi_H(hi_V)
                if (SCALE == 0) {
                    int oh_ = 0, ov_ = 0, od_ = 0, th_ = 0, tv_ = 0, td_ = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        oh_ += c_dwt_lo[k] * (int)tv[k].y; ov_ += c_dwt_hi[k] * (int)tv[k].x; od_ += c_dwt_hi[k] * (int)tv[k].y;
                        th_ += c_dwt_lo[k] * (int)tv[k].w; tv_ += c_dwt_hi[k] * (int)tv[k].z; td_ += c_dwt_hi[k] * (int)tv[k].w;
                    }
                    o[0] = (short)((oh_ + 32768) >> 16); o[1] = (short)((ov_ + 32768) >> 16); o[2] = (short)((od_ + 32768) >> 16);
                    t[0] = (short)((th_ + 32768) >> 16); t[1] = (short)((tv_ + 32768) >> 16); t[2] = (short)((td_ + 32768) >> 16);
                } else {
                    long long oh_ = 0, ov_ = 0, od_ = 0, th_ = 0, tv_ = 0, td_ = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        oh_ += (long long)c_dwt_lo[k] * tv[k].y; ov_ += (long long)c_dwt_hi[k] * tv[k].x; od_ += (long long)c_dwt_hi[k] * tv[k].y;
                        th_ += (long long)c_dwt_lo[k] * tv[k].w; tv_ += (long long)c_dwt_hi[k] * tv[k].z; td_ += (long long)c_dwt_hi[k] * tv[k].w;
                    }
                    constexpr int sh = AdmShifts<SCALE>::sh_h;
                    constexpr long long rnd = AdmShifts<SCALE>::rnd_h;
                    o[0] = (int)((oh_ + rnd) >> sh); o[1] = (int)((ov_ + rnd) >> sh); o[2] = (int)((od_ + rnd) >> sh);
                    t[0] = (int)((th_ + rnd) >> sh); t[1] = (int)((tv_ + rnd) >> sh); t[2] = (int)((td_ + rnd) >> sh);
                }
                int xabs[3], ccsum;
                adm_decouple_csf<SCALE>(o, t, a, xabs, cfsum, ccsum);
                if (interior) {
                    const int q = p;
                    s_x[q] = xabs[0]; s_x[AT_H * AT_W + q] = xabs[1]; s_x[2 * AT_H * AT_W + q] = xabs[2];
                    s_cc[q] = ccsum;
                }
                if (core) {
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        if (SCALE == 0) {
                            const unsigned long long v = (unsigned long long)(uint16_t)abs(o[b]);
                            dsum[b] = v * v * v;
                        } else {
                            const unsigned long long v = (unsigned long long)(unsigned)abs(o[b]);
                            dsum[b] = ((((v * v + a.sp.den_add_sq) >> a.sp.den_sh_sq) * v) + a.sp.den_add_cub) >> a.sp.den_sh_cub;
                        }
                    }
                }
            }
        }
        s_cf[r * AP_W + c] = cfsum;
        if (interior && __any_sync(0xffffffffu, core)) {
            // a warp covers 32 consecutive interior columns of one tile row
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const unsigned long long s = (unsigned long long)bv_warp_sum_redux((long long)dsum[b]);      // < 2^56
                // a (tile row, half row) pair belongs to exactly one warp and one trip: plain store, no atomic
                if ((tid & 31) == 0) s_row[((r - 1) * 2 + (c > 32)) * 6 + 3 + b] = s;
            }
        }
    }
    __syncthreads();

    // ---- phase D: 3x3 contrast-masking threshold and the cubed numerator ----
#pragma unroll 1
    for (int p = tid; p < AT_H * AT_W; p += AT_THREADS) {
        const int r = p / AT_W + 1, c = p % AT_W + 1;
        const bool core = ((s_rinfo[r].y & s_cinfo[c].y) & 10) == 10;
        if (!__any_sync(0xffffffffu, core)) continue;
        long long val[3] = { 0, 0, 0 };
        if (core) {
            unsigned thr = 0;
#pragma unroll
            for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
                for (int dc = -1; dc <= 1; ++dc) thr += (unsigned)s_cf[(r + dr) * AP_W + c + dc];
            thr = thr - (unsigned)s_cf[r * AP_W + c] + (unsigned)s_cc[p];
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                int x = s_x[b * AT_H * AT_W + p] - (int)(thr << a.sp.sh_sub[b]);
                x = max(x, 0);
                const int x_sq = (int)(((long long)x * x + (long long)a.sp.add_sq[b]) >> a.sp.sh_sq[b]);
                val[b] = ((long long)x_sq * x + (long long)a.sp.add_cub[b]) >> a.sp.sh_cub[b];
            }
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const long long s = bv_warp_sum_redux(val[b]);                                                // 0 <= val < 2^57
            if ((tid & 31) == 0) s_row[((r - 1) * 2 + (c > 32)) * 6 + b] = (unsigned long long)s;
        }
    }
    __syncthreads();
    if (tid < AT_H * 6) {
        const int rr = tid / 6, k = tid - rr * 6;
        const unsigned long long s = s_row[rr * 12 + k] + s_row[rr * 12 + 6 + k];
        s_row[rr * 12 + k] = 0ull; s_row[rr * 12 + 6 + k] = 0ull;
        if (s && ty0 + rr < oh)
            atomicAdd(a.rows + (size_t)f * a.rows_frame_stride + a.rows_offset + (size_t)(ty0 + rr) * 6 + k, s);
    }
    }   // tile loop
}

template <int SCALE> size_t adm_smem()
{
    return sizeof(typename AdmTypes<SCALE>::V) * AP_H * AN_P + sizeof(int) * (AP_H * AP_W + 4 * AT_H * AT_W) +
           sizeof(unsigned long long) * AT_H * 12 + sizeof(typename AdmTypes<SCALE>::Stage) * 2 * AN_R * AN_P;
}

struct AdmFinishArgs {
    BvAdmScaleParams sp[4];
    const unsigned long long *rows;
    size_t rows_frame_stride, rows_offset[4];
    unsigned long long *raw;
};

// libvmaf's per-row rounding: cm += (row_sum + add) >> shift, same for the denominator.
__global__ void __launch_bounds__(128)
adm_rows_finish_kernel(BvBatch batch, AdmFinishArgs a)
{
    __shared__ long long scratch[6 * 32];
    const int scale = blockIdx.x, f = blockIdx.y;
    if (batch.flags[f] & (BV_FRAME_LEAD_IN | BV_FRAME_SKIP_SPATIAL)) return;
    const BvAdmScaleParams &sp = a.sp[scale];
    const unsigned long long *rows = a.rows + (size_t)f * a.rows_frame_stride + a.rows_offset[scale];
    long long acc[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = sp.top + (int)threadIdx.x; i < sp.bottom; i += blockDim.x) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            acc[b] += (long long)((rows[(size_t)i * 6 + b] + sp.add_inner) >> sp.sh_inner);
            acc[3 + b] += (long long)((rows[(size_t)i * 6 + 3 + b] + sp.den_add_row) >> sp.den_sh_row);
        }
    }
    long long cmv[3] = { acc[0], acc[1], acc[2] }, dnv[3] = { acc[3], acc[4], acc[5] };
    bv_block_accumulate<3>(cmv, scratch, a.raw + (size_t)f * BV_RAW_WORDS + BV_RAW_ADM_CM + 3 * scale);
    __syncthreads();
    bv_block_accumulate<3>(dnv, scratch, a.raw + (size_t)f * BV_RAW_WORDS + BV_RAW_ADM_DEN + 3 * scale);
}

template <int SCALE, typename TIn>
void launch_scale(const BvBatch &b, const AdmArgs &a, cudaStream_t st)
{
    const size_t smem = adm_smem<SCALE>();
    bv_allow_smem<&adm_scale_kernel<SCALE, TIn>>(smem);
    const int tiles_x = (a.sp.w + AT_W - 1) / AT_W, tiles_per_frame = tiles_x * ((a.sp.h + AT_H - 1) / AT_H);
    const int total = tiles_per_frame * b.n;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int per_sm = SCALE == 0 ? 3 : 2;
    const int ctas = total < per_sm * sms ? total : per_sm * sms;
    adm_scale_kernel<SCALE, TIn><<<ctas, AT_THREADS, smem, st>>>(b, a, bv_make_div(tiles_x, tiles_per_frame), bv_make_div(tiles_per_frame, total), total);
}

float dwt_quant_step(int lambda, int theta, double view_dist, int display_h)
{
    // Watson et al. DWT quantisation model, evaluated the way libvmaf's adm_tools.h does (float
    // temporaries around double libm calls).
    static const float amp[4][4] = { SPEC_DWT79_AMP };
    static const float g[4] = { SPEC_DWT79_G };
    const float a = SPEC_DWT79_A, k = SPEC_DWT79_K, f0 = SPEC_DWT79_F0;
    float r = view_dist * display_h * M_PI / 180.0;
    float temp = log10(pow(2.0, lambda + 1) * f0 * g[theta] / r);
    float Q = 2.0 * a * pow(10.0, k * temp * temp) / amp[lambda][theta];
    return Q;
}

}  // namespace

void bv_adm_rfactor(int scale, double view_dist, int display_h, float rf[3])
{
    const float f1 = dwt_quant_step(scale, 1, view_dist, display_h);
    const float f2 = dwt_quant_step(scale, 2, view_dist, display_h);
    rf[0] = 1.0f / f1; rf[1] = 1.0f / f1; rf[2] = 1.0f / f2;
}

// Per-scale fixed-point parameters (libvmaf integer_adm.c: i4_adm_cm / adm_csf_den_scale shifts).
void bv_adm_make_params(int w, int h, double view_dist, int display_h, BvAdmScaleParams sp[4])
{
    int cw = w, ch = h;
    for (int s = 0; s < 4; ++s) {
        BvAdmScaleParams &p = sp[s];
        p.in_w = cw; p.in_h = ch;
        cw = (cw + 1) / 2; ch = (ch + 1) / 2;
        p.w = cw; p.h = ch;
        p.left = (int)(cw * SPEC_ADM_BORDER_FACTOR - 0.5); p.top = (int)(ch * SPEC_ADM_BORDER_FACTOR - 0.5);
        p.right = cw - p.left; p.bottom = ch - p.top;
        float rf[3];
        bv_adm_rfactor(s, view_dist, display_h, rf);
        if (s == 0) {
            if (fabs(view_dist * display_h - 3.0 * 1080) < 1.0e-8) {
                static const unsigned s0_rf[3] = { SPEC_ADM_S0_RF };
                p.rf[0] = s0_rf[0]; p.rf[1] = s0_rf[1]; p.rf[2] = s0_rf[2];
            } else {
                p.rf[0] = (uint16_t)(rf[0] * pow(2, 21)); p.rf[1] = (uint16_t)(rf[1] * pow(2, 21));
                p.rf[2] = (uint16_t)(rf[2] * pow(2, 23));
            }
            p.sh_sub[0] = 10; p.sh_sub[1] = 10; p.sh_sub[2] = 12;
            p.sh_sq[0] = 29; p.sh_sq[1] = 29; p.sh_sq[2] = 30;
            p.sh_cub[0] = p.sh_cub[1] = (int)(uint32_t)ceil(log2(cw) - 4);
            p.sh_cub[2] = (int)(uint32_t)ceil(log2(cw) - 3);
        } else {
            for (int b = 0; b < 3; ++b) {
                p.rf[b] = (uint32_t)(rf[b] * pow(2, 32));
                p.sh_sub[b] = 0; p.sh_sq[b] = 30; p.sh_cub[b] = (int)(uint32_t)ceil(log2(cw));
            }
        }
        for (int b = 0; b < 3; ++b) {
            p.add_sq[b] = 1ull << (p.sh_sq[b] - 1);
            p.add_cub[b] = (unsigned long long)(uint32_t)pow(2, (p.sh_cub[b] - 1));
        }
        p.sh_inner = (int)(uint32_t)ceil(log2(ch));
        p.add_inner = (unsigned long long)(uint32_t)pow(2, (p.sh_inner - 1));
        if (s == 0) {
            int sh_acc = (int)ceil(log2((double)(p.bottom - p.top) * (p.right - p.left)) - 20);
            if (sh_acc < 0) sh_acc = 0;
            p.den_sh_row = sh_acc; p.den_add_row = sh_acc > 0 ? (1ull << (sh_acc - 1)) : 0ull;
            p.den_sh_sq = 0; p.den_add_sq = 0; p.den_sh_cub = 0; p.den_add_cub = 0;
        } else {
            static const int sh_sqd[3] = { 31, 30, 31 };
            p.den_sh_sq = sh_sqd[s - 1]; p.den_add_sq = 1ull << (p.den_sh_sq - 1);
            p.den_sh_cub = (int)(uint32_t)ceil(log2(p.right - p.left));
            p.den_add_cub = (unsigned long long)(uint32_t)pow(2, (p.den_sh_cub - 1));
            p.den_sh_row = (int)(uint32_t)ceil(log2(p.bottom - p.top));
            p.den_add_row = (unsigned long long)(uint32_t)pow(2, (p.den_sh_row - 1));
        }
    }
}

// Scalar finalisation of one scale (host): integer accumulators -> float num / den as libvmaf
// stores them (powf on the host libm, float partial sums).
void bv_adm_finish_scale(const BvAdmScaleParams &p, int scale, double view_dist, int display_h,
                         const int64_t cm[3], const uint64_t dn[3], float *num_scale, float *den_scale)
{
    float rf[3];
    bv_adm_rfactor(scale, view_dist, display_h, rf);
    const float area_term = powf((p.bottom - p.top) * (p.right - p.left) / 32.0f, 1.0f / 3.0f);
    float f_acc[3];
    if (scale == 0) {
        f_acc[0] = (float)(cm[0] / pow(2, (52 - p.sh_cub[0] - p.sh_inner)));
        f_acc[1] = (float)(cm[1] / pow(2, (52 - p.sh_cub[1] - p.sh_inner)));
        f_acc[2] = (float)(cm[2] / pow(2, (57 - p.sh_cub[2] - p.sh_inner)));
    } else {
        static const int fs[3] = { 45, 39, 36 };
        const float final_shift = pow(2, (fs[scale - 1] - p.sh_cub[0] - p.sh_inner));
        for (int b = 0; b < 3; ++b) f_acc[b] = (float)(cm[b] / final_shift);
    }
    const float nh = powf(f_acc[0], 1.0f / 3.0f) + area_term;
    const float nv = powf(f_acc[1], 1.0f / 3.0f) + area_term;
    const float nd = powf(f_acc[2], 1.0f / 3.0f) + area_term;
    *num_scale = nh + nv + nd;

    double shift_csf;
    if (scale == 0) {
        shift_csf = pow(2, (18 - p.den_sh_row));
    } else {
        static const int conv[3] = { 32, 27, 23 };
        shift_csf = pow(2, (conv[scale - 1] - p.den_sh_row - p.den_sh_cub));
    }
    float part[3];
    for (int b = 0; b < 3; ++b) {
        const double csf = (double)(dn[b] / shift_csf) * pow(rf[b], 3);
        part[b] = powf(csf, 1.0f / 3.0f) + area_term;
    }
    *den_scale = part[0] + part[1] + part[2];
}

void bv_launch_adm(const BvBatch &b, BvPlane ref_y, BvPlane dis_y, int bpc, const BvAdmBuffers &ab,
                   const BvAdmScaleParams sp[4], double egl, unsigned long long *raw, const BvLaunch &L)
{
    const float cos_1deg_sq = (float)(cos(1.0 * M_PI / 180.0) * cos(1.0 * M_PI / 180.0));
    BvPlane cr = ref_y, cd = dis_y;
    cudaStream_t st = L.st;
    for (int s = 0; s < 4; ++s) {
        AdmArgs a;
        a.ref = cr; a.dis = cd;
        a.a_frame_elems = ab.band_plane_elems[s];
        const size_t esz = s == 0 ? 2 : 4;
        a.a_ref = s < 3 ? ab.bands[s] : nullptr;
        a.a_dis = s < 3 ? static_cast<uint8_t *>(ab.bands[s]) + (size_t)BV_MAX_BATCH * ab.band_plane_elems[s] * esz : nullptr;
        a.sp = sp[s]; a.bpc = bpc; a.div_lookup = ab.div_lookup; a.egl = egl; a.cos_1deg_sq = cos_1deg_sq;
        a.rows = ab.rows; a.rows_frame_stride = ab.rows_frame_stride; a.rows_offset = ab.rows_scale_offset[s];
        {
            const size_t in_sz = s == 0 ? (bpc == 8 ? 1 : 2) : (s == 1 ? 2 : 4);
            size_t bits = cr.pitch | cd.pitch;
            for (int k = 0; k < b.n; ++k) bits |= (size_t)cr.p[k] | (size_t)cd.p[k];
            a.vec_ok = (bits & (4 * in_sz - 1)) == 0;
        }
        bv_prof_begin(L, BVK_ADM_S0 + s);
        switch (s) {
        case 0:
            if (bpc == 8) launch_scale<0, uint8_t>(b, a, st); else launch_scale<0, uint16_t>(b, a, st);
            break;
        case 1: launch_scale<1, int16_t>(b, a, st); break;
        case 2: launch_scale<2, int32_t>(b, a, st); break;
        default: launch_scale<3, int32_t>(b, a, st); break;
        }
        bv_prof_end(L, BVK_ADM_S0 + s);
        if (s < 3) {
            cr = bv_plane_contig(a.a_ref, (size_t)sp[s].w * esz, ab.band_plane_elems[s] * esz, b.n);
            cd = bv_plane_contig(a.a_dis, (size_t)sp[s].w * esz, ab.band_plane_elems[s] * esz, b.n);
        }
    }
    AdmFinishArgs fa;
    for (int s = 0; s < 4; ++s) { fa.sp[s] = sp[s]; fa.rows_offset[s] = ab.rows_scale_offset[s]; }
    fa.rows = ab.rows; fa.rows_frame_stride = ab.rows_frame_stride; fa.raw = raw;
    bv_prof_begin(L, BVK_ADM_FINISH);
    adm_rows_finish_kernel<<<dim3(4, b.n), 128, 0, st>>>(b, fa);
    bv_prof_end(L, BVK_ADM_FINISH);
}
