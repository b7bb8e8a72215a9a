// motion_blur / motion_sad -- integer motion feature (replaces libvmaf integer_motion.c, reached
// from the reference at app/vmaf_analyzer.py:417; algorithm: SURVEY.md Appendix A.3).
//
// 5-tap Q16 blur {3571,16004,26386,16004,3571}, vertical then horizontal, MIRROR borders,
// u16 output; SAD against the previous frame's blurred picture.  HBM-bound:
// reads W*H*bytes (luma) + writes W*H*2 (blur) + SAD reads 2*W*H*2.
#include "bv_common.cuh"
#include "../../include/b200vmaf.h"
#include "../../include/libvmaf_spec.h"

namespace {

// Tile = 128 x 32 outputs; the staged window starts 4 columns left of the tile so that 4-pixel vector
// loads stay aligned.  Register-blocked: 8 outputs per item in both passes (12 inputs in registers).
constexpr int MB_TW = 128, MB_TH = 32, MB_R = 2;
constexpr int MB_IN_W = MB_TW + 8, MB_IN_H = MB_TH + 2 * MB_R, MB_G = MB_IN_W / 4, MB_O = 8;
constexpr int MB_P = MB_IN_W;            // u16 pitch of the staged tile: 8-byte rows (one 8-byte store per group; only read column-per-lane)
constexpr int MB_VP = MB_IN_W + 1;       // odd u32 pitch of the vertical-pass plane

__constant__ unsigned c_motion_filter[5] = { SPEC_MOTION_Q16_5 };

template <typename T>
__global__ void __launch_bounds__(256)
motion_blur_kernel(BvBatch batch, BvPlane src, int bpc, int w, int h, uint16_t *__restrict__ blur,
                   size_t blur_frame_elems, int vec_ok)
{
    __shared__ __align__(16) uint16_t s_in[MB_IN_H * MB_P];
    __shared__ unsigned s_v[MB_TH * MB_VP];

    const int f = blockIdx.z;
    const uint8_t *img = src.p[f];
    const int x0 = blockIdx.x * MB_TW - 4, y0 = blockIdx.y * MB_TH - MB_R;
    const int tid = threadIdx.x;

    using V4 = typename Px4<T>::V;
    bv_stage_tile<MB_IN_H, MB_G, V4>(tid,
        [&](int r, int gc) {
            const int gy = bv_mirror(min(y0 + r, h + MB_R - 1), h);
            return load_px4<T>(img + (size_t)gy * src.pitch, x0 + 4 * gc, w, w + MB_R - 1, vec_ok);
        },
        [&](int r, int gc, V4 v) {
            unsigned u[4];
            Px4<T>::raw(v, u);
            *reinterpret_cast<uint2 *>(s_in + r * MB_P + 4 * gc) = make_uint2(u[0] | (u[1] << 16), u[2] | (u[3] << 16));
        });
    __syncthreads();

    const unsigned add_v = 1u << (bpc - 1);
    for (int item = tid; item < MB_IN_W * (MB_TH / MB_O); item += 256) {
        const int c = item % MB_IN_W, strip = item / MB_IN_W;
        unsigned v[MB_O + 4];
#pragma unroll
        for (int i = 0; i < MB_O + 4; ++i) v[i] = s_in[(MB_O * strip + i) * MB_P + c];
#pragma unroll
        for (int o = 0; o < MB_O; ++o) {
            unsigned acc = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) acc += c_motion_filter[k] * v[o + k];
            s_v[(MB_O * strip + o) * MB_VP + c] = (acc + add_v) >> bpc;
        }
    }
    __syncthreads();

    uint16_t *out = blur + (size_t)f * blur_frame_elems;
    for (int item = tid; item < MB_TH * (MB_TW / MB_O); item += 256) {
        const int r = item % MB_TH, g = item / MB_TH;
        unsigned v[MB_O + 4];
#pragma unroll
        for (int i = 0; i < MB_O + 4; ++i) v[i] = s_v[r * MB_VP + MB_O * g + 2 + i];     // output col j <-> staged col j + 4
        const int gy = y0 + MB_R + r, gx0 = x0 + 4 + MB_O * g;
        if (gy >= h) continue;
        unsigned res[MB_O];
#pragma unroll
        for (int o = 0; o < MB_O; ++o) {
            unsigned acc = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) acc += c_motion_filter[k] * v[o + k];
            res[o] = (acc + 32768u) >> 16;
        }
        uint16_t *dst = out + (size_t)gy * w + gx0;
        if ((w & 7) == 0 && (blur_frame_elems & 7) == 0 && gx0 + MB_O <= w) {
            *reinterpret_cast<uint4 *>(dst) = make_uint4(res[0] | (res[1] << 16), res[2] | (res[3] << 16),
                                                         res[4] | (res[5] << 16), res[6] | (res[7] << 16));
        } else {
#pragma unroll
            for (int o = 0; o < MB_O; ++o)
                if (gx0 + o < w) dst[o] = (uint16_t)res[o];
        }
    }
}

// SAD of consecutive blurred frames.  Frame 0 of the group pairs with the last frame of the
// previous group (prev_last); frames flagged BV_FRAME_FIRST have no predecessor (sad = 0).
__global__ void __launch_bounds__(256)
motion_sad_kernel(BvBatch batch, const uint16_t *__restrict__ blur, const uint16_t *__restrict__ prev_last,
                  size_t frame_elems, size_t n_elems, unsigned long long *raw)
{
    __shared__ long long scratch[32];
    const int f = blockIdx.y;
    if (batch.flags[f] & BV_FRAME_FIRST) return;
    const uint16_t *cur = blur + (size_t)f * frame_elems;
    const uint16_t *prv = f == 0 ? prev_last : blur + (size_t)(f - 1) * frame_elems;
    unsigned long long sad = 0;
    const size_t n8 = n_elems / 8;
    const uint4 *c4 = reinterpret_cast<const uint4 *>(cur);
    const uint4 *p4 = reinterpret_cast<const uint4 *>(prv);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 a = __ldg(c4 + i), b = __ldg(p4 + i);
        unsigned s = 0;
        // per-halfword absolute differences, summed (max 8 * 65535 fits easily)
        s += __vsadu2(a.x, b.x);
        s += __vsadu2(a.y, b.y);
        s += __vsadu2(a.z, b.z);
        s += __vsadu2(a.w, b.w);
        sad += s;
    }
    if (blockIdx.x == 0) {
        for (size_t i = n8 * 8 + threadIdx.x; i < n_elems; i += blockDim.x)
            sad += (unsigned)abs((int)cur[i] - (int)prv[i]);
    }
    long long v[1] = { (long long)sad };
    bv_block_accumulate<1>(v, scratch, raw + (size_t)f * BV_RAW_WORDS + BV_RAW_SAD);
}

}  // namespace

void bv_launch_motion_blur(const BvBatch &b, BvPlane ref_y, int bpc, int w, int h, uint16_t *blur_cur,
                           size_t blur_frame_elems, const BvLaunch &L)
{
    dim3 grid((w + MB_TW - 1) / MB_TW, (h + MB_TH - 1) / MB_TH, b.n);
    bv_prof_begin(L, BVK_MOTION_BLUR);
    size_t bits = ref_y.pitch;
    for (int k = 0; k < b.n; ++k) bits |= (size_t)ref_y.p[k];
    const int vec_ok = (bits & (bpc == 8 ? 3 : 7)) == 0;
    if (bpc == 8)
        motion_blur_kernel<uint8_t><<<grid, 256, 0, L.st>>>(b, ref_y, bpc, w, h, blur_cur, blur_frame_elems, vec_ok);
    else
        motion_blur_kernel<uint16_t><<<grid, 256, 0, L.st>>>(b, ref_y, bpc, w, h, blur_cur, blur_frame_elems, vec_ok);
    bv_prof_end(L, BVK_MOTION_BLUR);
}

void bv_launch_motion_sad(const BvBatch &b, const uint16_t *blur_cur, const uint16_t *blur_prev_group_last,
                          size_t blur_frame_elems, int w, int h, unsigned long long *raw, const BvLaunch &L)
{
    const size_t n = (size_t)w * h;
    int gx = (int)((n / 8 + 255) / 256);
    if (gx > 148 * 2) gx = 148 * 2;
    if (gx < 1) gx = 1;
    dim3 grid(gx, b.n);
    bv_prof_begin(L, BVK_MOTION_SAD);
    motion_sad_kernel<<<grid, 256, 0, L.st>>>(b, blur_cur, blur_prev_group_last, blur_frame_elems, n, raw);
    bv_prof_end(L, BVK_MOTION_SAD);
}
