// psnr_sse / svr_predict -- the two small kernels of the path.
//
// psnr_sse: per-plane sum of squared differences (libvmaf integer_psnr.c `psnr=1`, reference
// app/vmaf_analyzer.py:385; FFmpeg vf_psnr.c, reference app/vmaf_analyzer.py:1032).  Pure streaming
// read: 16-byte vector loads per lane, dp4a on byte |a-b| for 8-bit, 64-bit integer accumulation.
//
// svr_predict: libsvm nu-SVR / RBF decision value with libvmaf predict.c's linear feature rescale
// (SURVEY.md Appendix A.6), one CTA per frame, one thread per support vector, fixed-order tree sum.
#include "bv_common.cuh"
#include "../../include/b200vmaf.h"
#include "../../include/libvmaf_spec.h"

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
sse_kernel(BvBatch batch, BvPlane ref, BvPlane dis, int w, int h, unsigned long long *raw, int raw_idx)
{
    __shared__ long long scratch[32];
    const int f = blockIdx.y;
    if (batch.flags[f] & (BV_FRAME_LEAD_IN | BV_FRAME_SKIP_SPATIAL)) return;
    const uint8_t *pr = ref.p[f], *pd = dis.p[f];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t row_bytes = (size_t)w * sizeof(T);
    const bool vec_ok = (((uintptr_t)pr | (uintptr_t)pd | ref.pitch | dis.pitch) & 15) == 0;
    unsigned long long acc = 0;
    for (int row = blockIdx.x * nwarps + warp; row < h; row += gridDim.x * nwarps) {
        const uint8_t *a = pr + (size_t)row * ref.pitch, *b = pd + (size_t)row * dis.pitch;
        size_t done = 0;
        if (vec_ok) {
            const size_t nv = row_bytes / 16;
            // 4 x 2 independent 16-byte loads in flight per lane (HBM-bound: bytes in flight are what matters)
            constexpr int U = 4;
            for (size_t v0 = lane; v0 < nv; v0 += 32 * U) {
                uint4 xv[U], yv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const size_t v = v0 + 32 * u;
                    if (v < nv) {
                        xv[u] = __ldg(reinterpret_cast<const uint4 *>(a) + v);
                        yv[u] = __ldg(reinterpret_cast<const uint4 *>(b) + v);
                    } else { xv[u] = make_uint4(0, 0, 0, 0); yv[u] = xv[u]; }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned xs[4] = { xv[u].x, xv[u].y, xv[u].z, xv[u].w }, ys[4] = { yv[u].x, yv[u].y, yv[u].z, yv[u].w };
                    if (sizeof(T) == 1) {
                        unsigned s = 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const unsigned d = __vabsdiffu4(xs[q], ys[q]);
                            s = __dp4a(d, d, s);
                        }
                        acc += s;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const unsigned d = __vabsdiffu2(xs[q], ys[q]);
                            const unsigned long long lo = d & 0xffffu, hi = d >> 16;
                            acc += lo * lo + hi * hi;
                        }
                    }
                }
            }
            done = nv * 16 / sizeof(T);
        }
        for (size_t j = done + lane; j < (size_t)w; j += 32) {
            const long long d = (long long)reinterpret_cast<const T *>(a)[j] - (long long)reinterpret_cast<const T *>(b)[j];
            acc += (unsigned long long)(d * d);
        }
    }
    long long v[1] = { (long long)acc };
    bv_block_accumulate<1>(v, scratch, raw + (size_t)f * BV_RAW_WORDS + raw_idx);
}

__global__ void __launch_bounds__(256)
svr_predict_kernel(const double *__restrict__ feat, int n_feat, const double *__restrict__ slopes,
                   const double *__restrict__ intercepts, const double *__restrict__ sv,
                   const double *__restrict__ coef, int n_sv, double gamma, double rho, double *__restrict__ out)
{
    __shared__ double s_x[64];
    __shared__ double s_part[256];
    const int frame = blockIdx.x, tid = threadIdx.x;
    if (tid < n_feat) s_x[tid] = slopes[tid + 1] * feat[(size_t)frame * n_feat + tid] + intercepts[tid + 1];
    __syncthreads();
    double sum = 0.0;
    for (int k = tid; k < n_sv; k += 256) {
        double d2 = 0.0;
        for (int i = 0; i < n_feat; ++i) {
            const double d = s_x[i] - sv[(size_t)k * n_feat + i];
            d2 = __dadd_rn(d2, __dmul_rn(d, d));
        }
        sum += coef[k] * exp(-gamma * d2);
    }
    s_part[tid] = sum;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) s_part[tid] += s_part[tid + s];
        __syncthreads();
    }
    if (tid == 0) out[frame] = ((s_part[0] - rho) - intercepts[0]) / slopes[0];
}

// ---- FFmpeg `ssim` filter (libavfilter/vf_ssim.c; reference app/vmaf_analyzer.py:1057-1064) ----
// x264-style integer SSIM: 4x4 block sums (s1, s2, ss, s12), 8x8 windows at stride 4 built from four
// blocks, ssim_end1 in float.  CTA = 32x8 windows = 33x9 blocks = 132x36 pixels staged in shared
// memory.  Output: the sum over windows (double) -> fraw[BV_FRAW_FFSSIM + plane].
constexpr int FS_WX = 32, FS_WY = 8, FS_BX = FS_WX + 1, FS_BY = FS_WY + 1, FS_PW = 4 * FS_BX, FS_PH = 4 * FS_BY;

template <typename T, typename S>
__global__ void __launch_bounds__(256)
ffssim_kernel(BvBatch batch, BvPlane ref, BvPlane dis, int bpc, int w, int h, double *fraw, int fraw_idx, int fraw_words)
{
    __shared__ T s_a[FS_PH][FS_PW + 4];
    __shared__ T s_b[FS_PH][FS_PW + 4];
    __shared__ S s_blk[FS_BY][FS_BX][4];
    __shared__ double scratch[32];
    const int f = blockIdx.z;
    if (batch.flags[f] & (BV_FRAME_LEAD_IN | BV_FRAME_SKIP_SPATIAL)) return;
    const int W4 = w >> 2, H4 = h >> 2;
    const int bx0 = blockIdx.x * FS_WX, by0 = blockIdx.y * FS_WY;
    const uint8_t *pa = ref.p[f], *pb = dis.p[f];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < FS_PH * FS_PW; idx += 256) {
        const int r = idx / FS_PW, c = idx - r * FS_PW;
        const int gy = min(by0 * 4 + r, h - 1), gx = min(bx0 * 4 + c, w - 1);
        s_a[r][c] = (T)bv_ld<T>(pa, ref.pitch, gy, gx);
        s_b[r][c] = (T)bv_ld<T>(pb, dis.pitch, gy, gx);
    }
    __syncthreads();
    for (int idx = tid; idx < FS_BY * FS_BX; idx += 256) {
        const int r = idx / FS_BX, c = idx - r * FS_BX;
        S s1 = 0, s2 = 0, ss = 0, s12 = 0;
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const S p = (S)s_a[4 * r + y][4 * c + x], q = (S)s_b[4 * r + y][4 * c + x];
                s1 += p; s2 += q; ss += p * p + q * q; s12 += p * q;
            }
        s_blk[r][c][0] = s1; s_blk[r][c][1] = s2; s_blk[r][c][2] = ss; s_blk[r][c][3] = s12;
    }
    __syncthreads();
    double acc = 0.0;
    {
        const int c = tid & 31, r = tid >> 5;
        if (bx0 + c < W4 - 1 && by0 + r < H4 - 1) {
            S s1 = 0, s2 = 0, ss = 0, s12 = 0;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    s1 += s_blk[r + dy][c + dx][0]; s2 += s_blk[r + dy][c + dx][1];
                    ss += s_blk[r + dy][c + dx][2]; s12 += s_blk[r + dy][c + dx][3];
                }
            S c1, c2;
            if (sizeof(T) == 1) { c1 = SPEC_FFSSIM_C1; c2 = SPEC_FFSSIM_C2; }
            else {
                const int maxv = (1 << bpc) - 1;
                c1 = (S)(long long)(.01 * .01 * maxv * maxv * 64 + .5);
                c2 = (S)(long long)(.03 * .03 * maxv * maxv * 64 * 63 + .5);
            }
            const S vars = ss * 64 - s1 * s1 - s2 * s2;
            const S covar = s12 * 64 - s1 * s2;
            const float v = __fdiv_rn(__fmul_rn((float)(2 * s1 * s2 + c1), (float)(2 * covar + c2)),
                                      __fmul_rn((float)(s1 * s1 + s2 * s2 + c1), (float)(vars + c2)));
            acc = (double)v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) scratch[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += scratch[k];
        if (s != 0.0) atomicAdd(fraw + (size_t)f * fraw_words + fraw_idx, s);
    }
}

// ---- luma statistics for bookend (white-frame) detection ------------------------------------------
// Replaces the per-frame cv2 scan of the reference's alignment step (app/bookend_alignment.py:997-1020,
// app/reference_analyzer.py:131-141): mean, standard deviation and the share of pixels above a
// threshold.  Exact integer sums per frame: sum y, sum y^2, count(y > thr[k]) for 3 thresholds.
// Pure streaming read (16-byte loads, one pass) -> HBM-bound.
template <typename T>
__global__ void __launch_bounds__(256)
luma_stats_kernel(const uint8_t *base, size_t pitch, size_t frame_stride, int w, int h, uint3 thr,
                  unsigned long long *out /* [frame][5] */)
{
    __shared__ long long scratch[5 * 32];
    const int f = blockIdx.y;
    const uint8_t *img = base + (size_t)f * frame_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t row_bytes = (size_t)w * sizeof(T);
    const bool vec_ok = (((uintptr_t)img | pitch) & 15) == 0;
    unsigned long long s1 = 0, s2 = 0;
    unsigned c0 = 0, c1 = 0, c2 = 0;
    auto px = [&](unsigned v) {
        s1 += v; s2 += (unsigned long long)v * v;
        c0 += v > thr.x; c1 += v > thr.y; c2 += v > thr.z;
    };
    for (int row = blockIdx.x * nwarps + warp; row < h; row += gridDim.x * nwarps) {
        const uint8_t *a = img + (size_t)row * pitch;
        size_t done = 0;
        if (vec_ok) {
            const size_t nv = row_bytes / 16;
            for (size_t v = lane; v < nv; v += 32) {
                const uint4 x = __ldg(reinterpret_cast<const uint4 *>(a) + v);
                const unsigned xs[4] = { x.x, x.y, x.z, x.w };
                if (sizeof(T) == 1) {
                    unsigned q1 = 0, q2 = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        q1 = __dp4a(xs[q], 0x01010101u, q1);
                        q2 = __dp4a(xs[q], xs[q], q2);
                        // per-byte compare: __vcmpgtu4 gives 0xff per true byte; popcount / 8 = number of true bytes
                        c0 += __popc(__vcmpgtu4(xs[q], thr.x * 0x01010101u)) >> 3;
                        c1 += __popc(__vcmpgtu4(xs[q], thr.y * 0x01010101u)) >> 3;
                        c2 += __popc(__vcmpgtu4(xs[q], thr.z * 0x01010101u)) >> 3;
                    }
                    s1 += q1; s2 += q2;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) { px(xs[q] & 0xffffu); px(xs[q] >> 16); }
                }
            }
            done = nv * 16 / sizeof(T);
        }
        for (size_t j = done + lane; j < (size_t)w; j += 32) px((unsigned)reinterpret_cast<const T *>(a)[j]);
    }
    long long v[5] = { (long long)s1, (long long)s2, (long long)c0, (long long)c1, (long long)c2 };
    bv_block_accumulate<5>(v, scratch, out + (size_t)f * 5);
}

}  // namespace

extern "C" int bv_luma_stats_device(int device, const void *d_luma, size_t pitch, size_t frame_stride, int n_frames,
                                    int w, int h, int bpc, const unsigned thr[3], unsigned long long *out_host)
{
    if (!d_luma || !thr || !out_host || n_frames <= 0 || w <= 0 || h <= 0) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    unsigned long long *d_out = nullptr;
    if (cudaMalloc(&d_out, sizeof(unsigned long long) * 5 * n_frames) != cudaSuccess) return -2;
    cudaMemset(d_out, 0, sizeof(unsigned long long) * 5 * n_frames);
    int gx = (h + 7) / 8;
    if (gx > 148 * 2) gx = 148 * 2;
    // the 8-bit compare works on bytes: thresholds above 254 can never be exceeded
    uint3 t = make_uint3(thr[0], thr[1], thr[2]);
    if (bpc == 8) { t.x = t.x > 255 ? 255 : t.x; t.y = t.y > 255 ? 255 : t.y; t.z = t.z > 255 ? 255 : t.z; }
    dim3 grid(gx, n_frames);
    if (bpc == 8) luma_stats_kernel<uint8_t><<<grid, 256>>>(static_cast<const uint8_t *>(d_luma), pitch, frame_stride, w, h, t, d_out);
    else luma_stats_kernel<uint16_t><<<grid, 256>>>(static_cast<const uint8_t *>(d_luma), pitch, frame_stride, w, h, t, d_out);
    cudaError_t e = cudaMemcpy(out_host, d_out, sizeof(unsigned long long) * 5 * n_frames, cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    return e == cudaSuccess ? 0 : -2;
}

namespace {
}

void bv_launch_ffssim(const BvBatch &b, BvPlane ref, BvPlane dis, int bpc, int w, int h, int plane_idx,
                      double *fraw, int fraw_idx, int fraw_words, const BvLaunch &L)
{
    const int W4 = w >> 2, H4 = h >> 2;
    if (W4 < 2 || H4 < 2) return;
    dim3 grid((W4 - 1 + FS_WX - 1) / FS_WX, (H4 - 1 + FS_WY - 1) / FS_WY, b.n);
    bv_prof_begin(L, BVK_FFSSIM_Y + plane_idx);
    if (bpc == 8) ffssim_kernel<uint8_t, int><<<grid, 256, 0, L.st>>>(b, ref, dis, bpc, w, h, fraw, fraw_idx, fraw_words);
    else ffssim_kernel<uint16_t, long long><<<grid, 256, 0, L.st>>>(b, ref, dis, bpc, w, h, fraw, fraw_idx, fraw_words);
    bv_prof_end(L, BVK_FFSSIM_Y + plane_idx);
}

void bv_launch_sse(const BvBatch &b, BvPlane ref, BvPlane dis, int bpc, int w, int h, int plane_idx,
                   unsigned long long *raw, const BvLaunch &L)
{
    // one wave of 8 CTAs per SM over the whole group (a warp streams whole rows)
    int gx = (h + 7) / 8;
    const int wave = (148 * 8 + b.n - 1) / b.n;
    if (gx > wave) gx = wave;
    dim3 grid(gx, b.n);
    bv_prof_begin(L, BVK_SSE_Y + plane_idx);
    if (bpc == 8) sse_kernel<uint8_t><<<grid, 256, 0, L.st>>>(b, ref, dis, w, h, raw, BV_RAW_SSE + plane_idx);
    else sse_kernel<uint16_t><<<grid, 256, 0, L.st>>>(b, ref, dis, w, h, raw, BV_RAW_SSE + plane_idx);
    bv_prof_end(L, BVK_SSE_Y + plane_idx);
}

void bv_launch_svr(const double *d_feat, int n_feat, const double *d_slopes, const double *d_intercepts,
                   const double *d_sv, const double *d_coef, int n_sv, double gamma, double rho, double *d_out,
                   long long n, cudaStream_t st)
{
    svr_predict_kernel<<<(unsigned)n, 256, 0, st>>>(d_feat, n_feat, d_slopes, d_intercepts, d_sv, d_coef, n_sv,
                                                   gamma, rho, d_out);
}
