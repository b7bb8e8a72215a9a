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
            for (size_t v = lane; v < nv; v += 32) {
                const uint4 x = __ldg(reinterpret_cast<const uint4 *>(a) + v);
                const uint4 y = __ldg(reinterpret_cast<const uint4 *>(b) + v);
                const unsigned xs[4] = { x.x, x.y, x.z, x.w }, ys[4] = { y.x, y.y, y.z, y.w };
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

}  // namespace

void bv_launch_sse(const BvBatch &b, BvPlane ref, BvPlane dis, int bpc, int w, int h, int plane_idx,
                   unsigned long long *raw, const BvLaunch &L)
{
    int gx = (h + 7) / 8;
    if (gx > 148) gx = 148;
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
