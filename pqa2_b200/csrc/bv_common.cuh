// Shared device helpers and the internal kernel-launch interface of libb200vmaf (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define BV_MAX_BATCH 32

// ---- border rules -------------------------------------------------------------------------
// libvmaf "MIRROR" (motion, ADM DWT, CM neighbourhood): -i -> i ; n+i -> n-1-i
__device__ __forceinline__ int bv_mirror(int i, int n)
{
    if (i < 0) return -i;
    if (i >= n) return 2 * n - i - 1;
    return i;
}
// integer VIF padding: reflect-101 on all four sides: -i -> i ; n-1+i -> n-1-i
__device__ __forceinline__ int bv_reflect101(int i, int n)
{
    if (i < 0) return -i;
    if (i >= n) return 2 * (n - 1) - i;
    return i;
}

// KBND_SYMMETRIC of the iqa convolutions (float_ssim / float_ms_ssim): -1 -> 0 ; n -> n-1
__device__ __forceinline__ int bv_sym(int i, int n)
{
    if (i < 0) return -1 - i;
    if (i >= n) return 2 * n - i - 1;
    return i;
}

// ---- opt-in to > 48 KB of dynamic shared memory -------------------------------------------------
// cudaFuncSetAttribute applies to the CURRENT device only, and one process may drive several GPUs (one context and one
// host thread per device): remember the opt-in per (kernel instantiation, device), not per process.
template <auto Kernel>
inline void bv_allow_smem(size_t bytes)
{
    static std::atomic<unsigned long long> done{ 0ull };
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(done.load(std::memory_order_acquire) & bit)) {
        cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        done.fetch_or(bit, std::memory_order_release);
    }
}

// ---- division by a launch-invariant divisor -------------------------------------------------
// The persistent kernels turn a linear work-item index into (frame, tile row, tile column) once per tile and once per
// prefetch; with run-time divisors each `/` was a ~20-instruction sequence (ncu: 3.6 % of vif_stat_s0).  magic =
// ceil(2^32 / d) makes n / d = umulhi(n, magic) exact whenever n * d < 2^32 (the host checks that for the largest
// dividend; otherwise, and for d == 1, magic is 0 and the ordinary division is used).
struct BvDiv {
    int d;
    unsigned magic;
    __host__ __device__ operator int() const { return d; }
};
inline BvDiv bv_make_div(int d, long long max_n)
{
    BvDiv v;
    v.d = d;
    v.magic = (d > 1 && max_n * (long long)d < (1ll << 32)) ? (unsigned)(((1ull << 32) + (unsigned)d - 1) / (unsigned)d) : 0u;
    return v;
}
__device__ __forceinline__ int operator/(int n, const BvDiv &dv)
{
    return dv.magic ? (int)__umulhi((unsigned)n, dv.magic) : n / dv.d;
}
__device__ __forceinline__ int operator%(int n, const BvDiv &dv) { return n - (n / dv) * dv.d; }

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ long long bv_warp_sum(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long bv_warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// 64-bit warp sum through three 32-bit hardware reductions (REDUX instead of 10 SHFL + 10 adds):
// v = c2 * 2^44 + c1 * 2^22 + c0 with 22-bit c0, c1 and a signed c2; every chunk sum fits 32 bits for |v| < 2^62.
__device__ __forceinline__ long long bv_warp_sum_redux(long long v)
{
    const unsigned c0 = (unsigned)v & 0x3fffffu, c1 = (unsigned)(v >> 22) & 0x3fffffu;
    const int c2 = (int)(v >> 44);
    const unsigned s0 = __reduce_add_sync(0xffffffffu, c0), s1 = __reduce_add_sync(0xffffffffu, c1);
    const int s2 = __reduce_add_sync(0xffffffffu, c2);
    return (long long)s0 + ((long long)s1 << 22) + ((long long)s2 << 44);
}
__device__ __forceinline__ int bv_warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum N per-thread 64-bit values over the block and atomically add them to dst[0..N).
// scratch: N * 32 long longs of shared memory.
// REDUX: the caller guarantees |v| < 2^62 for every partial sum (then the warp sums use bv_warp_sum_redux).
template <int N, bool REDUX = false>
__device__ __forceinline__ void bv_block_accumulate(const long long (&v)[N], long long *scratch,
                                                    unsigned long long *dst)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        long long s = REDUX ? bv_warp_sum_redux(v[k]) : bv_warp_sum(v[k]);
        if (lane == 0) scratch[k * 32 + warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            long long s = lane < nwarps ? scratch[k * 32 + lane] : 0;
            s = REDUX ? bv_warp_sum_redux(s) : bv_warp_sum(s);
            if (lane == 0 && s != 0) atomicAdd(dst + k, (unsigned long long)s);
        }
    }
}

// ---- pixel loads --------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ unsigned bv_ld(const uint8_t *base, size_t pitch, int i, int j)
{
    return (unsigned)__ldg(reinterpret_cast<const T *>(base + (size_t)i * pitch) + j);
}

// ---- 4-pixel groups: the unit of the tile prefetch ---------------------------------------------
template <typename T> struct Px4;
template <> struct Px4<uint8_t> {
    using V = unsigned;
    static __device__ __forceinline__ V pack(unsigned a, unsigned b, unsigned c, unsigned d) { return a | (b << 8) | (c << 16) | (d << 24); }
    static __device__ __forceinline__ void unpack(V v, float s, float o, float (&f)[4])
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) f[k] = fmaf((float)((v >> (8 * k)) & 0xffu), s, o);
    }
    static __device__ __forceinline__ void raw(V v, unsigned (&u)[4])
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = (v >> (8 * k)) & 0xffu;
    }
};
template <> struct Px4<uint16_t> {
    using V = uint2;
    static __device__ __forceinline__ V pack(unsigned a, unsigned b, unsigned c, unsigned d) { return make_uint2(a | (b << 16), c | (d << 16)); }
    static __device__ __forceinline__ void unpack(V v, float s, float o, float (&f)[4])
    {
        f[0] = fmaf((float)(v.x & 0xffffu), s, o); f[1] = fmaf((float)(v.x >> 16), s, o);
        f[2] = fmaf((float)(v.y & 0xffffu), s, o); f[3] = fmaf((float)(v.y >> 16), s, o);
    }
    static __device__ __forceinline__ void raw(V v, unsigned (&u)[4])
    {
        u[0] = v.x & 0xffffu; u[1] = v.x >> 16; u[2] = v.y & 0xffffu; u[3] = v.y >> 16;
    }
};
template <> struct Px4<float> {
    using V = float4;
    static __device__ __forceinline__ V pack(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
    static __device__ __forceinline__ void unpack(V v, float s, float o, float (&f)[4])
    {
        f[0] = fmaf(v.x, s, o); f[1] = fmaf(v.y, s, o); f[2] = fmaf(v.z, s, o); f[3] = fmaf(v.w, s, o);
    }
};

template <> struct Px4<int16_t> {
    using V = uint2;
    static __device__ __forceinline__ V pack(int a, int b, int c, int d)
    {
        return make_uint2(((unsigned)a & 0xffffu) | ((unsigned)b << 16), ((unsigned)c & 0xffffu) | ((unsigned)d << 16));
    }
    static __device__ __forceinline__ void raw(V v, int (&u)[4])
    {
        u[0] = (int)(short)(v.x & 0xffffu); u[1] = (int)v.x >> 16; u[2] = (int)(short)(v.y & 0xffffu); u[3] = (int)v.y >> 16;
    }
};
template <> struct Px4<int32_t> {
    using V = int4;
    static __device__ __forceinline__ V pack(int a, int b, int c, int d) { return make_int4(a, b, c, d); }
    static __device__ __forceinline__ void raw(V v, int (&u)[4]) { u[0] = v.x; u[1] = v.y; u[2] = v.z; u[3] = v.w; }
};

// Loads 4 consecutive pixels of row `row` starting at column gx0 (may hang over either image edge:
// resolved per element then, BORDER 0 = libvmaf MIRROR, 1 = reflect-101, 2 = iqa SYMMETRIC).  One vector load when the
// group is interior and aligned.
template <typename T, int BORDER = 0>
__device__ __forceinline__ typename Px4<T>::V load_px4(const uint8_t *row, int gx0, int w, int far, bool vec)
{
    if (vec && gx0 >= 0 && gx0 + 3 < w)
        return __ldg(reinterpret_cast<const typename Px4<T>::V *>(row + (size_t)gx0 * sizeof(T)));
    const T *p = reinterpret_cast<const T *>(row);
    const int lo = -(w - 1);
    int ix[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int g = min(max(gx0 + k, lo), far);
        ix[k] = BORDER == 0 ? bv_mirror(g, w) : (BORDER == 1 ? bv_reflect101(g, w) : bv_sym(g, w));
    }
    return Px4<T>::pack(__ldg(p + ix[0]), __ldg(p + ix[1]), __ldg(p + ix[2]), __ldg(p + ix[3]));
}

// Tile staging for the one-tile-per-CTA pyramid kernels (256 threads): a thread keeps its 4-sample column group and
// walks down the rows, RPP rows per pass, so the column arithmetic is done once, the row loop has a fixed trip count and
// all of a thread's loads are in flight before the first one is consumed.  ld(r, gc) -> V, st(r, gc, V).
template <int IN_H, int GPR, typename V, typename LD, typename ST>
__device__ __forceinline__ void bv_stage_tile(int tid, LD ld, ST st)
{
    constexpr int RPP = 256 / GPR, NP = (IN_H + RPP - 1) / RPP;
    const int rr = tid / GPR, gc = tid - rr * GPR;
    if (rr >= RPP) return;
    V v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int r = rr + i * RPP;
        if (r < IN_H) v[i] = ld(r, gc);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int r = rr + i * RPP;
        if (r < IN_H) st(r, gc, v[i]);
    }
}

// ---- launch-parameter blocks (passed by value) --------------------------------------------------
struct BvBatch {
    int n;                                  // frames in this launch group
    unsigned flags[BV_MAX_BATCH];           // BV_FRAME_*
};

struct BvPlane {                            // one image plane per frame of the group
    const uint8_t *p[BV_MAX_BATCH];         // frame f of the group (device pointer; zero-copy for bv_submit_device)
    size_t pitch;                           // bytes per row
};
static inline BvPlane bv_plane_contig(const void *base, size_t pitch, size_t frame_stride_bytes, int n)
{
    BvPlane pl;
    for (int f = 0; f < BV_MAX_BATCH; ++f)
        pl.p[f] = f < n ? static_cast<const uint8_t *>(base) + (size_t)f * frame_stride_bytes : nullptr;
    pl.pitch = pitch;
    return pl;
}

struct BvAdmScaleParams {
    int in_w, in_h;                         // DWT input dims
    int w, h;                               // band dims
    int left, top, right, bottom;           // contrast-masking / denominator region
    unsigned rf[3];                         // fixed-point csf factors (h, v, d)
    int sh_sub[3], sh_sq[3], sh_cub[3];     // numerator shifts per band
    unsigned long long add_sq[3], add_cub[3];
    int sh_inner; unsigned long long add_inner;   // per-row numerator shift
    // denominator
    int den_sh_sq; unsigned long long den_add_sq; // scales 1..3
    int den_sh_cub; unsigned long long den_add_cub;
    int den_sh_row; unsigned long long den_add_row;
};

// ---- per-kernel profiling (CUDA events on the launching stream) ----------------------------
enum BvKernelId {
    BVK_MOTION_BLUR = 0, BVK_MOTION_SAD,
    BVK_VIF_STAT0, BVK_VIF_SUB1, BVK_VIF_STAT1, BVK_VIF_SUB2, BVK_VIF_STAT2, BVK_VIF_SUB3, BVK_VIF_STAT3,
    BVK_ADM_S0, BVK_ADM_S1, BVK_ADM_S2, BVK_ADM_S3, BVK_ADM_FINISH,
    BVK_SSE_Y, BVK_SSE_U, BVK_SSE_V,
    BVK_FFSSIM_Y, BVK_FFSSIM_U, BVK_FFSSIM_V,
    BVK_F_FIRST,                       // float kernels: ids BVK_F_FIRST .. BV_MAX_KERNELS-1 (bv_float.cuh)
    BV_MAX_KERNELS = 64
};
struct BvProf {
    bool on = false;
    bool have_events = false;
    cudaEvent_t ev[BV_MAX_KERNELS][2];
    bool used[BV_MAX_KERNELS];
};
struct BvLaunch {
    cudaStream_t st;
    BvProf *prof;
    long long *nlaunch;
};
static inline void bv_prof_begin(const BvLaunch &l, int id)
{
    if (l.prof && l.prof->on) { cudaEventRecord(l.prof->ev[id][0], l.st); l.prof->used[id] = true; }
}
static inline void bv_prof_end(const BvLaunch &l, int id)
{
    if (l.prof && l.prof->on) cudaEventRecord(l.prof->ev[id][1], l.st);
    ++*l.nlaunch;
}

// ---- kernel launchers (one per .cu file) ---------------------------------------------------

// motion
void bv_launch_motion_blur(const BvBatch &b, BvPlane ref_y, int bpc, int w, int h, uint16_t *blur_cur,
                           size_t blur_frame_elems, const BvLaunch &L);
void bv_launch_motion_sad(const BvBatch &b, const uint16_t *blur_cur, const uint16_t *blur_prev_group_last,
                          size_t blur_frame_elems, int w, int h, unsigned long long *raw, const BvLaunch &L);
// vif
struct BvVifLevels {                        // u16 pyramid levels 1..3 (tight pitch), ref and dis
    uint16_t *ref[4], *dis[4];              // [0] unused
    size_t frame_elems[4];
    int w[4], h[4];
};
#define BV_LOG2C_BYTES (1024 + 16384)       // 512 u16 bases + 32768 4-bit deltas
// motion_blur_out != nullptr: the scale-0 kernel also writes the integer motion feature's blurred reference (for every
// frame of the group, scored or not), so the caller skips bv_launch_motion_blur.
void bv_launch_vif(const BvBatch &b, BvPlane ref_y, BvPlane dis_y, int bpc, const BvVifLevels &lv,
                   const uint16_t *log2_table, const uint8_t *log2_packed, double egl, unsigned long long *raw,
                   const BvLaunch &L, uint16_t *motion_blur_out = nullptr, size_t motion_blur_frame_elems = 0);
// adm
struct BvAdmBuffers {
    void *bands[4];                         // scale s: [frame][ref/dis][a,v,h,d][h][w], i16 (s=0) / i32
    size_t band_plane_elems[4];             // w*h of scale s
    unsigned long long *rows;               // [frame][scale][row][6] row accumulators
    size_t rows_frame_stride;               // elements between frames
    size_t rows_scale_offset[4];
    const int *div_lookup;                  // 65537 entries
};
void bv_launch_adm(const BvBatch &b, BvPlane ref_y, BvPlane dis_y, int bpc, const BvAdmBuffers &ab,
                   const BvAdmScaleParams sp[4], double egl, unsigned long long *raw, const BvLaunch &L);
// psnr
void bv_launch_sse(const BvBatch &b, BvPlane ref, BvPlane dis, int bpc, int w, int h, int plane_idx,
                   unsigned long long *raw, const BvLaunch &L);
// ffmpeg ssim filter (sum over 8x8 windows of one plane -> fraw[fraw_idx], atomically)
void bv_launch_ffssim(const BvBatch &b, BvPlane ref, BvPlane dis, int bpc, int w, int h, int plane_idx,
                      double *fraw, int fraw_idx, int fraw_words, const BvLaunch &L);
// adm host-side helpers (bv_adm.cu)
void bv_adm_rfactor(int scale, double view_dist, int display_h, float rf[3]);
void bv_adm_make_params(int w, int h, double view_dist, int display_h, BvAdmScaleParams sp[4]);
void bv_adm_finish_scale(const BvAdmScaleParams &p, int scale, double view_dist, int display_h,
                         const int64_t cm[3], const uint64_t dn[3], float *num_scale, float *den_scale);
// svr
void bv_launch_svr(const double *d_feat, int n_feat, const double *d_slopes, const double *d_intercepts,
                   const double *d_sv, const double *d_coef, int n_sv, double gamma, double rho, double *d_out,
                   long long n, cudaStream_t st);
