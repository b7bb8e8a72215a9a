// SVR fusion: libsvm nu-SVR / RBF decision value + libvmaf predict.c normalisation, transform and
// clip (reached from the reference at app/vmaf_analyzer.py:417; model data:
// reference models/*.json).  bv_predict evaluates on the host in libsvm's sequential order;
// bv_predict_device runs the svr_predict kernel (bv_misc.cu) for whole clips at once.
#include "bv_common.cuh"
#include "../../include/b200vmaf.h"
#include <math.h>
#include <mutex>
#include <vector>

struct bv_model {
    int n_feat = 0, n_sv = 0;
    std::vector<double> sv, coef, slopes, intercepts;
    double gamma = 0, rho = 0;
    double clip[2] = { 0, 0 };
    int has_clip = 0;
    double tp[3] = { 0, 0, 0 };
    unsigned tflags = 0;
    // Device mirrors: one per GPU, created on first use under `mu` and never freed before bv_model_free, so
    // a VmafModel shared by one worker thread per GPU (engine.analyze_batch, sweep.run_sweep) is safe: a thread on
    // device A never sees (or frees) the buffers of device B.
    // Each mirror also owns a stream and grow-only scratch for the feature rows and the scores: a predict call must not
    // cudaMalloc / cudaFree (cudaFree waits for EVERY stream of the device, i.e. for the extractor kernels still draining
    // while the host already scores the first part of a clip -- engine.analyze builds those frames during the drain).
    struct Mirror {
        int dev = -1;
        double *d_sv = nullptr, *d_coef = nullptr, *d_slopes = nullptr, *d_intercepts = nullptr;
        double *d_feat = nullptr, *d_out = nullptr;
        int64_t cap = 0;
        cudaStream_t st = nullptr;
    };
    std::mutex mu;                       // held for a whole bv_predict_device call (they take ~0.1 ms)
    std::vector<Mirror> mirrors;
};

namespace {

// libvmaf predict.c: score transform (polynomial, then out_lte_in / out_gte_in), then clip
double post(const bv_model *m, double y, unsigned flags)
{
    if (flags & BV_MODEL_ENABLE_TRANSFORM) {
        const double in = y;
        double v = 0.0;
        if (m->tflags & 1u) v += m->tp[0];
        if (m->tflags & 2u) v += m->tp[1] * in;
        if (m->tflags & 4u) v += m->tp[2] * in * in;
        if (m->tflags & 7u) y = v;
        if ((m->tflags & 8u) && y > in) y = in;
        if ((m->tflags & 16u) && y < in) y = in;
    }
    if (!(flags & BV_MODEL_DISABLE_CLIP) && m->has_clip) {
        if (y < m->clip[0]) y = m->clip[0];
        if (y > m->clip[1]) y = m->clip[1];
    }
    return y;
}

}  // namespace

extern "C" {

bv_model *bv_model_create(int n_feat, int n_sv, const double *sv, const double *coef, double gamma, double rho,
                          const double *slopes, const double *intercepts, const double clip[2], int has_clip,
                          const double transform_p[3], unsigned transform_flags)
{
    if (n_feat <= 0 || n_feat > 64 || n_sv <= 0 || !sv || !coef || !slopes || !intercepts) return nullptr;
    bv_model *m = new bv_model();
    m->n_feat = n_feat; m->n_sv = n_sv;
    m->sv.assign(sv, sv + (size_t)n_sv * n_feat);
    m->coef.assign(coef, coef + n_sv);
    m->slopes.assign(slopes, slopes + n_feat + 1);
    m->intercepts.assign(intercepts, intercepts + n_feat + 1);
    m->gamma = gamma; m->rho = rho;
    m->has_clip = has_clip && clip;
    if (m->has_clip) { m->clip[0] = clip[0]; m->clip[1] = clip[1]; }
    if (transform_p) { m->tp[0] = transform_p[0]; m->tp[1] = transform_p[1]; m->tp[2] = transform_p[2]; }
    m->tflags = transform_flags;
    return m;
}

void bv_model_free(bv_model *m)
{
    if (!m) return;
    for (const bv_model::Mirror &r : m->mirrors) {
        if (cudaSetDevice(r.dev) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaFree(r.d_sv); cudaFree(r.d_coef); cudaFree(r.d_slopes); cudaFree(r.d_intercepts);
        cudaFree(r.d_feat); cudaFree(r.d_out);
        if (r.st) cudaStreamDestroy(r.st);
    }
    delete m;
}

int bv_predict(const bv_model *m, const double *feat, int64_t n, unsigned flags, double *out)
{
    if (!m || !feat || !out || n < 0) return BV_ERR_ARG;
    const int nf = m->n_feat;
    for (int64_t r = 0; r < n; ++r) {
        double x[64];
        for (int i = 0; i < nf; ++i) x[i] = m->slopes[i + 1] * feat[r * nf + i] + m->intercepts[i + 1];
        double sum = 0;
        for (int k = 0; k < m->n_sv; ++k) {
            double d2 = 0;
            for (int i = 0; i < nf; ++i) {
                const double d = x[i] - m->sv[(size_t)k * nf + i];
                d2 += d * d;
            }
            sum += m->coef[k] * exp(-m->gamma * d2);
        }
        sum -= m->rho;
        out[r] = post(m, (sum - m->intercepts[0]) / m->slopes[0], flags);
    }
    return 0;
}

int bv_predict_device(const bv_model *cm, int device, const double *feat, int64_t n, unsigned flags, double *out)
{
    bv_model *m = const_cast<bv_model *>(cm);
    if (!m || !feat || !out || n < 0) return BV_ERR_ARG;
    if (n == 0) return 0;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return BV_ERR_CUDA; }
    std::lock_guard<std::mutex> lock(m->mu);
    bv_model::Mirror *mir = nullptr;
    for (bv_model::Mirror &r : m->mirrors) if (r.dev == device) mir = &r;
    if (!mir) {
        bv_model::Mirror r;
        const size_t nsv = m->sv.size() * sizeof(double), ncoef = m->coef.size() * sizeof(double),
                     nsl = m->slopes.size() * sizeof(double);
        if (cudaMalloc(&r.d_sv, nsv) || cudaMalloc(&r.d_coef, ncoef) || cudaMalloc(&r.d_slopes, nsl) ||
            cudaMalloc(&r.d_intercepts, nsl) ||
            cudaMemcpy(r.d_sv, m->sv.data(), nsv, cudaMemcpyHostToDevice) ||
            cudaMemcpy(r.d_coef, m->coef.data(), ncoef, cudaMemcpyHostToDevice) ||
            cudaMemcpy(r.d_slopes, m->slopes.data(), nsl, cudaMemcpyHostToDevice) ||
            cudaMemcpy(r.d_intercepts, m->intercepts.data(), nsl, cudaMemcpyHostToDevice) ||
            cudaStreamCreateWithFlags(&r.st, cudaStreamNonBlocking)) {
            cudaGetLastError();
            cudaFree(r.d_sv); cudaFree(r.d_coef); cudaFree(r.d_slopes); cudaFree(r.d_intercepts);
            return BV_ERR_CUDA;
        }
        r.dev = device;
        m->mirrors.push_back(r);
        mir = &m->mirrors.back();
    }
    if (n > mir->cap) {                  // rare: the first clip, or a longer one than any before
        const int64_t cap = n > 2 * mir->cap ? n : 2 * mir->cap;
        cudaFree(mir->d_feat); cudaFree(mir->d_out);
        mir->d_feat = mir->d_out = nullptr;
        mir->cap = 0;
        if (cudaMalloc(&mir->d_feat, sizeof(double) * cap * m->n_feat) || cudaMalloc(&mir->d_out, sizeof(double) * cap)) {
            cudaGetLastError(); cudaFree(mir->d_feat); mir->d_feat = nullptr; return BV_ERR_CUDA;
        }
        mir->cap = cap;
    }
    cudaMemcpyAsync(mir->d_feat, feat, sizeof(double) * n * m->n_feat, cudaMemcpyHostToDevice, mir->st);
    bv_launch_svr(mir->d_feat, m->n_feat, mir->d_slopes, mir->d_intercepts, mir->d_sv, mir->d_coef, m->n_sv, m->gamma,
                  m->rho, mir->d_out, n, mir->st);
    cudaMemcpyAsync(out, mir->d_out, sizeof(double) * n, cudaMemcpyDeviceToHost, mir->st);
    if (cudaStreamSynchronize(mir->st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return BV_ERR_CUDA;
    for (int64_t r = 0; r < n; ++r) out[r] = post(m, out[r], flags);
    return 0;
}

}  // extern "C"
