// DRAFT -- row f1, NOT part of libscouter_b200.so and never run on a GPU yet.
//
// Train-mode BatchNorm2d of the backbone (nn.BatchNorm2d defaults: eps 1e-5, momentum 0.1; timm/models/resnet.py:401-420,
// resnest.py:84-105, split_attn.py:46-50) on NHWC activations viewed as (M = B*H*W rows, C channels): in train mode the
// statistics come from the batch, so BatchNorm cannot be folded into the conv weights as the eval path does.
//   1. bn_stats_partial   per-channel sum and sum of squares; thread = (row lane, channel quad), float4 loads that are
//                         coalesced across the quads of a row, fp64 accumulators (E[x^2] - mean^2 cancels badly in fp32
//                         when |mean| >> std), one fp64 atomicAdd pair per thread and channel into a (C, 2) workspace;
//   2. bn_finalize        per channel: batch mean / biased variance -> scale = gamma * rstd, shift = beta - mean * scale,
//                         running_mean / running_var (unbiased) momentum update;
//   3. bn_apply           y = x * scale + shift [+ residual] [ReLU], float4 per thread.
// All three are HBM-bound: one read of x for the statistics, one read + one write for the apply (the apply can later
// move into the consumer conv's TS-mode splitter, which already touches every element).
// The bodies compile as host code for the emulation in tests/test_bn_train_draft.py.
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define BN_HD __device__ __forceinline__
#define BN_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define BN_HD static inline
#define BN_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

namespace scouter_draft {

struct BnTrainArgs {
    long long M;                 // rows = B*H*W
    int C;                       // channels, multiple of 4
    const float* x;              // (M, C)
    double* sums;                // (C, 2) workspace, zeroed by the caller: [sum, sum of squares]
    const float *gamma, *beta;   // (C)
    float *running_mean, *running_var;   // (C), updated in place
    float *scale, *shift;        // (C) outputs of bn_finalize, inputs of bn_apply
    float *save_mean, *save_rstd;// (C) batch mean and 1/sqrt(var + eps), kept for the backward (or NULL)
    float eps, momentum;
    const float* residual;       // (M, C) or NULL
    float* y;                    // (M, C); may alias x
    int relu;
};

// On the GPU the lanes of a CTA that hold the same channel quad are merged in shared memory first (fixed order), so a CTA
// issues ONE fp64 atomic pair per channel instead of one per thread (592 CTAs x 256 threads x 8 atomics on 2*C addresses
// was 140 us per launch).  `red` = (nthreads, 8) doubles of shared memory, NULL in the host emulation (no merge needed).
// Returns whether this thread adds `v` to the global sums.  All threads of the CTA must call it (it synchronises).
BN_HD bool bn_cta_merge(double (*red)[8], int tid, int qpr, int lanes, double v[8]) {
#ifdef __CUDACC__
    if (red && lanes > 1) {
        for (int k = 0; k < 8; ++k) red[tid][k] = v[k];
        __syncthreads();
        if (tid < qpr)
            for (int l = 1; l < lanes; ++l)
                for (int k = 0; k < 8; ++k) v[k] += red[tid + l * qpr][k];
        __syncthreads();
        return tid < qpr;
    }
#endif
    (void)red; (void)tid; (void)qpr; (void)lanes; (void)v;
    return true;
}

// CTA `cta` of `n_ctas` owns a contiguous slab of rows; inside it thread `tid` owns channel quad tid % qpr and the rows
// r0 + tid / qpr, + lanes, + 2*lanes ... (qpr = quads per row handled at once = min(C/4, nthreads)).
BN_HD void bn_stats_partial(const BnTrainArgs& a, int cta, int n_ctas, int tid, int nthreads, double (*red)[8] = nullptr) {
    const int quads = a.C / 4;
    const int qpr = quads < nthreads ? quads : nthreads;
    const int lanes = nthreads / qpr;
    const long long per = (a.M + n_ctas - 1) / n_ctas;
    const long long r0 = (long long)cta * per, r1 = r0 + per < a.M ? r0 + per : a.M;
    const int lane = tid / qpr;
    for (int q0 = 0; q0 < quads; q0 += qpr) {            // more than nthreads quads per row: loop (uniform trip count)
        const int q = q0 + tid % qpr;
        const bool active = tid < lanes * qpr && q < quads;
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};          // [sum, sum of squares] x 4 channels
        if (active)
            for (long long r = r0 + lane; r < r1; r += lanes) {
                const float* p = a.x + r * a.C + 4 * q;
                for (int k = 0; k < 4; ++k) { const double x = p[k]; v[2 * k] += x; v[2 * k + 1] += x * x; }
            }
        if (bn_cta_merge(red, tid, qpr, lanes, v) && active)
            for (int k = 0; k < 8; ++k) BN_ATOMIC_ADD(a.sums + 2 * (4 * q) + k, v[k]);
    }
}

BN_HD void bn_finalize(const BnTrainArgs& a, int c) {
    const double n = (double)a.M;
    const double mean = a.sums[2 * c] / n;
    double var = a.sums[2 * c + 1] / n - mean * mean;    // biased: what the normalisation uses
    if (var < 0) var = 0;
    const double rstd = 1.0 / sqrt(var + (double)a.eps);
    const double sc = (double)a.gamma[c] * rstd;
    a.scale[c] = (float)sc;
    a.shift[c] = (float)((double)a.beta[c] - mean * sc);
    if (a.save_mean) { a.save_mean[c] = (float)mean; a.save_rstd[c] = (float)rstd; }
    const double unbiased = a.M > 1 ? var * n / (n - 1.0) : var;
    a.running_mean[c] = (float)((1.0 - a.momentum) * a.running_mean[c] + a.momentum * mean);
    a.running_var[c] = (float)((1.0 - a.momentum) * a.running_var[c] + a.momentum * unbiased);
}

BN_HD void bn_apply(const BnTrainArgs& a, long long i4) {    // i4 indexes float4s of the (M, C) map
    const int c = (int)((i4 * 4) % a.C);
    for (int k = 0; k < 4; ++k) {
        const long long i = i4 * 4 + k;
        float v = fmaf(a.x[i], a.scale[c + k], a.shift[c + k]);
        if (a.residual) v += a.residual[i];
        if (a.relu && v < 0.f) v = 0.f;
        a.y[i] = v;
    }
}

// ---- backward of  out = [relu]( bn(x) [+ residual] )  ---------------------------------------------------------------
//   dy = d_out * [out > 0]  (ReLU mask from the saved output);  d_residual = dy (the caller aliases / accumulates it);
//   d_beta = sum dy;  d_gamma = sum dy * xhat,  xhat = (x - mean) * rstd;
//   dx = gamma * rstd * (dy - d_beta / M - xhat * d_gamma / M).
// Same decomposition as the forward: one statistics pass (fp64 sums, atomics), a per-channel finalize, one apply pass.
struct BnBwdArgs {
    long long M;
    int C;
    const float *x, *out, *d_out;        // (M, C): conv output, block output (for the ReLU mask; may be NULL when !relu), its gradient
    const float *gamma, *save_mean, *save_rstd;
    double* sums;                        // (C, 2) zeroed workspace: [sum dy, sum dy * xhat]
    float *d_gamma, *d_beta;             // (C) accumulated into (+=): parameters can be shared across calls of a step
    float *coef;                         // (C, 3) from finalize: gamma*rstd, d_beta/M, d_gamma/M
    float *dx;                           // (M, C); may alias d_out
    float *d_residual;                   // (M, C) or NULL: receives dy
    int relu;
};

BN_HD float bn_bwd_dy(const BnBwdArgs& a, long long i) {
    const float g = a.d_out[i];
    return (a.relu && !(a.out[i] > 0.f)) ? 0.f : g;
}

BN_HD void bn_bwd_stats_partial(const BnBwdArgs& a, int cta, int n_ctas, int tid, int nthreads, double (*red)[8] = nullptr) {
    const int quads = a.C / 4;
    const int qpr = quads < nthreads ? quads : nthreads;
    const int lanes = nthreads / qpr;
    const long long per = (a.M + n_ctas - 1) / n_ctas;
    const long long r0 = (long long)cta * per, r1 = r0 + per < a.M ? r0 + per : a.M;
    const int lane = tid / qpr;
    for (int q0 = 0; q0 < quads; q0 += qpr) {
        const int q = q0 + tid % qpr;
        const bool active = tid < lanes * qpr && q < quads;
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};          // [sum dy, sum dy * xhat] x 4 channels
        if (active)
            for (long long r = r0 + lane; r < r1; r += lanes) {
                for (int k = 0; k < 4; ++k) {
                    const int c = 4 * q + k;
                    const long long i = r * a.C + c;
                    const double dy = bn_bwd_dy(a, i);
                    v[2 * k] += dy;
                    v[2 * k + 1] += dy * (((double)a.x[i] - a.save_mean[c]) * a.save_rstd[c]);
                }
            }
        if (bn_cta_merge(red, tid, qpr, lanes, v) && active)
            for (int k = 0; k < 8; ++k) BN_ATOMIC_ADD(a.sums + 2 * (4 * q) + k, v[k]);
    }
}

BN_HD void bn_bwd_finalize(const BnBwdArgs& a, int c) {
    const double n = (double)a.M;
    a.d_beta[c] += (float)a.sums[2 * c];
    a.d_gamma[c] += (float)a.sums[2 * c + 1];
    a.coef[3 * c] = a.gamma[c] * a.save_rstd[c];
    a.coef[3 * c + 1] = (float)(a.sums[2 * c] / n);
    a.coef[3 * c + 2] = (float)(a.sums[2 * c + 1] / n);
}

BN_HD void bn_bwd_apply(const BnBwdArgs& a, long long i4) {
    const int c0 = (int)((i4 * 4) % a.C);
    for (int k = 0; k < 4; ++k) {
        const long long i = i4 * 4 + k;
        const int c = c0 + k;
        const float dy = bn_bwd_dy(a, i);
        const float xhat = (a.x[i] - a.save_mean[c]) * a.save_rstd[c];
        if (a.d_residual) a.d_residual[i] = dy;
        a.dx[i] = a.coef[3 * c] * (dy - a.coef[3 * c + 1] - xhat * a.coef[3 * c + 2]);
    }
}

#ifdef __CUDACC__
static __global__ void __launch_bounds__(256) bn_bwd_stats_kernel(BnBwdArgs a) {
    __shared__ double red[256][8];
    bn_bwd_stats_partial(a, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x, red);
}
static __global__ void bn_bwd_finalize_kernel(BnBwdArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < a.C) bn_bwd_finalize(a, c);
}
static __global__ void __launch_bounds__(256) bn_bwd_apply_kernel(BnBwdArgs a) {
    const long long n4 = a.M * a.C / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) bn_bwd_apply(a, i);
}
static __global__ void __launch_bounds__(256) bn_stats_kernel(BnTrainArgs a) {
    __shared__ double red[256][8];
    bn_stats_partial(a, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x, red);
}
static __global__ void bn_finalize_kernel(BnTrainArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < a.C) bn_finalize(a, c);
}
static __global__ void __launch_bounds__(256) bn_apply_kernel(BnTrainArgs a) {
    const long long n4 = a.M * a.C / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) bn_apply(a, i);
}
#endif

}  // namespace scouter_draft
