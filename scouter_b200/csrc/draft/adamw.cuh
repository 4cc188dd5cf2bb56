// DRAFT -- row f1, NOT part of libscouter_b200.so and never run on a GPU yet.
//
// The optimizer step of the reference: torch.optim.AdamW(params, lr=args.lr) (train.py:146 -- the parser's
// --weight_decay is never passed, so torch's defaults apply: betas (0.9, 0.999), eps 1e-8, weight_decay 0.01),
// optimizer.step() in engine.py:34.  One elementwise pass over the flat fp32 buffers of dist.GradientBuckets
// (parameters, gradients, exp_avg, exp_avg_sq), in the operation order of torch's single-tensor implementation so the
// result is the same to the last bit or two:
//     p *= 1 - lr*wd;  m += (g - m)*(1 - b1);  v = v*b2 + g*g*(1 - b2);
//     p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Memory-bound: 16 B read + 12 B written per parameter (15.2 M parameters: 0.43 GB per step, ~65 us at HBM speed).
// The body compiles as host code for tests/test_head_backward_draft.py-style emulation (tests/test_adamw_draft.py).
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define AW_HD __device__ __forceinline__
#else
#define AW_HD static inline
#endif

namespace scouter_draft {

struct AdamWArgs {       // every scalar is computed by the host in double (as torch does in Python) and rounded once
    float decay;                 // 1 - lr * weight_decay
    float one_minus_beta1, beta2, one_minus_beta2, eps;
    float step_size;             // lr / (1 - beta1^t)
    float bias_correction2_sqrt; // sqrt(1 - beta2^t)
    size_t n;
    float* p;
    const float* g;
    float *m, *v;
};

AW_HD void adamw_element(const AdamWArgs& a, size_t i) {
    float p = a.p[i] * a.decay;
    const float g = a.g[i];
    const float m = a.m[i] + (g - a.m[i]) * a.one_minus_beta1;          // lerp_
    const float v = a.v[i] * a.beta2 + g * g * a.one_minus_beta2;       // mul_ + addcmul_
    const float denom = sqrtf(v) / a.bias_correction2_sqrt + a.eps;
    p -= a.step_size * (m / denom);                                     // addcdiv_
    a.p[i] = p; a.m[i] = m; a.v[i] = v;
}

#ifdef __CUDACC__
static __global__ void __launch_bounds__(256) adamw_kernel(AdamWArgs a) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x)
        adamw_element(a, i);
}
#endif

}  // namespace scouter_draft
