// DRAFT -- row f1, NOT part of libscouter_b200.so and never run on a GPU yet.
//
// Weight gradient of nn.Conv2d (any k, stride, padding, groups; timm/models/resnet.py:401-408, resnest.py:84-105,
// split_attn.py:43-52), NHWC activations and this library's OHWI weights:
//     dW[o, r, s, c] = sum_{b, y, x} dY[b, y, x, o] * X[b, y*stride + r - pad, x*stride + s - pad, g(o)*Cin_g + c]
// i.e. a GEMM with K = B*Ho*Wo.  This is the correctness-first CUDA-core version: thread = one weight element (adjacent
// threads = adjacent input channels: the X loads coalesce, the dY load is a broadcast), the K range is split across
// gridDim.y CTAs and merged with fp32 atomics.  The tcgen05 version (MN-major operands, K = rows) replaces it later;
// this one is what a first end-to-end training step can run on.  (The data gradient of the stride-1 convs needs no
// kernel at all: plan.dgrad_weights turns it into a forward conv.)  Bias gradients (fc1 / fc2 / conv1x1 have biases)
// are column sums of dY: conv_bgrad.
// The bodies compile as host code for the emulation in tests/test_conv_wgrad_draft.py.
#pragma once
#include <stddef.h>
#include <math.h>

#ifdef __CUDACC__
#define WG_HD __device__ __forceinline__
#define WG_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define WG_HD static inline
#define WG_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

namespace scouter_draft {

struct WgradArgs {
    int B, H, W, Cin, Ho, Wo, Cout, k, stride, pad, groups;
    const float* x;     // (B, H, W, Cin)
    const float* dy;    // (B, Ho, Wo, Cout)
    float* dw;          // (Cout, k, k, Cin/groups), zeroed by the caller, accumulated here
    float* db;          // (Cout) or NULL, zeroed by the caller
};

// weight element `e` of Cout*k*k*Cin_g, rows [m0, m1) of B*Ho*Wo
WG_HD void conv_wgrad_element(const WgradArgs& a, long long e, long long m0, long long m1) {
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    const int c = (int)(e % cin_g);
    long long t = e / cin_g;
    const int s = (int)(t % a.k); t /= a.k;
    const int r = (int)(t % a.k);
    const int o = (int)(t / a.k);
    const int ci = (o / cout_g) * cin_g + c;
    double acc = 0.0;      // K = B*Ho*Wo reaches 10^5..10^6 terms: a sequential fp32 sum would carry ~1e-3 of rounding noise
    for (long long m = m0; m < m1; ++m) {
        const int xo = (int)(m % a.Wo);
        const long long q = m / a.Wo;
        const int yo = (int)(q % a.Ho);
        const int b = (int)(q / a.Ho);
        const int yi = yo * a.stride + r - a.pad, xi = xo * a.stride + s - a.pad;
        if (yi < 0 || yi >= a.H || xi < 0 || xi >= a.W) continue;
        acc += (double)a.dy[(size_t)m * a.Cout + o] * (double)a.x[(((size_t)b * a.H + yi) * a.W + xi) * a.Cin + ci];
    }
    WG_ATOMIC_ADD(a.dw + e, (float)acc);
}

WG_HD void conv_bgrad_element(const WgradArgs& a, int o, long long m0, long long m1) {
    double acc = 0.0;
    for (long long m = m0; m < m1; ++m) acc += (double)a.dy[(size_t)m * a.Cout + o];
    WG_ATOMIC_ADD(a.db + o, (float)acc);
}

// Data gradient for ANY stride as a gather (thread = one input element): the strided convs of resnet18 (3x3 s2 and the
// 1x1 s2 shortcuts, resnet.py:172-199, 276-289) cannot use the forward-conv trick of plan.dgrad_weights.
//     dX[b, yi, xi, ci] = sum_{o in group(ci), r, s : (yi + pad - r) % stride == 0, ...} dY[b, yo, xo, o] * W[o, r, s, c]
struct DgradArgs {
    int B, H, W, Cin, Ho, Wo, Cout, k, stride, pad, groups;
    const float* dy;    // (B, Ho, Wo, Cout)
    const float* w;     // (Cout, k, k, Cin/groups)
    float* dx;          // (B, H, W, Cin)
};

WG_HD void conv_dgrad_element(const DgradArgs& a, long long idx) {     // idx over B*H*W*Cin
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    const int ci = (int)(idx % a.Cin);
    long long t = idx / a.Cin;
    const int xi = (int)(t % a.W); t /= a.W;
    const int yi = (int)(t % a.H);
    const int b = (int)(t / a.H);
    const int g = ci / cin_g, c = ci % cin_g;
    float acc = 0.f;
    for (int r = 0; r < a.k; ++r) {
        const int ty = yi + a.pad - r;
        if (ty < 0 || ty % a.stride) continue;
        const int yo = ty / a.stride;
        if (yo >= a.Ho) continue;
        for (int s = 0; s < a.k; ++s) {
            const int tx = xi + a.pad - s;
            if (tx < 0 || tx % a.stride) continue;
            const int xo = tx / a.stride;
            if (xo >= a.Wo) continue;
            const float* dyp = a.dy + (((size_t)b * a.Ho + yo) * a.Wo + xo) * a.Cout + (size_t)g * cout_g;
            for (int o = 0; o < cout_g; ++o)
                acc = fmaf(dyp[o], a.w[(((size_t)(g * cout_g + o) * a.k + r) * a.k + s) * cin_g + c], acc);
        }
    }
    a.dx[idx] = acc;
}

#ifdef __CUDACC__
static __global__ void __launch_bounds__(256) conv_dgrad_kernel(DgradArgs a) {
    const long long n = (long long)a.B * a.H * a.W * a.Cin;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) conv_dgrad_element(a, i);
}

// grid (ceil(elements / 256), k_splits)
static __global__ void __launch_bounds__(256) conv_wgrad_kernel(WgradArgs a) {
    const long long elems = (long long)a.Cout * a.k * a.k * (a.Cin / a.groups);
    const long long M = (long long)a.B * a.Ho * a.Wo;
    const long long per = (M + gridDim.y - 1) / gridDim.y;
    const long long m0 = (long long)blockIdx.y * per, m1 = m0 + per < M ? m0 + per : M;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < elems) conv_wgrad_element(a, e, m0, m1);
    if (a.db && e < a.Cout) conv_bgrad_element(a, (int)e, m0, m1);
}
#endif

}  // namespace scouter_draft
