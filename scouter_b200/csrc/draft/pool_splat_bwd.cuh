// DRAFT -- row f1, NOT part of libscouter_b200.so and never run on a GPU yet.
//
// Backward of the memory-bound helpers of the backbone (forward: csrc/aux_kernels.cu), NHWC fp32, written as gathers
// (thread = one input element, no atomics, deterministic):
//   maxpool_bwd      nn.MaxPool2d(3, 2, 1), resnet.py:420 -- the window's arg-max is recomputed; ties go to the first
//                    maximum in (kh, kw) scan order like PyTorch (`val > max`), which matters after a ReLU (zeros tie)
//   avgpool2_bwd     AvgPool2d(2, 2, ceil_mode=True, count_include_pad=False), resnet.py:300 (shortcut avg-down)
//   avgpool3_bwd     AvgPool2d(3, 2, padding=1) (count_include_pad=True: divisor 9), resnest.py:101 (avd_last)
//   splat_bwd_*      split attention, split_attn.py:62-79 with radix 2:
//                      out = sum_r x_r * att_r,  att = softmax_r(fc2(relu(bn1(fc1(gap))))),  gap = mean_hw(x_0 + x_1)
//                    reduce : d_att[b,r,c] = sum_hw d_out[b,hw,c] * x_r[b,hw,c]
//                    softmax: d_logit[b,r,c] = att_r * (d_att_r - sum_r' att_r' * d_att_r')
//                    apply  : d_x_r[b,hw,c] = d_out[b,hw,c] * att[b,r,c] + d_gap[b,c] / (H*W)
//                    (fc2 / bn1 / fc1 in between are ordinary 1x1 convs and a BatchNorm on (B,1,1,.) maps)
// The bodies compile as host code for the emulation in tests/test_pool_splat_bwd_draft.py.
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define PS_HD __device__ __forceinline__
#else
#define PS_HD static inline
#endif

namespace scouter_draft {

struct PoolBwdArgs {
    int B, H, W, C, Ho, Wo;
    const float* x;        // (B, H, W, C) forward input (max-pool only)
    const float* dy;       // (B, Ho, Wo, C)
    float* dx;             // (B, H, W, C)
};

PS_HD void maxpool_bwd(const PoolBwdArgs& a, long long idx) {          // idx over B*H*W*C
    const int c = (int)(idx % a.C);
    long long t = idx / a.C;
    const int w = (int)(t % a.W); t /= a.W;
    const int h = (int)(t % a.H);
    const int b = (int)(t / a.H);
    const float* xb = a.x + (size_t)b * a.H * a.W * a.C;
    float g = 0.f;
    // output windows (k=3, s=2, p=1) that contain (h, w): ho in {ceil((h-1)/2) .. floor((h+1)/2)}
    for (int ho = (h >> 1); ho <= ((h + 1) >> 1); ++ho) {
        if (ho >= a.Ho) continue;
        for (int wo = (w >> 1); wo <= ((w + 1) >> 1); ++wo) {
            if (wo >= a.Wo) continue;
            int bh = -1, bw = -1;
            float best = -INFINITY;
            for (int r = 0; r < 3; ++r) {
                const int hh = 2 * ho - 1 + r;
                if (hh < 0 || hh >= a.H) continue;
                for (int s = 0; s < 3; ++s) {
                    const int ww = 2 * wo - 1 + s;
                    if (ww < 0 || ww >= a.W) continue;
                    const float v = xb[((size_t)hh * a.W + ww) * a.C + c];
                    if (v > best || bh < 0) { best = v; bh = hh; bw = ww; }
                }
            }
            if (bh == h && bw == w) g += a.dy[(((size_t)b * a.Ho + ho) * a.Wo + wo) * a.C + c];
        }
    }
    a.dx[idx] = g;
}

PS_HD void avgpool2_bwd(const PoolBwdArgs& a, long long idx) {         // k=2, s=2, ceil_mode, count_include_pad=False
    const int c = (int)(idx % a.C);
    long long t = idx / a.C;
    const int w = (int)(t % a.W); t /= a.W;
    const int h = (int)(t % a.H);
    const int b = (int)(t / a.H);
    const int ho = h >> 1, wo = w >> 1;
    const int nh = (2 * ho + 2 <= a.H) ? 2 : 1, nw = (2 * wo + 2 <= a.W) ? 2 : 1;     // valid elements of the window
    a.dx[idx] = a.dy[(((size_t)b * a.Ho + ho) * a.Wo + wo) * a.C + c] / (float)(nh * nw);
}

PS_HD void avgpool3_bwd(const PoolBwdArgs& a, long long idx) {         // k=3, s=2, p=1, divisor 9
    const int c = (int)(idx % a.C);
    long long t = idx / a.C;
    const int w = (int)(t % a.W); t /= a.W;
    const int h = (int)(t % a.H);
    const int b = (int)(t / a.H);
    float g = 0.f;
    for (int ho = (h >> 1); ho <= ((h + 1) >> 1); ++ho) {
        if (ho >= a.Ho) continue;
        for (int wo = (w >> 1); wo <= ((w + 1) >> 1); ++wo) {
            if (wo >= a.Wo) continue;
            g += a.dy[(((size_t)b * a.Ho + ho) * a.Wo + wo) * a.C + c];
        }
    }
    a.dx[idx] = g * (1.0f / 9.0f);
}

struct SplatBwdArgs {
    int B, HW, C;              // x2 is (B, HW, 2C) radix-major channels; out / d_out are (B, HW, C)
    const float* x2;
    const float* d_out;
    const float* att;          // (B, 2, C) softmax over the radix
    float* d_att;              // (B, 2, C)
    float* d_logit;            // (B, 2, C): gradient at fc2's output (radix-major, like the forward's logits)
    const float* d_gap;        // (B, C): gradient at the pooled descriptor (from fc1's dgrad)
    float* d_x2;               // (B, HW, 2C)
};

PS_HD void splat_bwd_reduce(const SplatBwdArgs& a, long long idx) {     // idx over B*2*C
    const int c = (int)(idx % a.C);
    const int r = (int)((idx / a.C) % 2);
    const int b = (int)(idx / (2 * a.C));
    float s = 0.f;
    for (int p = 0; p < a.HW; ++p)
        s = fmaf(a.d_out[((size_t)b * a.HW + p) * a.C + c], a.x2[((size_t)b * a.HW + p) * 2 * a.C + r * a.C + c], s);
    a.d_att[idx] = s;
}

PS_HD void splat_bwd_softmax(const SplatBwdArgs& a, long long idx) {    // idx over B*C
    const int c = (int)(idx % a.C);
    const int b = (int)(idx / a.C);
    const size_t i0 = ((size_t)b * 2) * a.C + c, i1 = i0 + a.C;
    const float dot = a.att[i0] * a.d_att[i0] + a.att[i1] * a.d_att[i1];
    a.d_logit[i0] = a.att[i0] * (a.d_att[i0] - dot);
    a.d_logit[i1] = a.att[i1] * (a.d_att[i1] - dot);
}

PS_HD void splat_bwd_apply(const SplatBwdArgs& a, long long idx) {      // idx over B*HW*2C
    const int cc = (int)(idx % (2 * a.C));
    const long long t = idx / (2 * a.C);
    const int b = (int)(t / a.HW);
    const int r = cc / a.C, c = cc % a.C;
    a.d_x2[idx] = a.d_out[(size_t)t * a.C + c] * a.att[((size_t)b * 2 + r) * a.C + c] + a.d_gap[(size_t)b * a.C + c] / (float)a.HW;
}

#ifdef __CUDACC__
#define PS_KERNEL(name, fn, Args, count)                                                                     \
    static __global__ void __launch_bounds__(256) name(Args a) {                                                   \
        const long long n_ = (count);                                                                       \
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_; i += (long long)gridDim.x * blockDim.x) fn(a, i); \
    }
PS_KERNEL(maxpool_bwd_kernel, maxpool_bwd, PoolBwdArgs, (long long)a.B * a.H * a.W * a.C)
PS_KERNEL(avgpool2_bwd_kernel, avgpool2_bwd, PoolBwdArgs, (long long)a.B * a.H * a.W * a.C)
PS_KERNEL(avgpool3_bwd_kernel, avgpool3_bwd, PoolBwdArgs, (long long)a.B * a.H * a.W * a.C)
PS_KERNEL(splat_bwd_reduce_kernel, splat_bwd_reduce, SplatBwdArgs, (long long)a.B * 2 * a.C)
PS_KERNEL(splat_bwd_softmax_kernel, splat_bwd_softmax, SplatBwdArgs, (long long)a.B * a.C)
PS_KERNEL(splat_bwd_apply_kernel, splat_bwd_apply, SplatBwdArgs, (long long)a.B * a.HW * 2 * a.C)
#endif

}  // namespace scouter_draft
