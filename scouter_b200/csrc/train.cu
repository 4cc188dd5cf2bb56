// Row f1: C-ABI entries of the training step -- train-mode BatchNorm, the backward of every op on the hot
// path, AdamW -- plus the forward single-op entries the train-mode forward needs (it keeps every buffer alive for the
// backward, so it does not go through the eval plan's arena).  The kernels are the bodies of csrc/draft/*.cuh (also
// compiled as host code by the CPU emulation tests); the structs of include/scouter_b200.h mirror theirs field by field.
#include <cuda_runtime.h>

#include "common.cuh"
#include "draft/adamw.cuh"
#include "draft/bn_train.cuh"
#include "draft/conv_wgrad.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "draft/head_backward.cuh"
#include "draft/pool_splat_bwd.cuh"

namespace scouter_draft {
int head_backward_launch(const HeadBwdArgs& a, cudaStream_t stream);
size_t head_backward_scratch_floats(int n, int S, int L, int iters);
int bn_train_launch(const BnTrainArgs& a, int sms, cudaStream_t stream);
int bn_train_backward_launch(const BnBwdArgs& a, int sms, cudaStream_t stream);
int maxpool_bwd_launch(const PoolBwdArgs& a, int sms, cudaStream_t s);
int avgpool2_bwd_launch(const PoolBwdArgs& a, int sms, cudaStream_t s);
int avgpool3_bwd_launch(const PoolBwdArgs& a, int sms, cudaStream_t s);
int splat_bwd_reduce_launch(const SplatBwdArgs& a, int sms, cudaStream_t s);
int splat_bwd_apply_launch(const SplatBwdArgs& a, int sms, cudaStream_t s);
int conv_wgrad_launch(const WgradArgs& a, int sms, cudaStream_t stream);
int conv_dgrad_launch(const DgradArgs& a, int sms, cudaStream_t stream);
int adamw_launch(const AdamWArgs& a, int sms, cudaStream_t stream);
}  // namespace scouter_draft

using namespace scouter;
namespace sd = scouter_draft;

static_assert(sizeof(scouter_bn_train_args_t) == sizeof(sd::BnTrainArgs), "bn_train args layout");
static_assert(sizeof(scouter_bn_bwd_args_t) == sizeof(sd::BnBwdArgs), "bn_bwd args layout");
static_assert(sizeof(scouter_wgrad_args_t) == sizeof(sd::WgradArgs), "wgrad args layout");
static_assert(sizeof(scouter_dgrad_args_t) == sizeof(sd::DgradArgs), "dgrad args layout");
static_assert(sizeof(scouter_pool_bwd_args_t) == sizeof(sd::PoolBwdArgs), "pool_bwd args layout");
static_assert(sizeof(scouter_splat_bwd_args_t) == sizeof(sd::SplatBwdArgs), "splat_bwd args layout");
static_assert(sizeof(scouter_head_bwd_args_t) == sizeof(sd::HeadBwdArgs), "head_bwd args layout");
static_assert(sizeof(scouter_adamw_args_t) == sizeof(sd::AdamWArgs), "adamw args layout");
static_assert(SCOUTER_MAX_TO_K_LAYERS == sd::HB_MAX_L, "to_k layer limit");

static int sm_count() {
    static int sms = [] {
        int dev = 0, n = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    return sms;
}
static int done(int rc, const char* what) {
    if (rc) set_error("%s: kernel launch failed: %s", what, cudaGetErrorString((cudaError_t)rc));
    return rc;
}

extern "C" int scouter_train_bn_forward(const scouter_bn_train_args_t* a, scouter_stream_t stream) {
    SC_CHECK_ARG(a && a->x && a->y && a->sums && a->gamma && a->beta && a->running_mean && a->running_var && a->scale && a->shift,
                 SCOUTER_E_INVALID, "train_bn_forward: NULL pointer argument");
    SC_CHECK_ARG(a->M > 0 && a->C > 0 && a->C % 4 == 0, SCOUTER_E_UNSUPPORTED, "train_bn_forward: M=%lld C=%d (C must be a multiple of 4)", a->M, a->C);
    cudaStream_t s = (cudaStream_t)stream;
    SC_CUDA(cudaMemsetAsync(a->sums, 0, (size_t)a->C * 2 * sizeof(double), s));
    return done(sd::bn_train_launch(*reinterpret_cast<const sd::BnTrainArgs*>(a), sm_count(), s), "train_bn_forward");
}

extern "C" int scouter_train_bn_backward(const scouter_bn_bwd_args_t* a, scouter_stream_t stream) {
    SC_CHECK_ARG(a && a->x && a->d_out && a->gamma && a->save_mean && a->save_rstd && a->sums && a->d_gamma && a->d_beta && a->coef && a->dx,
                 SCOUTER_E_INVALID, "train_bn_backward: NULL pointer argument");
    SC_CHECK_ARG(!a->relu || a->out, SCOUTER_E_INVALID, "train_bn_backward: relu needs the forward output for its mask");
    SC_CHECK_ARG(a->M > 0 && a->C > 0 && a->C % 4 == 0, SCOUTER_E_UNSUPPORTED, "train_bn_backward: M=%lld C=%d", a->M, a->C);
    cudaStream_t s = (cudaStream_t)stream;
    SC_CUDA(cudaMemsetAsync(a->sums, 0, (size_t)a->C * 2 * sizeof(double), s));
    return done(sd::bn_train_backward_launch(*reinterpret_cast<const sd::BnBwdArgs*>(a), sm_count(), s), "train_bn_backward");
}

extern "C" int scouter_train_conv_wgrad(const scouter_wgrad_args_t* a, scouter_stream_t stream) {
    SC_CHECK_ARG(a && a->x && a->dy && a->dw, SCOUTER_E_INVALID, "train_conv_wgrad: NULL pointer argument");
    SC_CHECK_ARG(a->groups >= 1 && a->Cin % a->groups == 0 && a->Cout % a->groups == 0 && a->k >= 1 && a->stride >= 1, SCOUTER_E_INVALID,
                 "train_conv_wgrad: geometry");
    return done(sd::conv_wgrad_launch(*reinterpret_cast<const sd::WgradArgs*>(a), sm_count(), (cudaStream_t)stream), "train_conv_wgrad");
}

extern "C" int scouter_train_conv_dgrad(const scouter_dgrad_args_t* a, scouter_stream_t stream) {
    SC_CHECK_ARG(a && a->dy && a->w && a->dx, SCOUTER_E_INVALID, "train_conv_dgrad: NULL pointer argument");
    SC_CHECK_ARG(a->groups >= 1 && a->Cin % a->groups == 0 && a->Cout % a->groups == 0 && a->k >= 1 && a->stride >= 1, SCOUTER_E_INVALID,
                 "train_conv_dgrad: geometry");
    return done(sd::conv_dgrad_launch(*reinterpret_cast<const sd::DgradArgs*>(a), sm_count(), (cudaStream_t)stream), "train_conv_dgrad");
}

extern "C" int scouter_train_pool_backward(const scouter_pool_bwd_args_t* a, int kind, scouter_stream_t stream) {
    SC_CHECK_ARG(a && a->dy && a->dx && (kind != 0 || a->x), SCOUTER_E_INVALID, "train_pool_backward: NULL pointer argument");
    SC_CHECK_ARG(kind >= 0 && kind <= 2, SCOUTER_E_INVALID, "train_pool_backward: kind=%d", kind);
    const sd::PoolBwdArgs& p = *reinterpret_cast<const sd::PoolBwdArgs*>(a);
    cudaStream_t s = (cudaStream_t)stream;
    return done(kind == 0 ? sd::maxpool_bwd_launch(p, sm_count(), s) : kind == 1 ? sd::avgpool2_bwd_launch(p, sm_count(), s)
                                                                                  : sd::avgpool3_bwd_launch(p, sm_count(), s), "train_pool_backward");
}

extern "C" int scouter_pool_forward(int kind, const float* in, float* out, int batch, int h, int w, int c, scouter_stream_t stream) {
    SC_CHECK_ARG(in && out && batch > 0 && h > 0 && w > 0 && c > 0, SCOUTER_E_INVALID, "pool_forward: bad arguments");
    SC_CHECK_ARG(kind >= 0 && kind <= 2, SCOUTER_E_INVALID, "pool_forward: kind=%d", kind);
    cudaStream_t s = (cudaStream_t)stream;
    if (kind == 1) {                       // AvgPool2d(2, 2, ceil_mode=True, count_include_pad=False)
        const int ho = (h + 1) / 2, wo = (w + 1) / 2;
        return launch_avgpool(in, out, batch, h, w, c, ho, wo, 2, 2, 0, 0, 0, s);
    }
    const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
    if (kind == 0) return launch_maxpool(in, out, batch, h, w, c, ho, wo, 3, 2, 1, s);
    return launch_avgpool(in, out, batch, h, w, c, ho, wo, 3, 2, 1, 1, 0, s);   // count_include_pad=True (resnest.py:101)
}

extern "C" int scouter_train_splat_backward(const scouter_splat_bwd_args_t* a, int stage, scouter_stream_t stream) {
    SC_CHECK_ARG(a && a->x2 && a->d_out && a->att, SCOUTER_E_INVALID, "train_splat_backward: NULL pointer argument");
    SC_CHECK_ARG(stage == 0 ? (a->d_att && a->d_logit) : (stage == 1 && a->d_gap && a->d_x2), SCOUTER_E_INVALID,
                 "train_splat_backward: stage %d outputs missing", stage);
    const sd::SplatBwdArgs& p = *reinterpret_cast<const sd::SplatBwdArgs*>(a);
    cudaStream_t s = (cudaStream_t)stream;
    return done(stage == 0 ? sd::splat_bwd_reduce_launch(p, sm_count(), s) : sd::splat_bwd_apply_launch(p, sm_count(), s), "train_splat_backward");
}

extern "C" size_t scouter_splat_gap_scratch_floats(int batch, int hw, int c) {
    if (batch <= 0 || hw <= 0 || c <= 0) return 0;
    return (size_t)batch * splat_gap_splits(batch, hw) * 2 * c;
}
extern "C" int scouter_splat_gap_forward(const float* in, float* scratch, float* gap, int batch, int hw, int c, scouter_stream_t stream) {
    SC_CHECK_ARG(in && scratch && gap && batch > 0 && hw > 0 && c > 0, SCOUTER_E_INVALID, "splat_gap_forward: bad arguments");
    return launch_splat_gap(in, scratch, gap, batch, hw, c, (cudaStream_t)stream);
}
extern "C" int scouter_splat_apply_forward(const float* in, const float* logits, float* out, int batch, int h, int w, int c,
                                           scouter_stream_t stream) {
    SC_CHECK_ARG(in && logits && out && batch > 0 && h > 0 && w > 0 && c > 0, SCOUTER_E_INVALID, "splat_apply_forward: bad arguments");
    return launch_splat_apply(in, logits, out, batch, h, w, c, h, w, 0, 0, (cudaStream_t)stream);
}

extern "C" size_t scouter_train_head_backward_scratch_floats(int n, int s, int to_k_layers, int iters) {
    if (n <= 0 || s <= 0 || to_k_layers < 1 || to_k_layers > SCOUTER_MAX_TO_K_LAYERS || iters < 1) return 0;
    return sd::head_backward_scratch_floats(n, s, to_k_layers, iters);
}
extern "C" int scouter_train_head_backward(const scouter_head_bwd_args_t* a, scouter_stream_t stream) {
    SC_CHECK_ARG(a && a->feat && a->conv_w && a->conv_b && a->pe && a->w_ih && a->w_hh && a->b_ih && a->b_hh && a->slots0 && a->g_logits &&
                     a->attn_coef && a->scratch, SCOUTER_E_INVALID, "train_head_backward: NULL pointer argument");
    SC_CHECK_ARG(a->B > 0 && a->n > 0 && a->ch > 0 && a->S == a->C * a->spc && a->L >= 1 && a->L <= SCOUTER_MAX_TO_K_LAYERS && a->iters >= 1,
                 SCOUTER_E_INVALID, "train_head_backward: B=%d n=%d ch=%d S=%d C=%d spc=%d L=%d iters=%d", a->B, a->n, a->ch, a->S, a->C, a->spc, a->L, a->iters);
    SC_CHECK_ARG(a->scratch_per_image >= sd::head_backward_scratch_floats(a->n, a->S, a->L, a->iters), SCOUTER_E_INVALID,
                 "train_head_backward: scratch_per_image=%zu too small", a->scratch_per_image);
    SC_CHECK_ARG(a->g_conv_w && a->g_conv_b && a->g_w_ih && a->g_w_hh && a->g_b_ih && a->g_b_hh && a->g_slots0 && (a->d_feat || a->d_pre),
                 SCOUTER_E_INVALID, "train_head_backward: NULL gradient output");
    return done(sd::head_backward_launch(*reinterpret_cast<const sd::HeadBwdArgs*>(a), (cudaStream_t)stream), "train_head_backward");
}

namespace {
__global__ void __launch_bounds__(256) split_weights_f16_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = w[i];
        const __half h = __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
        const __nv_bfloat16 r = __float2bfloat16_rn(x - __half2float(h));
        out[i] = __half_as_ushort(h);
        out[n + i] = __bfloat16_as_ushort(r);
    }
}
}  // namespace

extern "C" int scouter_split_weights_f16(const float* w, void* out16, size_t n, scouter_stream_t stream) {
    SC_CHECK_ARG(w && out16 && n > 0, SCOUTER_E_INVALID, "split_weights_f16: NULL pointer / empty tensor");
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    split_weights_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, (uint16_t*)out16, n);
    return done((int)cudaGetLastError(), "split_weights_f16");
}

extern "C" int scouter_train_adamw_step(const scouter_adamw_args_t* a, scouter_stream_t stream) {
    SC_CHECK_ARG(a && a->p && a->g && a->m && a->v, SCOUTER_E_INVALID, "train_adamw_step: NULL pointer argument");
    if (a->n == 0) return 0;
    return done(sd::adamw_launch(*reinterpret_cast<const sd::AdamWArgs*>(a), sm_count(), (cudaStream_t)stream), "train_adamw_step");
}
