// DRAFT (row f1): extern "C" launch wrappers around the draft kernels so that the SAME ctypes argument blocks the host
// emulation uses (tests/draft_emu.py) can be pointed at device memory and launched on a GPU.  Built only by
// tests/draft_emu.py (SCOUTER_DRAFT_BACKEND=gpu) into its own libscouter_draft.so -- never into libscouter_b200.so.
// Nothing here has run yet; `pytest -m gpu_draft` (tests/test_gpu_draft_kernels.py) is the first thing to do with it.
#include <cuda_runtime.h>

#include "adamw.cuh"
#include "bn_train.cuh"
#include "conv_wgrad.cuh"
#include "head_backward.cuh"
#include "pool_splat_bwd.cuh"

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

using namespace scouter_draft;

static int sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

extern "C" {
size_t draft_head_backward_scratch_floats(int n, int S, int L, int iters) { return head_backward_scratch_floats(n, S, L, iters); }
int draft_head_backward(const HeadBwdArgs* a) { return head_backward_launch(*a, 0); }
int draft_bn_train(const BnTrainArgs* a) { return bn_train_launch(*a, sm_count(), 0); }
int draft_bn_train_backward(const BnBwdArgs* a) { return bn_train_backward_launch(*a, sm_count(), 0); }
int draft_pool_bwd(const PoolBwdArgs* a, int kind) {
    return kind == 0 ? maxpool_bwd_launch(*a, sm_count(), 0) : kind == 1 ? avgpool2_bwd_launch(*a, sm_count(), 0) : avgpool3_bwd_launch(*a, sm_count(), 0);
}
int draft_splat_bwd(const SplatBwdArgs* a, int stage) {
    return stage == 0 ? splat_bwd_reduce_launch(*a, sm_count(), 0) : splat_bwd_apply_launch(*a, sm_count(), 0);
}
int draft_conv_wgrad(const WgradArgs* a) { return conv_wgrad_launch(*a, sm_count(), 0); }
int draft_conv_dgrad(const DgradArgs* a) { return conv_dgrad_launch(*a, sm_count(), 0); }
int draft_adamw(const AdamWArgs* a) { return adamw_launch(*a, sm_count(), 0); }
}
