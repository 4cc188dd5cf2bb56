// Row f1 -- see head_backward.cuh.
// Part of libscouter_b200.so through csrc/train.cu (scouter_train_head_backward).
#include <cuda_runtime.h>

#include "head_backward.cuh"

namespace scouter_draft {

// threads per image (the phases are element loops strided by the block size); 512 / 1024 were measured: no change of the step time,
// which at B = 32 is bound by the ~1300 launches the Python driver issues, not by this kernel
#ifndef SCOUTER_HB_THREADS
#define SCOUTER_HB_THREADS 256
#endif
__global__ void __launch_bounds__(SCOUTER_HB_THREADS) head_backward_kernel(HeadBwdArgs a) {
    head_backward_image(a, blockIdx.x, threadIdx.x, blockDim.x);
}

// attn_coef = g_attn_loss * power * m^(power-1) / (B*S*n),  m = sum_b attn_sum[b] / (B*S*n)   (slot_attention.py:93-96)
__global__ void head_backward_coef_kernel(const float* __restrict__ attn_sum, int B, int S, int n, float power,
                                          float g_attn_loss, float* __restrict__ coef) {
    if (blockIdx.x || threadIdx.x) return;
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += attn_sum[b];
    const double cnt = (double)B * S * n;
    *coef = (float)(g_attn_loss * power * pow(s / cnt, (double)power - 1.0) / cnt);
}

// scratch bytes per batch: B * head_bwd_layout(n, S, L, iters).total * 4
size_t head_backward_scratch_floats(int n, int S, int L, int iters) { return head_bwd_layout(n, S, L, iters).total; }

int head_backward_launch(const HeadBwdArgs& a, cudaStream_t stream) {
    head_backward_kernel<<<a.B, SCOUTER_HB_THREADS, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace scouter_draft
