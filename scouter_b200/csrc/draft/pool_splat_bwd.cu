// DRAFT (row f1) -- see pool_splat_bwd.cuh.  Not listed in scouter_b200/_lib.py SOURCES: the library does not contain it.
#include <cuda_runtime.h>

#include "pool_splat_bwd.cuh"

namespace scouter_draft {

int maxpool_bwd_launch(const PoolBwdArgs& a, int sms, cudaStream_t s) { maxpool_bwd_kernel<<<sms * 8, 256, 0, s>>>(a); return (int)cudaGetLastError(); }
int avgpool2_bwd_launch(const PoolBwdArgs& a, int sms, cudaStream_t s) { avgpool2_bwd_kernel<<<sms * 8, 256, 0, s>>>(a); return (int)cudaGetLastError(); }
int avgpool3_bwd_launch(const PoolBwdArgs& a, int sms, cudaStream_t s) { avgpool3_bwd_kernel<<<sms * 8, 256, 0, s>>>(a); return (int)cudaGetLastError(); }
// reduce -> softmax -> (fc2 / bn1 / fc1 backward by the caller, producing d_gap) -> apply
// d_att[b, r, c] = sum_p d_out[b, p, c] * x2[b, p, r*C + c]: the thread-per-output kernel of pool_splat_bwd.cuh walks HW pixels per
// thread with B*2C threads in all (16 CTAs on layer 1).  Here CTA = (64-pixel chunk, image), thread = channels cc, cc+256, ...; the
// chunk sums are merged with fp32 atomics (d_att zeroed first).
static __global__ void __launch_bounds__(256) splat_bwd_reduce_chunk_kernel(SplatBwdArgs a) {
    const int b = blockIdx.y, C2 = 2 * a.C;
    const int p0 = blockIdx.x * 64, p1 = min(p0 + 64, a.HW);
    for (int cc = threadIdx.x; cc < C2; cc += 256) {
        const int c = cc % a.C;
        const float* dp = a.d_out + ((size_t)b * a.HW + p0) * a.C + c;
        const float* xp = a.x2 + ((size_t)b * a.HW + p0) * C2 + cc;
        float s0 = 0.f, s1 = 0.f;
        int p = p0;
        for (; p + 1 < p1; p += 2) {
            s0 = fmaf(__ldg(dp), __ldg(xp), s0);
            s1 = fmaf(__ldg(dp + a.C), __ldg(xp + C2), s1);
            dp += 2 * a.C;
            xp += 2 * C2;
        }
        if (p < p1) s0 = fmaf(__ldg(dp), __ldg(xp), s0);
        atomicAdd(a.d_att + (size_t)b * C2 + cc, s0 + s1);
    }
}

int splat_bwd_reduce_launch(const SplatBwdArgs& a, int sms, cudaStream_t s) {
    if (a.HW >= 256 && a.B <= 65535) {
        cudaMemsetAsync(a.d_att, 0, (size_t)a.B * 2 * a.C * sizeof(float), s);
        splat_bwd_reduce_chunk_kernel<<<dim3((a.HW + 63) / 64, a.B), 256, 0, s>>>(a);
        splat_bwd_softmax_kernel<<<sms, 256, 0, s>>>(a);
        return (int)cudaGetLastError();
    }
    splat_bwd_reduce_kernel<<<sms * 4, 256, 0, s>>>(a);
    splat_bwd_softmax_kernel<<<sms, 256, 0, s>>>(a);
    return (int)cudaGetLastError();
}
int splat_bwd_apply_launch(const SplatBwdArgs& a, int sms, cudaStream_t s) { splat_bwd_apply_kernel<<<sms * 8, 256, 0, s>>>(a); return (int)cudaGetLastError(); }

}  // namespace scouter_draft
