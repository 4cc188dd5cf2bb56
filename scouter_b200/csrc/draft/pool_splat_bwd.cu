// DRAFT (row f1) -- see pool_splat_bwd.cuh.  Not listed in scouter_b200/_lib.py SOURCES: the library does not contain it.
#include <cuda_runtime.h>

#include "pool_splat_bwd.cuh"

namespace scouter_draft {

int maxpool_bwd_launch(const PoolBwdArgs& a, int sms, cudaStream_t s) { maxpool_bwd_kernel<<<sms * 8, 256, 0, s>>>(a); return (int)cudaGetLastError(); }
int avgpool2_bwd_launch(const PoolBwdArgs& a, int sms, cudaStream_t s) { avgpool2_bwd_kernel<<<sms * 8, 256, 0, s>>>(a); return (int)cudaGetLastError(); }
int avgpool3_bwd_launch(const PoolBwdArgs& a, int sms, cudaStream_t s) { avgpool3_bwd_kernel<<<sms * 8, 256, 0, s>>>(a); return (int)cudaGetLastError(); }
// reduce -> softmax -> (fc2 / bn1 / fc1 backward by the caller, producing d_gap) -> apply
int splat_bwd_reduce_launch(const SplatBwdArgs& a, int sms, cudaStream_t s) {
    splat_bwd_reduce_kernel<<<sms * 4, 256, 0, s>>>(a);
    splat_bwd_softmax_kernel<<<sms, 256, 0, s>>>(a);
    return (int)cudaGetLastError();
}
int splat_bwd_apply_launch(const SplatBwdArgs& a, int sms, cudaStream_t s) { splat_bwd_apply_kernel<<<sms * 8, 256, 0, s>>>(a); return (int)cudaGetLastError(); }

}  // namespace scouter_draft
