// DRAFT (row f1) -- see bn_train.cuh.  Not listed in scouter_b200/_lib.py SOURCES: the library does not contain it.
#include <cuda_runtime.h>

#include "bn_train.cuh"

namespace scouter_draft {

// Statistics passes: at most 4 CTAs per SM, at least 64 rows per CTA (small maps: fewer CTAs, fewer atomics).
static int stats_grid(long long M, int sms) {
    const long long want = (M + 63) / 64;
    return (int)(want < 1 ? 1 : (want < 4LL * sms ? want : 4LL * sms));
}

// `sums` must be zero on entry (cudaMemsetAsync by the caller); grids are multiples of the SM count.
int bn_train_launch(const BnTrainArgs& a, int sms, cudaStream_t stream) {
    bn_stats_kernel<<<stats_grid(a.M, sms), 256, 0, stream>>>(a);
    bn_finalize_kernel<<<(a.C + 127) / 128, 128, 0, stream>>>(a);
    bn_apply_kernel<<<sms * 8, 256, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

// out = [relu](bn(x) [+ residual]) backward; `sums` zero on entry.  NOTE the apply pass must not start before the
// statistics pass has read all of d_out when dx aliases d_out: stream order between the launches guarantees it.
int bn_train_backward_launch(const BnBwdArgs& a, int sms, cudaStream_t stream) {
    bn_bwd_stats_kernel<<<stats_grid(a.M, sms), 256, 0, stream>>>(a);
    bn_bwd_finalize_kernel<<<(a.C + 127) / 128, 128, 0, stream>>>(a);
    bn_bwd_apply_kernel<<<sms * 8, 256, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace scouter_draft
