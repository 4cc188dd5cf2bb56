// DRAFT (row f1) -- see conv_wgrad.cuh.  Not listed in scouter_b200/_lib.py SOURCES: the library does not contain it.
#include <cuda_runtime.h>

#include "conv_wgrad.cuh"

namespace scouter_draft {

int conv_wgrad_launch(const WgradArgs& a, int sms, cudaStream_t stream) {
    const long long elems = (long long)a.Cout * a.k * a.k * (a.Cin / a.groups);
    const int gx = (int)((elems + 255) / 256);
    int splits = (4 * sms + gx - 1) / gx;                  // about four CTAs per SM in total
    if (splits < 1) splits = 1;
    conv_wgrad_kernel<<<dim3(gx, splits), 256, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

int conv_dgrad_launch(const DgradArgs& a, int sms, cudaStream_t stream) {
    conv_dgrad_kernel<<<sms * 8, 256, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace scouter_draft
