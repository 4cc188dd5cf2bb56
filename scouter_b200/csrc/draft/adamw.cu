// DRAFT (row f1) -- see adamw.cuh.  Not listed in scouter_b200/_lib.py SOURCES: the library does not contain it.
#include <cuda_runtime.h>

#include "adamw.cuh"

namespace scouter_draft {

int adamw_launch(const AdamWArgs& a, int sms, cudaStream_t stream) {
    adamw_kernel<<<sms * 8, 256, 0, stream>>>(a);      // grid-stride, a multiple of the SM count
    return (int)cudaGetLastError();
}

}  // namespace scouter_draft
