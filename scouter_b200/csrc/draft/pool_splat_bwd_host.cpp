// Host emulation of the DRAFT pooling / split-attention backward (pool_splat_bwd.cuh), built by
// tests/test_pool_splat_bwd_draft.py: every kernel is a loop over its index space.
#include "pool_splat_bwd.cuh"

using namespace scouter_draft;

extern "C" void pool_bwd_host(const PoolBwdArgs* a, int kind) {
    const long long n = (long long)a->B * a->H * a->W * a->C;
    for (long long i = 0; i < n; ++i) {
        if (kind == 0) maxpool_bwd(*a, i);
        else if (kind == 1) avgpool2_bwd(*a, i);
        else avgpool3_bwd(*a, i);
    }
}

extern "C" void splat_bwd_host(const SplatBwdArgs* a, int stage) {
    if (stage == 0) {
        for (long long i = 0; i < (long long)a->B * 2 * a->C; ++i) splat_bwd_reduce(*a, i);
        for (long long i = 0; i < (long long)a->B * a->C; ++i) splat_bwd_softmax(*a, i);
    } else {
        for (long long i = 0; i < (long long)a->B * a->HW * 2 * a->C; ++i) splat_bwd_apply(*a, i);
    }
}
