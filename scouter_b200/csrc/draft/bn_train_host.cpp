// Host emulation of the DRAFT train-mode BatchNorm (bn_train.cuh), built by tests/test_bn_train_draft.py: the three
// kernels run as loops over (cta, tid) / channels / float4s with the same decomposition a launch would use.
#include "bn_train.cuh"

extern "C" void bn_train_host(const scouter_draft::BnTrainArgs* a, int n_ctas, int nthreads) {
    for (int cta = 0; cta < n_ctas; ++cta)
        for (int tid = 0; tid < nthreads; ++tid) scouter_draft::bn_stats_partial(*a, cta, n_ctas, tid, nthreads);
    for (int c = 0; c < a->C; ++c) scouter_draft::bn_finalize(*a, c);
    for (long long i = 0; i < a->M * a->C / 4; ++i) scouter_draft::bn_apply(*a, i);
}

extern "C" void bn_train_backward_host(const scouter_draft::BnBwdArgs* a, int n_ctas, int nthreads) {
    for (int cta = 0; cta < n_ctas; ++cta)
        for (int tid = 0; tid < nthreads; ++tid) scouter_draft::bn_bwd_stats_partial(*a, cta, n_ctas, tid, nthreads);
    for (int c = 0; c < a->C; ++c) scouter_draft::bn_bwd_finalize(*a, c);
    for (long long i = 0; i < a->M * a->C / 4; ++i) scouter_draft::bn_bwd_apply(*a, i);
}
