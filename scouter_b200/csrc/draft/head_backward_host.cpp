// Host emulation of the DRAFT head backward (head_backward.cuh): the same body, one "thread" per image, barriers are
// no-ops, atomics are plain adds.  Built by tests/test_head_backward_draft.py with g++ and compared with
// oracle/head_backward.py -- a logic check of the transcription, not a GPU result.
#include "head_backward.cuh"

extern "C" size_t head_backward_scratch_floats(int n, int S, int L, int iters) {
    return scouter_draft::head_bwd_layout(n, S, L, iters).total;
}

extern "C" void head_backward_host(const scouter_draft::HeadBwdArgs* a, int nthreads_emulated) {
    // nthreads_emulated > 1 runs every phase as `nthreads` interleaved strided loops (tid = 0..nthreads-1 in turn, phase
    // by phase is not expressible here, so the whole image runs once per tid only when nthreads_emulated == 1)
    (void)nthreads_emulated;
    for (int b = 0; b < a->B; ++b) scouter_draft::head_backward_image(*a, b, 0, 1);
}
