// Host emulation of the DRAFT conv weight gradient (conv_wgrad.cuh), built by tests/test_conv_wgrad_draft.py with the
// same (element, K-split) decomposition a launch uses.
#include "conv_wgrad.cuh"

extern "C" void conv_wgrad_host(const scouter_draft::WgradArgs* a, int splits) {
    const long long elems = (long long)a->Cout * a->k * a->k * (a->Cin / a->groups);
    const long long M = (long long)a->B * a->Ho * a->Wo;
    const long long per = (M + splits - 1) / splits;
    for (int y = 0; y < splits; ++y) {
        const long long m0 = (long long)y * per, m1 = m0 + per < M ? m0 + per : M;
        for (long long e = 0; e < elems; ++e) scouter_draft::conv_wgrad_element(*a, e, m0, m1);
        if (a->db) for (int o = 0; o < a->Cout; ++o) scouter_draft::conv_bgrad_element(*a, o, m0, m1);
    }
}

extern "C" void conv_dgrad_host(const scouter_draft::DgradArgs* a) {
    const long long n = (long long)a->B * a->H * a->W * a->Cin;
    for (long long i = 0; i < n; ++i) scouter_draft::conv_dgrad_element(*a, i);
}
