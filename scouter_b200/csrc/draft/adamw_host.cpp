// Host emulation of the DRAFT AdamW step (adamw.cuh), built by tests/test_adamw_draft.py.
#include "adamw.cuh"

extern "C" void adamw_host(const scouter_draft::AdamWArgs* a) {
    for (size_t i = 0; i < a->n; ++i) scouter_draft::adamw_element(*a, i);
}
