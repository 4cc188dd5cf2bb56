// Packed-parameter layout of the xSlot module, shared by the pack kernel and the forward kernels.
#pragma once
#include "common.cuh"

namespace scouter {

constexpr int XD = 64;        // hidden dim (train.py default hidden_dim=64; the only one implemented)
constexpr int XG = 3 * XD;    // GRU gate rows

// Offsets in floats inside the packed buffer.
struct XSlotPacked {
    int S, L;
    __host__ __device__ size_t slots() const { return 0; }                               // (S, 64)
    __host__ __device__ size_t tok_wt(int l) const { return (size_t)S * XD + (size_t)l * (XD * XD + XD); }  // WT[e][o]
    __host__ __device__ size_t tok_b(int l) const { return tok_wt(l) + XD * XD; }        // (64)
    __host__ __device__ size_t gru_wih_t() const { return tok_wt(L); }                   // [e][g] (64,192)
    __host__ __device__ size_t gru_whh_t() const { return gru_wih_t() + XD * XG; }
    __host__ __device__ size_t gru_bih() const { return gru_whh_t() + XD * XG; }         // (192)
    __host__ __device__ size_t gru_bhh() const { return gru_bih() + XG; }
    // tensor-core operands (SCOUTER_MATH_TC): row-major [out][in] hi/lo tf32 splits, K-major for UMMA B
    __host__ __device__ size_t tok_w_hi(int l) const { return gru_bhh() + XG + (size_t)l * 2 * XD * XD; }
    __host__ __device__ size_t tok_w_lo(int l) const { return tok_w_hi(l) + XD * XD; }
    __host__ __device__ size_t total() const { return tok_w_hi(L); }
};

int validate_xslot_desc(const scouter_xslot_desc_t* d);

}  // namespace scouter
