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
    // tensor-core operands of the fused head's to_k MLP (K-major UMMA B tiles, fetched by TMA):
    //   raw:  (L*64, 64) fp32, row l*64+o = to_k[l].weight[o][:]
    //   pair: (L*128, 64) bf16, rows l*128+[0,64) = bf16(W), rows l*128+[64,128) = bf16(W - trunc19(W))
    __host__ __device__ size_t tok_w_raw(int l) const { return gru_bhh() + XG + (size_t)l * XD * XD; }
    __host__ __device__ size_t tok_w_pair(int l) const { return tok_w_raw(L) + (size_t)l * XD * XD; }   // 8192 bf16 = 4096 floats per layer
    __host__ __device__ size_t total() const { return tok_w_pair(L); }
};

int validate_xslot_desc(const scouter_xslot_desc_t* d);

// Fast loop kernel (xslot_fast.cu): contiguous tokens + PE table; optionally finishes a split-K projection.
struct XSlotFastIO {
    int batch, n;
    const float* x;          // (B, n, 64) or null when xpart is given
    const float* xpart;      // (nsplit, B*n, 64) partial projections (no bias), or null
    const float* conv_bias;  // (64), with xpart
    long long split_stride;  // floats between slabs
    int nsplit;
    const float* pe;         // (n, 64)
    float* x_out;            // optional (B, n, 64): the finished projection (tests)
    float* logits;
    float* attn;
    float* attn_sum;
};
bool xslot_fast_supported(const scouter_xslot_desc_t* d, int n);
int xslot_fast_launch(const scouter_xslot_desc_t* d, const void* packed, const XSlotFastIO& io, cudaStream_t s);


// Whole head in one kernel (head_fused.cu): projection on the tensor cores feeding the loop through shared memory.
bool head_fused_supported(const scouter_xslot_desc_t* d, int batch, int n, int channel);
size_t head_fused_workspace_bytes(int channel);   // bf16 [W ; W_r] of conv1x1.weight when the caller does not supply it
int head_fused_launch(const scouter_xslot_desc_t* d, const void* packed, const scouter_head_io_t* io, const float* feat_nhwc,
                      void* workspace, cudaStream_t s);

}  // namespace scouter
