// Shared epilogue of the tcgen05 conv kernels: a thread holds 16 consecutive output channels of ONE output row
// (that is how tcgen05.ld 32x32b hands out the accumulator: lane = row).  Writing them directly makes every 16-byte
// store of a warp hit a different 128-byte line (32 L1 wavefronts per instruction), which is what bounds the
// memory-bound 1x1 convs.  Here the 32x16 block goes through a per-warp padded scratch so that each instruction
// covers 8 rows x 64 contiguous bytes (8 wavefronts); bias, residual and ReLU are applied on the coalesced side.
#pragma once
#include "common.cuh"

namespace scouter {

struct EpiOut {
    const float* bias;
    const float* res;
    float* out;
    int Cout, relu, round_out;
};

constexpr int EPI_LD = 20;                          // floats per scratch row (16 + 4 pad: conflict-free both ways)
constexpr int EPI_SCRATCH_FLOATS = 32 * EPI_LD;     // per epilogue warp

// my_row: global output row of this lane (-1 = nothing to store); ch: first of the 16 channels; v: the 16 values.
__device__ __forceinline__ void epi_emit16(const EpiOut& e, float* scratch, int lane, int my_row, int ch, const float* v) {
    float4* srow = reinterpret_cast<float4*>(scratch + lane * EPI_LD);
#pragma unroll
    for (int j = 0; j < 4; ++j) srow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int quad = lane >> 3, rsub = lane & 7;
    const int c = ch + quad * 4;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias) bv = __ldg(reinterpret_cast<const float4*>(e.bias + c));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + rsub;
        const int gr = __shfl_sync(0xffffffffu, my_row, r);
        float4 x = *reinterpret_cast<const float4*>(scratch + r * EPI_LD + quad * 4);
        if (gr >= 0) {
            const size_t o = (size_t)gr * e.Cout + c;
            x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
            if (e.res) {
                const float4 rv = __ldg(reinterpret_cast<const float4*>(e.res + o));
                x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w;
            }
            if (e.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            if (e.round_out) { x.x = to_tf32(x.x); x.y = to_tf32(x.y); x.z = to_tf32(x.z); x.w = to_tf32(x.w); }
            *reinterpret_cast<float4*>(e.out + o) = x;
        }
    }
    __syncwarp();
}

}  // namespace scouter
