// Shared helpers for libscouter_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/scouter_b200.h"

namespace scouter {

void set_error(const char* fmt, ...);

#define SC_CHECK_ARG(cond, code, ...)          \
    do {                                       \
        if (!(cond)) {                         \
            ::scouter::set_error(__VA_ARGS__); \
            return (code);                     \
        }                                      \
    } while (0)

#define SC_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::scouter::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                                 __FILE__, __LINE__);                                        \
            return (int)_e;                                                                  \
        }                                                                                    \
    } while (0)

#define SC_LAUNCH_CHECK()                                                                    \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            ::scouter::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                        \
            return (int)_e;                                                                  \
        }                                                                                    \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Round-to-nearest fp32 -> tf32 (10-bit mantissa), result kept in an fp32 container.
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---- launchers implemented in the .cu files (all return 0 / cudaError_t) -------------------------

struct ConvArgs {
    const float* in;    // NHWC (B,H,W,Cin)
    const float* w;     // (Cout, kh, kw, Cin/groups)
    const float* bias;  // (Cout) or null
    const float* res;   // NHWC (B,Ho,Wo,Cout) or null
    float* out;         // NHWC (B,Ho,Wo,Cout)
    int B, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad, groups, relu;
    int round_out;  // round the stored activations to tf32 (SCOUTER_MATH_TC: the next conv's MMA reads them as tf32)
    int split;      // tcgen05 only: error-compensated product (operands split into fp16 + remainder)
    const void* w_rem;   // optional host-pre-split 16-bit weights [fp16 W ; bf16 (W - fp16 W)], (2*Cout, kh, kw, Cin/g)
    int ksplit;          // tcgen05 flat kernel only: > 1 writes `ksplit` raw partial-sum slabs (M*Cout floats apart) to out
    float* gap_part;     // halo kernel only, optional: per-(image, tile, epilogue warp) column sums of the stored output,
    int gap_slots;       //   (B, gap_slots, Cout) -- the split-attention GAP's partial sums (split_attn.py:64-66), see halo_gap_slots()
};
int launch_conv_simt(const ConvArgs& a, cudaStream_t s);

struct StemArgs {
    const float* in;   // NCHW (B,Cin,H,W)
    const float* w;    // (Cout, kh, kw, Cin)
    const float* bias;
    float* out;        // NHWC
    int B, H, W, Cin, Ho, Wo, Cout, k, stride, pad, relu;
    int round_out;
    const float* w_host;  // host copies of w / bias (plan-owned) for the constant-bank stem kernel, or null
    const float* b_host;
    int tc;               // SCOUTER_MATH_TC: the tensor-core stem (stem_ts.cu) may take it
};
int launch_stem_conv(const StemArgs& a, cudaStream_t s);
bool stem_ts_supported(const StemArgs& a);
int launch_stem_ts(const StemArgs& a, cudaStream_t s);

int launch_maxpool(const float* in, float* out, int B, int H, int W, int C, int Ho, int Wo, int k, int stride,
                   int pad, cudaStream_t s);
int launch_avgpool(const float* in, float* out, int B, int H, int W, int C, int Ho, int Wo, int k, int stride,
                   int pad, int count_include_pad, int round_out, cudaStream_t s);
int splat_gap_splits(int B, int HW);  // pixel slices of the split-attention GAP (workspace = B * splits * 2C floats)
int launch_splat_gap(const float* in, float* part, float* gap, int B, int HW, int C /*per radix*/, cudaStream_t s);
// finish only: `part` (B, nslots, 2C) was written by the producing conv's epilogue (ConvArgs::gap_part)
int launch_splat_gap_finish(const float* part, float* gap, int B, int HW, int C /*per radix*/, int nslots, cudaStream_t s);
int launch_splat_apply(const float* in, const float* logit, float* out, int B, int H, int W, int C, int Ho, int Wo,
                       int avd, int round_out, cudaStream_t s);
int launch_gap(const float* in, float* out, int B, int HW, int C, cudaStream_t s);
int launch_preprocess_u8(const uint8_t* img, int B, int H, int W, int C, const double* mean, const double* stdv, float* out,
                         cudaStream_t s);
int launch_nhwc_to_nchw(const float* in, float* out, int B, int HW, int C, cudaStream_t s);
int launch_nchw_to_nhwc(const float* in, float* out, int B, int C, int HW, cudaStream_t s);

}  // namespace scouter
