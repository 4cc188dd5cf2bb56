// tcgen05 (UMMA) + TMA implicit-GEMM convolution: interface used by the plan executor.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace scouter {

// Per-op cached launch state: the two TMA descriptors and the spatial tile they were built for.
struct UmmaConvPlan {
    bool valid = false;
    bool halo = false;
    bool presplit = false;
    CUtensorMap tmA, tmB, tmB2, tmO, tmR;
    const float* out = nullptr;
    const float* res = nullptr;
    int ksplit = 0;
    const float* in = nullptr;
    const float* w = nullptr;
    int B = 0, H = 0, W = 0, Cin = 0, Cout = 0, kh = 0, groups = 0, BN = 0;
    int Wb = 0, Hb = 0, Nb = 0;
};

bool umma_conv_supported(const ConvArgs& a);
int launch_conv_umma(const ConvArgs& a, UmmaConvPlan& plan, cudaStream_t s);
// 3x3 from one halo tile per channel block (umma_halo.cu); needs a.split and a.w_rem
bool halo_conv_supported(const ConvArgs& a);
int launch_conv_halo(const ConvArgs& a, UmmaConvPlan& plan, cudaStream_t s);
// slots per image of ConvArgs::gap_part when the halo kernel takes this conv (tiles per image x 4 epilogue warps), else 0
int halo_gap_slots(const ConvArgs& a);
// dispatch: halo kernel when it applies, else the tap-reload / flat kernel
inline bool tc_conv_supported(const ConvArgs& a) { return halo_conv_supported(a) || umma_conv_supported(a); }
inline int launch_conv_tc(const ConvArgs& a, UmmaConvPlan& plan, cudaStream_t s) {
    return halo_conv_supported(a) ? launch_conv_halo(a, plan, s) : launch_conv_umma(a, plan, s);
}

}  // namespace scouter
