// tcgen05 (UMMA) + TMA implicit-GEMM convolution: interface used by the plan executor.
#pragma once
#include "common.cuh"

namespace scouter {

// Per-op cached launch state (tensor maps etc.); filled lazily by launch_conv_umma.
struct UmmaConvPlan {
    bool valid = false;
};

bool umma_conv_supported(const ConvArgs& a);
int launch_conv_umma(const ConvArgs& a, UmmaConvPlan& plan, cudaStream_t s);

}  // namespace scouter
