#include "umma.cuh"

namespace scouter {

bool umma_conv_supported(const ConvArgs&) { return false; }

int launch_conv_umma(const ConvArgs&, UmmaConvPlan&, cudaStream_t) {
    set_error("tcgen05 convolution is not built in");
    return SCOUTER_E_UNSUPPORTED;
}

}  // namespace scouter
