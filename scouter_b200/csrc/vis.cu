// Row f3, the explanation output path: up-sampling of the uint8 slot maps to the input image's size and the attention
// ratio, on the device -- replaces the PNG round trip of test.py:33-35
//     np.array(Image.open(f'sloter/vis/slot_{id}.png').resize(image_raw.size, resample=Image.BILINEAR))
// and test.py:40-44 (attention_ratio = sum(map) / (h*w*255)).
//
// The arithmetic is Pillow's (Pillow==7.2.0, requirements.txt:14; src/libImaging/Resample.c, 8 bits per channel,
// single band): per axis a triangle filter of support max(1, in/out) around center = (xx + 0.5) * in/out, weights
// normalised in double, converted to 22-bit fixed point, and two separable passes -- horizontal first -- each
// clip8((2^21 + sum(pixel * k)) >> 22) with a uint8 image between them.  Byte work: the result is bit-identical to
// Pillow's (oracle/vis.py restates it; tests/golden/vis_upsample.npz holds Pillow's own outputs).
//
// The coefficients are recomputed per thread in IEEE double with the explicitly rounded intrinsics (__dadd_rn ...),
// which the compiler never contracts into FMAs -- the host library is compiled without FMA contraction, and a fused
// (xx + 0.5) * scale - support would change xmin at exact ties.  No table, no workspace, graph-capturable.  The maps are
// a few KB (C x 7 x 7 ... 9 x 9) and stay in L1/L2; the kernel is bound by its one coalesced byte store per thread.
#include "common.cuh"

namespace scouter {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS

struct Axis {
    double scale, support, ss;
    int in_size;
    __device__ Axis(int in, int out) : in_size(in) {
        scale = __ddiv_rn((double)in, (double)out);
        const double filterscale = scale < 1.0 ? 1.0 : scale;
        support = filterscale;                   // bilinear_filter.support (1.0) * filterscale
        ss = __ddiv_rn(1.0, filterscale);
    }
    // precompute_coeffs for output index xx: first tap, tap count, centre, sum of the raw weights
    __device__ void window(int xx, int& xmin, int& cnt, double& center, double& ww) const {
        center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
        xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
        if (xmin < 0) xmin = 0;
        int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
        if (xmax > in_size) xmax = in_size;
        cnt = xmax - xmin;
        ww = 0.0;
        for (int x = 0; x < cnt; ++x) ww = __dadd_rn(ww, raw(xmin + x, center));
    }
    __device__ double raw(int idx, double center) const {       // bilinear_filter((idx - center + 0.5) * ss)
        double t = fabs(__dmul_rn(__dadd_rn(__dsub_rn((double)idx, center), 0.5), ss));
        return t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
    }
    __device__ int coeff(int idx, double center, double ww) const {   // normalize_coeffs_8bpc
        double k = raw(idx, center);
        if (ww != 0.0) k = __ddiv_rn(k, ww);
        const double f = __dmul_rn(k, (double)(1 << kPrecisionBits));
        return k < 0.0 ? __double2int_rz(__dadd_rn(-0.5, f)) : __double2int_rz(__dadd_rn(0.5, f));
    }
};

__device__ __forceinline__ int clip8(int v) {
    v >>= kPrecisionBits;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// grid (ceil(out_w/256), out_h, count); thread = one output byte
__global__ void __launch_bounds__(256) vis_upsample_kernel(const uint8_t* __restrict__ maps, int h, int w, int out_h, int out_w,
                                                           uint8_t* __restrict__ out) {
    const int X = blockIdx.x * 256 + threadIdx.x, Y = blockIdx.y;
    if (X >= out_w) return;
    const uint8_t* src = maps + (size_t)blockIdx.z * h * w;
    const Axis ax(w, out_w), ay(h, out_h);
    int xmin, xcnt, ymin, ycnt;
    double xc, xww, yc, yww;
    ax.window(X, xmin, xcnt, xc, xww);
    ay.window(Y, ymin, ycnt, yc, yww);
    int acc = 1 << (kPrecisionBits - 1);
    for (int ky = 0; ky < ycnt; ++ky) {
        const uint8_t* row = src + (size_t)(ymin + ky) * w + xmin;
        int r = 1 << (kPrecisionBits - 1);
        for (int kx = 0; kx < xcnt; ++kx) r += (int)row[kx] * ax.coeff(xmin + kx, xc, xww);
        acc += clip8(r) * ay.coeff(ymin + ky, yc, yww);          // the horizontal pass lands in a uint8 image
    }
    out[((size_t)blockIdx.z * out_h + Y) * out_w + X] = (uint8_t)clip8(acc);
}

// one CTA per map: exact integer sum, one double division (test.py:43)
__global__ void __launch_bounds__(256) vis_ratio_kernel(const uint8_t* __restrict__ maps, int hw, double* __restrict__ ratios) {
    __shared__ unsigned long long part[8];
    const uint8_t* src = maps + (size_t)blockIdx.x * hw;
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < hw; i += 256) s += src[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i) t += part[i];
        ratios[blockIdx.x] = __ddiv_rn((double)t, (double)((long long)hw * 255));
    }
}

}  // namespace
}  // namespace scouter

extern "C" int scouter_vis_upsample_u8(const uint8_t* maps, int count, int h, int w, int out_h, int out_w, uint8_t* out,
                                       double* ratios, scouter_stream_t stream) {
    using namespace scouter;
    SC_CHECK_ARG(maps != nullptr && (out != nullptr || ratios != nullptr), SCOUTER_E_INVALID, "vis_upsample: NULL buffer");
    SC_CHECK_ARG(count >= 1 && count <= 65535 && h >= 1 && w >= 1 && (long long)h * w <= (1 << 24), SCOUTER_E_INVALID,
                 "vis_upsample: count=%d maps of %dx%d", count, h, w);
    if (out != nullptr) {
        SC_CHECK_ARG(out_h >= 1 && out_h <= 65535 && out_w >= 1 && out_w <= (1 << 24), SCOUTER_E_INVALID,
                     "vis_upsample: output size %dx%d", out_h, out_w);
        dim3 grid(cdiv(out_w, 256), out_h, count);
        vis_upsample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(maps, h, w, out_h, out_w, out);
        SC_LAUNCH_CHECK();
    }
    if (ratios != nullptr) {
        vis_ratio_kernel<<<count, 256, 0, (cudaStream_t)stream>>>(maps, h * w, ratios);
        SC_LAUNCH_CHECK();
    }
    return 0;
}
