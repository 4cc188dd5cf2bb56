// DRAFT (row f1) -- see conv_wgrad.cuh.  Not listed in scouter_b200/_lib.py SOURCES: the library does not contain it.
#include <cuda_runtime.h>

#include "conv_wgrad.cuh"

namespace scouter_draft {

// Tiled CUDA-core weight gradient (the naive thread-per-weight kernel of conv_wgrad.cuh spends 95 % of a training step):
// dW[o, (r,s), c] = sum_m dY[m, o] * X[m shifted by tap (r,s), c] is a GEMM whose K dimension is the pixel index m, and in
// NHWC both operands are K-major rows of contiguous channels.  CTA = BO output channels x BC input channels of one tap,
// 256 threads with (BO/16) x (BC/16) register tiles, 32 pixels per shared-memory stage, the pixel range split over gridDim.z
// and merged with fp32 atomics (like the naive kernel).
template <int BO, int BC>
__global__ void __launch_bounds__(256) conv_wgrad_tiled_kernel(WgradArgs a) {
    constexpr int KP = 32, TO = BO / 16, TC = BC / 16;
    __shared__ float dy_s[KP][BO + 4];
    __shared__ float x_s[KP][BC + 4];
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    const int o_tiles = cout_g / BO, c_tiles = cin_g / BC;
    int t = blockIdx.x;
    const int ct = t % c_tiles; t /= c_tiles;
    const int ot = t % o_tiles; t /= o_tiles;
    const int g = t;
    const int tap = blockIdx.y, r = tap / a.k, sx = tap - r * a.k;
    const int o0 = g * cout_g + ot * BO, c0 = g * cin_g + ct * BC;       // absolute channels
    const long long M = (long long)a.B * a.Ho * a.Wo;
    const long long per = ((M + gridDim.z - 1) / gridDim.z + KP - 1) / KP * KP;
    const long long m0 = (long long)blockIdx.z * per, m1 = m0 + per < M ? m0 + per : M;
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    float acc[TO][TC];
#pragma unroll
    for (int i = 0; i < TO; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;
    for (long long mb = m0; mb < m1; mb += KP) {
        // stage 32 pixels: dY rows (64 channels) and the tap-shifted X rows (BC channels); out-of-image taps are zeros
        for (int i = tid; i < KP * (BO / 4); i += 256) {
            const int k = i / (BO / 4), q = i % (BO / 4);
            const long long m = mb + k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < m1) v = __ldg(reinterpret_cast<const float4*>(a.dy + (size_t)m * a.Cout + o0 + 4 * q));
            *reinterpret_cast<float4*>(&dy_s[k][4 * q]) = v;
        }
        for (int i = tid; i < KP * (BC / 4); i += 256) {
            const int k = i / (BC / 4), q = i % (BC / 4);
            const long long m = mb + k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < m1) {
                const int xo = (int)(m % a.Wo);
                const long long qq = m / a.Wo;
                const int yo = (int)(qq % a.Ho), b = (int)(qq / a.Ho);
                const int yi = yo * a.stride + r - a.pad, xi = xo * a.stride + sx - a.pad;
                if (yi >= 0 && yi < a.H && xi >= 0 && xi < a.W)
                    v = __ldg(reinterpret_cast<const float4*>(a.x + (((size_t)b * a.H + yi) * a.W + xi) * a.Cin + c0 + 4 * q));
            }
            *reinterpret_cast<float4*>(&x_s[k][4 * q]) = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < KP; ++k) {
            float dv[TO], xv[TC];
#pragma unroll
            for (int i = 0; i < TO; ++i) dv[i] = dy_s[k][ty * TO + i];
#pragma unroll
            for (int j = 0; j < TC; ++j) xv[j] = x_s[k][tx * TC + j];
#pragma unroll
            for (int i = 0; i < TO; ++i)
#pragma unroll
                for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TO; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) {
            const int o = o0 + TO * ty + i, c = ct * BC + tx * TC + j;             // c inside the group
            atomicAdd(a.dw + (((size_t)o * a.k + r) * a.k + sx) * cin_g + c, acc[i][j]);
        }
}

// Few input channels (the stem: Cin = 3 or 1, k = 3): the flattened (r, s, c) index j < 32 plays the channel role.  CTA = 32 output
// channels x all taps, 64 pixels per shared-memory stage; thread = (output channel, taps jj, jj+8, jj+16, jj+24).  The naive kernel
// had 864 threads' worth of parallelism for 401k pixels (2.1 ms of a 34 ms step).
__global__ void __launch_bounds__(256) conv_wgrad_smallc_kernel(WgradArgs a, int pix_per_cta) {
    __shared__ float dy_s[64][33];
    __shared__ float x_s[64][33];
    const int taps = a.k * a.k * a.Cin;                 // <= 32
    const int o0 = blockIdx.y * 32;
    const long long M = (long long)a.B * a.Ho * a.Wo;
    const long long m0 = (long long)blockIdx.x * pix_per_cta, m1 = m0 + pix_per_cta < M ? m0 + pix_per_cta : M;
    const int tid = threadIdx.x, o = tid & 31, jj = tid >> 5;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long mb = m0; mb < m1; mb += 64) {
        for (int i = tid; i < 64 * 32; i += 256) {
            const int pl = i >> 5, q = i & 31;
            const long long m = mb + pl;
            float dv = 0.f, xv = 0.f;
            if (m < m1) {
                dv = __ldg(a.dy + (size_t)m * a.Cout + o0 + q);
                if (q < taps) {
                    const int c = q % a.Cin, rs = q / a.Cin, sx = rs % a.k, r = rs / a.k;
                    const int xo = (int)(m % a.Wo);
                    const long long t = m / a.Wo;
                    const int yo = (int)(t % a.Ho), b = (int)(t / a.Ho);
                    const int yi = yo * a.stride + r - a.pad, xi = xo * a.stride + sx - a.pad;
                    if (yi >= 0 && yi < a.H && xi >= 0 && xi < a.W) xv = __ldg(a.x + (((size_t)b * a.H + yi) * a.W + xi) * a.Cin + c);
                }
            }
            dy_s[pl][q] = dv;
            x_s[pl][q] = xv;
        }
        __syncthreads();
#pragma unroll 8
        for (int pl = 0; pl < 64; ++pl) {
            const float d = dy_s[pl][o];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(d, x_s[pl][jj + 8 * q], acc[q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = jj + 8 * q;
        if (j < taps) atomicAdd(a.dw + (size_t)(o0 + o) * taps + j, acc[q]);
    }
}

int conv_wgrad_launch(const WgradArgs& a, int sms, cudaStream_t stream) {
    if (!a.db && a.groups == 1 && a.k * a.k * a.Cin <= 32 && a.Cout % 32 == 0) {
        const long long M = (long long)a.B * a.Ho * a.Wo;
        const int o_tiles = a.Cout / 32;
        long long ctas = (8LL * sms + o_tiles - 1) / o_tiles;
        long long per = ((M + ctas - 1) / ctas + 63) / 64 * 64;
        if (per < 256) per = 256;
        const int gx = (int)((M + per - 1) / per);
        conv_wgrad_smallc_kernel<<<dim3(gx, o_tiles), 256, 0, stream>>>(a, (int)per);
        return (int)cudaGetLastError();
    }
    {
        const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
        const long long M = (long long)a.B * a.Ho * a.Wo;
        // big maps without a bias: the tiled GEMM form (fc convs on (B,1,1,C) maps and the 1..4-channel stem stay on the naive kernel)
        if (!a.db && cout_g % 32 == 0 && cin_g % 32 == 0 && M >= 1024 && ((a.Cin | a.Cout) & 3) == 0) {
            const int bc = cin_g % 64 == 0 ? 64 : 32, bo = cout_g % 64 == 0 ? 64 : 32;
            const int ctas = a.groups * (cout_g / bo) * (cin_g / bc) * a.k * a.k;
            int splits = (6 * sms + ctas - 1) / ctas;
            const int max_splits = (int)((M + 255) / 256);            // at least 256 pixels per CTA
            if (splits > max_splits) splits = max_splits;
            if (splits < 1) splits = 1;
            if (splits > 65535) splits = 65535;
            dim3 grid(a.groups * (cout_g / bo) * (cin_g / bc), a.k * a.k, splits);
            if (bo == 64 && bc == 64) conv_wgrad_tiled_kernel<64, 64><<<grid, 256, 0, stream>>>(a);
            else if (bo == 64) conv_wgrad_tiled_kernel<64, 32><<<grid, 256, 0, stream>>>(a);
            else if (bc == 64) conv_wgrad_tiled_kernel<32, 64><<<grid, 256, 0, stream>>>(a);
            else conv_wgrad_tiled_kernel<32, 32><<<grid, 256, 0, stream>>>(a);
            return (int)cudaGetLastError();
        }
    }
    const long long elems = (long long)a.Cout * a.k * a.k * (a.Cin / a.groups);
    const int gx = (int)((elems + 255) / 256);
    int splits = (4 * sms + gx - 1) / gx;                  // about four CTAs per SM in total
    if (splits < 1) splits = 1;
    conv_wgrad_kernel<<<dim3(gx, splits), 256, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

int conv_dgrad_launch(const DgradArgs& a, int sms, cudaStream_t stream) {
    conv_dgrad_kernel<<<sms * 8, 256, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace scouter_draft
