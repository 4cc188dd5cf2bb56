// Exact-fp32 backbone kernels on CUDA cores (SCOUTER_MATH_FP32): the bit-for-bit-auditable path that the
// tensor-core kernels are checked against on the device, and the fallback for shapes the tcgen05 path
// does not take (strided 3x3, tiny channel counts).  NHWC activations, OHWI weights, BN already folded.
//
// Reference ops replaced (eval mode): nn.Conv2d + BatchNorm2d + ReLU chains of timm/models/resnet.py:401-420,
// resnest.py:111-143, split_attn.py:54-80, the pools at resnet.py:300,420 and resnest.py:101.
#include <algorithm>

#include "common.cuh"

namespace scouter {

// ------------------------------------------------------------------------------------------------
// Generic implicit-GEMM convolution:  M = B*Ho*Wo, N = Cout/groups, K = kh*kw*Cin/groups.
// 128x64 output tile per 256-thread CTA, 8x4 register micro-tile, K in chunks of 16 staged through
// double-buffered shared memory.  Requires (Cin/groups) % 16 == 0 so a K-chunk sits inside one tap.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
constexpr int LDA = BM + 4, LDB = BN + 4;

__global__ void __launch_bounds__(256) conv_igemm_f32_kernel(ConvArgs p) {
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int g = blockIdx.z;
    const int cin_g = p.Cin / p.groups, cout_g = p.Cout / p.groups;
    const int M = p.B * p.Ho * p.Wo;
    const int K = p.kh * p.kw * cin_g;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // A loader: 2 float4 per thread: rows (tid/4) and (tid/4 + 64), k-quad tid%4.
    const int kq = tid % 4;
    int a_hi0[2], a_wi0[2];
    long long a_base[2];
    bool a_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int m = m0 + tid / 4 + i * 64;
        a_ok[i] = m < M;
        int mm = a_ok[i] ? m : 0;
        int wo = mm % p.Wo;
        int t = mm / p.Wo;
        int ho = t % p.Ho;
        int b = t / p.Ho;
        a_hi0[i] = ho * p.stride - p.pad;
        a_wi0[i] = wo * p.stride - p.pad;
        a_base[i] = (long long)b * p.H * p.W * p.Cin + (long long)g * cin_g + kq * 4;
    }
    // B loader: 1 float4 per thread: n = tid/4, k-quad tid%4.
    const int bn = n0 + tid / 4;
    const bool b_ok = bn < cout_g;
    const float* wrow = p.w + (long long)(g * cout_g + (b_ok ? bn : 0)) * K + kq * 4;

    float4 ra[2], rb;
    auto load_chunk = [&](int kc) {
        int k0 = kc * BK;
        int tap = k0 / cin_g;
        int c0 = k0 - tap * cin_g;
        int r = tap / p.kw, s = tap - r * p.kw;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int hi = a_hi0[i] + r, wi = a_wi0[i] + s;
            bool ok = a_ok[i] && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            ra[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.in + a_base[i] + ((long long)hi * p.W + wi) * p.Cin + c0))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        rb = b_ok ? __ldg(reinterpret_cast<const float4*>(wrow + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int row = tid / 4 + i * 64;
            As[buf][kq * 4 + 0][row] = ra[i].x;
            As[buf][kq * 4 + 1][row] = ra[i].y;
            As[buf][kq * 4 + 2][row] = ra[i].z;
            As[buf][kq * 4 + 3][row] = ra[i].w;
        }
        int col = tid / 4;
        Bs[buf][kq * 4 + 0][col] = rb.x;
        Bs[buf][kq * 4 + 1][col] = rb.y;
        Bs[buf][kq * 4 + 2][col] = rb.z;
        Bs[buf][kq * 4 + 3][col] = rb.w;
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = K / BK;
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int kc = 0; kc < nk; ++kc) {
        const int cur = kc & 1;
        if (kc + 1 < nk) load_chunk(kc + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * TM]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][ty * TM + 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * TN]);
            float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[TN] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kc + 1 < nk) store_chunk(cur ^ 1);
        __syncthreads();
    }

    // Epilogue: + bias (+ residual) (ReLU), NHWC store.
    const int n = n0 + tx * TN;
    const bool vec = (cout_g % 4 == 0) && (n + 3 < cout_g);
    float bv[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = (p.bias && n + j < cout_g) ? __ldg(p.bias + g * cout_g + n + j) : 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= M) continue;
        long long o = (long long)m * p.Cout + g * cout_g + n;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) v[j] = acc[i][j] + bv[j];
        if (vec) {
            if (p.res) {
                float4 r = __ldg(reinterpret_cast<const float4*>(p.res + o));
                v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
            }
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < TN; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.round_out) {
#pragma unroll
                for (int j = 0; j < TN; ++j) v[j] = to_tf32(v[j]);
            }
            *reinterpret_cast<float4*>(p.out + o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                if (n + j >= cout_g) continue;
                float t = v[j];
                if (p.res) t += __ldg(p.res + o + j);
                if (p.relu) t = fmaxf(t, 0.f);
                if (p.round_out) t = to_tf32(t);
                p.out[o + j] = t;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Stem convolution: NCHW input with <= 4 channels (3 RGB, 1 MNIST), k x k, writes NHWC.
// One thread = one output pixel x 4 output channels; the (Cout,k,k,Cin) filter bank sits in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_conv_kernel(StemArgs p) {
    extern __shared__ __align__(16) float sw[];  // [k*k*Cin][Cout]  (tap-major so a thread's 4 channels are contiguous)
    const int taps = p.k * p.k * p.Cin;
    for (int i = threadIdx.x; i < taps * p.Cout; i += blockDim.x) {
        int co = i / taps, t = i - co * taps;  // source layout (Cout, k, k, Cin)
        sw[t * p.Cout + co] = p.w[i];
    }
    __syncthreads();
    const int cq = p.Cout / 4;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)p.B * p.Ho * p.Wo * cq;
    if (idx >= total) return;
    int q = (int)(idx % cq);
    long long pix = idx / cq;
    int wo = (int)(pix % p.Wo);
    long long t2 = pix / p.Wo;
    int ho = (int)(t2 % p.Ho);
    int b = (int)(t2 / p.Ho);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int hi0 = ho * p.stride - p.pad, wi0 = wo * p.stride - p.pad;
    for (int r = 0; r < p.k; ++r) {
        int hi = hi0 + r;
        if (hi < 0 || hi >= p.H) continue;
        for (int s = 0; s < p.k; ++s) {
            int wi = wi0 + s;
            if (wi < 0 || wi >= p.W) continue;
            for (int c = 0; c < p.Cin; ++c) {
                float v = __ldg(p.in + (((long long)b * p.Cin + c) * p.H + hi) * p.W + wi);
                const float4 wv = *reinterpret_cast<const float4*>(&sw[((r * p.k + s) * p.Cin + c) * p.Cout + q * 4]);
                acc[0] = fmaf(v, wv.x, acc[0]);
                acc[1] = fmaf(v, wv.y, acc[1]);
                acc[2] = fmaf(v, wv.z, acc[2]);
                acc[3] = fmaf(v, wv.w, acc[3]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (p.bias) acc[j] += __ldg(p.bias + q * 4 + j);
        if (p.relu) acc[j] = fmaxf(acc[j], 0.f);
        if (p.round_out) acc[j] = to_tf32(acc[j]);
    }
    *reinterpret_cast<float4*>(p.out + pix * p.Cout + q * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// ------------------------------------------------------------------------------------------------
// Tiled 3x3 stride-2 pad-1 stem (the deep stem's conv1.0 and the MNIST stem): the filter bank travels as a
// kernel parameter, i.e. in the constant bank, so every FFMA takes its weight as an immediate constant operand;
// the CTA stages a (CIN, 17, 65) input patch in shared memory and each thread produces one output pixel x COUT.
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct StemConst {
    float w[9 * CIN * COUT];  // [tap = (r*3+s)*CIN + c][cout]
    float b[COUT];
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(256) stem3x3s2_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                        int Ho, int Wo, int relu, int round_out,
                                                        const __grid_constant__ StemConst<CIN, COUT> cw) {
    constexpr int TH = 8, TW = 32, PH = 2 * TH + 1, PW = 2 * TW + 1, LDP = PW + 2;
    __shared__ float patch[CIN][PH][LDP];
    const int b = blockIdx.z;
    const int h0 = blockIdx.y * TH, w0 = blockIdx.x * TW;
    const int hi0 = 2 * h0 - 1, wi0 = 2 * w0 - 1;
    for (int i = threadIdx.x; i < CIN * PH * PW; i += 256) {
        int c = i / (PH * PW), r = (i / PW) % PH, x = i % PW;
        int hi = hi0 + r, wi = wi0 + x;
        float v = 0.f;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(in + (((long long)b * CIN + c) * H + hi) * W + wi);
        patch[c][r][x] = v;
    }
    __syncthreads();
    const int ty = threadIdx.x / TW, tx = threadIdx.x % TW;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const float v = patch[c][2 * ty + r][2 * tx + s];
#pragma unroll
                for (int co = 0; co < COUT; ++co) acc[co] = fmaf(v, cw.w[((r * 3 + s) * CIN + c) * COUT + co], acc[co]);
            }
    // The tile's 32 pixels of one output row are one contiguous run of 32*COUT floats in NHWC: stage the results in shared
    // memory (pixel stride COUT+4 floats: conflict-free 16-byte stores) and write whole 512-byte segments per warp
    // instruction instead of 32 scattered 16-byte pieces.
    extern __shared__ __align__(16) float stage[];
    constexpr int LDS_ = COUT + 4;
    float* sp = stage + threadIdx.x * LDS_;
#pragma unroll
    for (int co = 0; co < COUT; co += 4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = acc[co + j] + cw.b[co + j];
            if (relu) v[j] = fmaxf(v[j], 0.f);
            if (round_out) v[j] = to_tf32(v[j]);
        }
        *reinterpret_cast<float4*>(sp + co) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    constexpr int Q = COUT / 4;                 // float4 per pixel
    for (int i = threadIdx.x; i < TH * TW * Q; i += 256) {
        const int pix = i / Q, c4 = i - pix * Q;
        const int py = pix / TW, px = pix - py * TW;
        const int ho = h0 + py, wo = w0 + px;
        if (ho < Ho && wo < Wo)
            *reinterpret_cast<float4*>(out + (((long long)b * Ho + ho) * Wo + wo) * COUT + c4 * 4) =
                *reinterpret_cast<const float4*>(stage + pix * LDS_ + c4 * 4);
    }
}

template <int CIN, int COUT>
int launch_stem_tiled(const StemArgs& a, cudaStream_t s) {
    static_assert(sizeof(StemConst<CIN, COUT>) <= 3800, "filter bank must fit the 4 KB kernel-parameter space");
    StemConst<CIN, COUT> cw;
    // host copy is laid out (Cout, 3, 3, Cin): transpose to tap-major
    for (int co = 0; co < COUT; ++co)
        for (int t = 0; t < 9 * CIN; ++t) cw.w[t * COUT + co] = a.w_host[co * 9 * CIN + t];
    for (int co = 0; co < COUT; ++co) cw.b[co] = a.b_host ? a.b_host[co] : 0.f;
    dim3 grid(cdiv(a.Wo, 32), cdiv(a.Ho, 8), a.B);
    constexpr int stage_bytes = 256 * (COUT + 4) * 4;
    if (stage_bytes + (int)sizeof(float) * CIN * 17 * 67 > 48 * 1024)
        SC_CUDA(cudaFuncSetAttribute(stem3x3s2_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage_bytes));
    stem3x3s2_kernel<CIN, COUT><<<grid, 256, stage_bytes, s>>>(a.in, a.out, a.H, a.W, a.Ho, a.Wo, a.relu, a.round_out, cw);
    SC_LAUNCH_CHECK();
    return 0;
}

// Tiled transposes between (B, HW, C) and (B, C, HW).
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Ccols) {
    // in: (B, R, Ccols) -> out: (B, Ccols, R)
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const float* src = in + (long long)b * R * Ccols;
    float* dst = out + (long long)b * R * Ccols;
    int c = blockIdx.x * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = blockIdx.y * 32 + i;
        if (r < R && c < Ccols) tile[i][threadIdx.x] = src[(long long)r * Ccols + c];
    }
    __syncthreads();
    int r2 = blockIdx.y * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c2 = blockIdx.x * 32 + i;
        if (r2 < R && c2 < Ccols) dst[(long long)c2 * R + r2] = tile[threadIdx.x][i];
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Launchers
// ------------------------------------------------------------------------------------------------
int launch_conv_simt(const ConvArgs& a, cudaStream_t s) {
    SC_CHECK_ARG(a.groups >= 1 && a.Cin % a.groups == 0 && a.Cout % a.groups == 0, SCOUTER_E_INVALID,
                 "conv: channels (%d,%d) not divisible by groups %d", a.Cin, a.Cout, a.groups);
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    SC_CHECK_ARG(cin_g % 16 == 0, SCOUTER_E_UNSUPPORTED, "conv: Cin/groups = %d is not a multiple of 16", cin_g);
    const long long M = (long long)a.B * a.Ho * a.Wo;
    SC_CHECK_ARG(M > 0 && M < (1ll << 31), SCOUTER_E_INVALID, "conv: M = %lld out of range", M);
    dim3 grid(cdiv((int)M, BM), cdiv(cout_g, BN), a.groups);
    conv_igemm_f32_kernel<<<grid, 256, 0, s>>>(a);
    SC_LAUNCH_CHECK();
    return 0;
}

int launch_stem_conv(const StemArgs& a, cudaStream_t s) {
    if (stem_ts_supported(a)) return launch_stem_ts(a, s);
    if (a.w_host && a.k == 3 && a.stride == 2 && a.pad == 1 && a.B <= 65535) {
        if (a.Cin == 3 && a.Cout == 32) return launch_stem_tiled<3, 32>(a, s);
        if (a.Cin == 1 && a.Cout == 64) return launch_stem_tiled<1, 64>(a, s);
    }
    SC_CHECK_ARG(a.Cin >= 1 && a.Cin <= 4, SCOUTER_E_UNSUPPORTED, "stem conv: Cin = %d (supports 1..4)", a.Cin);
    SC_CHECK_ARG(a.Cout % 4 == 0, SCOUTER_E_UNSUPPORTED, "stem conv: Cout = %d not a multiple of 4", a.Cout);
    size_t smem = (size_t)a.k * a.k * a.Cin * a.Cout * sizeof(float);
    SC_CHECK_ARG(smem <= 48 * 1024, SCOUTER_E_UNSUPPORTED, "stem conv: filter bank of %zu bytes exceeds 48 KB", smem);
    long long total = (long long)a.B * a.Ho * a.Wo * (a.Cout / 4);
    stem_conv_kernel<<<(unsigned)((total + 255) / 256), 256, smem, s>>>(a);
    SC_LAUNCH_CHECK();
    return 0;
}

int launch_nhwc_to_nchw(const float* in, float* out, int B, int HW, int C, cudaStream_t s) {
    dim3 grid(cdiv(C, 32), cdiv(HW, 32), B);
    transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(in, out, HW, C);
    SC_LAUNCH_CHECK();
    return 0;
}

int launch_nchw_to_nhwc(const float* in, float* out, int B, int C, int HW, cudaStream_t s) {
    dim3 grid(cdiv(HW, 32), cdiv(C, 32), B);
    transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(in, out, C, HW);
    SC_LAUNCH_CHECK();
    return 0;
}

}  // namespace scouter
