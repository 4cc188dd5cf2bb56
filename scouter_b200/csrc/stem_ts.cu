// Deep-stem conv1.0 (timm/models/resnet.py:401: Conv2d(3, 32, 3, stride 2, pad 1) + BN + ReLU, NCHW fp32 in, NHWC fp32 out)
// on the tensor cores, im2col in TENSOR MEMORY.
//
// The CUDA-core stem does 864 FMAs per output pixel against constant-bank weights and runs at 22 % of the FMA rate
// (344 us at B = 256); its traffic (154 MB in, 411 MB out) is worth ~90 us.  Here a tile is 8 x 16 output pixels:
//   * warp 0 streams the 17 x 33 x 3 input patch of the tile with ONE 3-D TMA box per tile (out-of-bound zero fill = the
//     padding; the box is 40 columns wide so that its innermost start coordinate is 16-byte aligned) through a 4-deep ring;
//   * four splitter warps (thread = output pixel) gather the pixel's 3x3x3 window (27 values, K order (r, s, c) like the
//     OHWI weights, padded to 32) and tcgen05.st it as the A operand: 32 columns fp32 (kind::tf32 reads trunc19 exactly) +
//     16 columns bf16(x - trunc19(x));
//   * the weights (32 x 27 -> W, W_r = W - trunc19(W) as tf32 tiles, bf16 W) are resident K-major swizzled tiles built once
//     per CTA; the issuer runs 4 + 4 tf32 and 2 bf16 TS-mode MMAs (M128 x N32) per tile: A_t W_t + A_t W_r + A_r W;
//   * four epilogue warps add bias, ReLU and leave through swizzled staging + 4-D TMA stores.
// Same error-compensated product as every other tensor-core conv here (dropped term A_r W_r ~ 2^-22).
#include <cuda_bf16.h>

#include <cstdint>

#include "ptx.cuh"
#include "umma.cuh"

namespace scouter {
using namespace ptx;

namespace {

struct StemTsArgs {
    const float* w;      // (32, 3, 3, 3) fp32 OHWI, BN folded
    const float* bias;   // (32)
    int B, H, W, Ho, Wo;
    int tw, th, tiles;
    int relu;
};

constexpr int ST_THREADS = 384;
constexpr int ST_TH = 8, ST_TW = 16;                       // output tile: 128 pixels
// patch rows; patch row pitch in floats: the box starts at column 2*w0 - 4 (TMA wants a 16-byte aligned innermost
// coordinate: 2*w0 - 1 faults) and the window of the tile is columns 3..35 of it
constexpr int ST_PH = 2 * ST_TH + 1, ST_PWB = 40, ST_PX0 = 3;
constexpr int ST_PATCH = 3 * ST_PH * ST_PWB * 4;           // 8160 bytes per patch
constexpr int ST_PATCH_ALLOC = 8192;
constexpr int ST_PST = 4;
constexpr int ST_W32 = 32 * 128, ST_W16 = 32 * 64;         // weight tiles: fp32 SW128 / bf16 SW64, 32 rows (cout) x 32 K
constexpr int ST_OUT_STAGE = 128 * 64;
constexpr int ST_SMEM = 2 * ST_W32 + ST_W16 + ST_PST * ST_PATCH_ALLOC + 2 * ST_OUT_STAGE + 512 + 1024;
constexpr int ST_ACC = 0, ST_OP = 64;                      // TMEM: two 32-column accumulators, four 64-column operand buffers

__global__ void __launch_bounds__(ST_THREADS, 1)
stem_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmO, const StemTsArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* w32 = smem;                                   // W (tf32 operand)
    uint8_t* wr32 = w32 + ST_W32;                          // W - trunc19(W)
    uint8_t* wb16 = wr32 + ST_W32;                         // bf16(W)
    uint8_t* out_stage = wb16 + ST_W16;                    // [2][OUT_STAGE] (1024-aligned: 4096 + 4096 + 2048)
    uint8_t* patch0 = out_stage + 2 * ST_OUT_STAGE;        // PST patches
    uint64_t* bars = reinterpret_cast<uint64_t*>(patch0 + ST_PST * ST_PATCH_ALLOC);
    uint64_t* pfull = bars;            // [4] patch landed
    uint64_t* pempty = pfull + 4;      // [4] patch gathered (4 splitter warps)
    uint64_t* opfull = pempty + 4;     // [4] operands of the tile are in TMEM (4 splitter warps)
    uint64_t* opfree = opfull + 4;     // [4] the tile's MMAs retired: operand buffer reusable
    uint64_t* cfull = opfree + 4;      // [2] accumulator complete
    uint64_t* cempty = cfull + 2;      // [2] accumulator drained (128 epilogue threads)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(cempty + 2);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmO);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&pfull[i], 1);
            mbar_init(&pempty[i], 4);
            mbar_init(&opfull[i], 4);
            mbar_init(&opfree[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&cfull[i], 1); mbar_init(&cempty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    // resident weights: element (cout, k), k = (r*3 + s)*3 + c < 27, zero above
    for (int idx = threadIdx.x; idx < 32 * 32; idx += ST_THREADS) {
        const int co = idx >> 5, k = idx & 31;
        const float v = k < 27 ? __ldg(p.w + co * 27 + k) : 0.f;
        const float r = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        const int o32 = co * 128 + (((k >> 2) ^ (co & 7)) << 4) + ((k & 3) << 2);
        *reinterpret_cast<float*>(w32 + o32) = v;
        *reinterpret_cast<float*>(wr32 + o32) = r;
        const int o16 = co * 64 + (((k >> 3) ^ ((co >> 1) & 3)) << 4) + ((k & 7) << 1);
        *reinterpret_cast<__nv_bfloat16*>(wb16 + o16) = __float2bfloat16_rn(v);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto tile_coords = [&](int t, int& w0, int& h0, int& b) {
        w0 = (t % p.tw) * ST_TW;
        h0 = ((t / p.tw) % p.th) * ST_TH;
        b = t / (p.tw * p.th);
    };

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer: one {40, 17, 3} box of the NCHW input per tile =====
            int n = 0;
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++n) {
                const int ps = n & 3;
                int w0, h0, b;
                tile_coords(t, w0, h0, b);
                if (n >= ST_PST) mbar_wait(&pempty[ps], (uint32_t)((n >> 2) - 1) & 1u);
                mbar_arrive_expect_tx(&pfull[ps], (uint32_t)ST_PATCH);
                tma_load_3d(patch0 + ps * ST_PATCH_ALLOC, &tmA, &pfull[ps], 2 * w0 - 1 - ST_PX0, 2 * h0 - 1, b * 3);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = idesc_tf32(128, 32), idesc_b = idesc_bf16(128, 32);
            const uint32_t wt = desc_lo(smem_u32(w32)), wr = desc_lo(smem_u32(wr32)), wb = desc_lo(smem_u32(wb16));
            uint32_t n = 0;
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++n) {
                const uint32_t ob = n & 3, ab = n & 1;
                mbar_wait(&opfull[ob], (n >> 2) & 1u);
                mbar_wait(&cempty[ab], ((n >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + ST_ACC + 32 * ab, a32 = tmem_base + ST_OP + 64 * ob, ar16 = a32 + 32;
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)   // A_t * W_t
                    umma_tf32_ts(d, a32 + 8 * j, desc_make(DESC_HI_SW128, wt + 2 * j), idesc, j != 0);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)   // A_t * W_r
                    umma_tf32_ts(d, a32 + 8 * j, desc_make(DESC_HI_SW128, wr + 2 * j), idesc, 1);
#pragma unroll
                for (uint32_t j = 0; j < 2; ++j)   // A_r * W (bf16)
                    umma_bf16_ts(d, ar16 + 8 * j, desc_make(DESC_HI_SW64, wb + 2 * j), idesc_b, 1);
                umma_commit(&opfree[ob]);
                umma_commit(&cfull[ab]);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== epilogue: thread = output pixel (compact index ph*16 + pw), 32 channels =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        float bias[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 v = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            bias[4 * j] = v.x; bias[4 * j + 1] = v.y; bias[4 * j + 2] = v.z; bias[4 * j + 3] = v.w;
        }
        uint32_t n = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++n) {
            const uint32_t ab = n & 1;
            int w0, h0, b;
            tile_coords(t, w0, h0, b);
            mbar_wait(&cfull[ab], (n >> 1) & 1u);
            tc_fence_after();
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + ST_ACC + 32 * ab, r);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&cempty[ab]);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                v[j] = __uint_as_float(r[j]) + bias[j];
                if (p.relu) v[j] = fmaxf(v[j], 0.f);
            }
#pragma unroll
            for (int c16 = 0; c16 < 2; ++c16) {   // 16 channels of every tile pixel per bulk tensor store
                uint8_t* stg = out_stage + c16 * ST_OUT_STAGE;
                if (row == 0) bulk_wait_read<1>();
                named_bar_sync(1, 128);
#pragma unroll
                for (int j = 0; j < 4; ++j)   // SWIZZLE_64B: 16-byte chunk index ^= (row / 2) % 4
                    *reinterpret_cast<float4*>(stg + row * 64 + ((j ^ ((row >> 1) & 3)) << 4)) =
                        make_float4(v[c16 * 16 + 4 * j], v[c16 * 16 + 4 * j + 1], v[c16 * 16 + 4 * j + 2], v[c16 * 16 + 4 * j + 3]);
                fence_proxy_async();
                named_bar_sync(1, 128);
                if (row == 0) {
                    tma_store_4d(&tmO, stg, c16 * 16, w0, h0, b);
                    bulk_commit();
                }
            }
        }
        if (row == 0) bulk_wait<0>();
    } else if (warp >= 8 && warp < 12) {
        // ===== splitters: thread = output pixel; its 27-value window -> TMEM (fp32 and bf16 remainder) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int ph = row >> 4, pw = row & 15;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + ST_OP;
        uint32_t n = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++n) {
            const uint32_t ps = n & 3, ob = n & 3;
            mbar_wait(&pfull[ps], (n >> 2) & 1u);
            const float* patch = reinterpret_cast<const float*>(patch0 + ps * ST_PATCH_ALLOC) + (2 * ph) * ST_PWB + 2 * pw + ST_PX0;
            uint32_t f[32], rb[16];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                float x = 0.f;
                if (k < 27) {
                    const int c = k % 3, s_ = (k / 3) % 3, r_ = k / 9;
                    x = patch[(c * ST_PH + r_) * ST_PWB + s_];
                }
                f[k] = __float_as_uint(x);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pempty[ps]);      // the patch has been gathered
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float r0 = __uint_as_float(f[2 * i]) - __uint_as_float(f[2 * i] & 0xFFFFE000u);
                const float r1 = __uint_as_float(f[2 * i + 1]) - __uint_as_float(f[2 * i + 1] & 0xFFFFE000u);
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(rb[i]) : "f"(r1), "f"(r0));
            }
            if (n >= 4) {
                mbar_wait(&opfree[ob], ((n >> 2) - 1) & 1u);   // the MMAs that read this operand buffer have retired
                tc_fence_after();
            }
            tmem_st_32x32(t_lane + 64 * ob, f);
            tmem_st_32x16(t_lane + 64 * ob + 32, rb);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&opfull[ob]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn stem_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

}  // namespace

bool stem_ts_supported(const StemArgs& a) {
    static bool off = getenv("SCOUTER_NO_STEM_TS") != nullptr;
    if (off || !a.tc || a.Cin != 3 || a.Cout != 32 || a.k != 3 || a.stride != 2 || a.pad != 1 || a.round_out) return false;
    if ((a.W * 4) % 16 || ((uintptr_t)a.in & 15) || a.B * 3 > (1 << 30)) return false;
    return stem_encode_fn() != nullptr;
}

int launch_stem_ts(const StemArgs& a, cudaStream_t s) {
    EncodeTiledFn enc = stem_encode_fn();
    SC_CHECK_ARG(enc && stem_ts_supported(a), SCOUTER_E_UNSUPPORTED, "stem_ts: unsupported geometry");
    StemTsArgs u;
    u.w = a.w; u.bias = a.bias;
    u.B = a.B; u.H = a.H; u.W = a.W; u.Ho = a.Ho; u.Wo = a.Wo;
    u.tw = cdiv(a.Wo, ST_TW); u.th = cdiv(a.Ho, ST_TH);
    u.tiles = u.tw * u.th * a.B;
    u.relu = a.relu;
    CUtensorMap tmA, tmO;
    {
        // input NCHW viewed as (W, H, B*3) planes; box {40, 17, 3}: the 17 x 33 window (+ alignment columns) of the three planes
        cuuint64_t dims[3] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B * 3};
        cuuint64_t strides[2] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.H * a.W * 4};
        cuuint32_t box[3] = {(cuuint32_t)ST_PWB, (cuuint32_t)ST_PH, 3};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "stem_ts: cuTensorMapEncodeTiled(input) failed with %d", (int)r);
        cuuint64_t dimsO[4] = {32, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)a.B};
        cuuint64_t stridesO[3] = {32 * 4, (cuuint64_t)a.Wo * 32 * 4, (cuuint64_t)a.Ho * a.Wo * 32 * 4};
        cuuint32_t boxO[4] = {16, (cuuint32_t)ST_TW, (cuuint32_t)ST_TH, 1};
        cuuint32_t esO[4] = {1, 1, 1, 1};
        r = enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a.out, dimsO, stridesO, boxO, esO, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "stem_ts: cuTensorMapEncodeTiled(output) failed with %d", (int)r);
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        SC_CUDA(cudaGetDevice(&dev));
        SC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    SC_CUDA(cudaFuncSetAttribute(stem_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    stem_ts_kernel<<<std::min(u.tiles, sms), ST_THREADS, ST_SMEM, s>>>(tmA, tmO, u);
    SC_LAUNCH_CHECK();
    return 0;
}

}  // namespace scouter
