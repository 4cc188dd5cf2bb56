// Probe: per-SM throughput of the conv epilogue's output path, 148 CTAs, each writing `iters` tiles of 128 rows.
// Output tensor (M, C) fp32 row-major (row stride C*4 bytes), like an NHWC activation; a CTA tile = 128 rows x 128 columns.
//   mode 0: TMA store, box {16 cols, 128 rows} (64-byte rows, what emit_tma16 does), 8 stores per tile, 2 in flight
//   mode 1: TMA store, box {32 cols, 128 rows} (128-byte rows), 4 stores per tile, 2 in flight
//   mode 2: st.global.v4 row-per-thread (thread = row, 32 consecutive float4 = the direct-store epilogue)
//   mode 3: st.global.v4 coalesced (each warp instruction writes one 512-byte row segment)
//   mode 4: TMA LOAD, box {32 cols, 128 rows} into 2 buffers (residual prefetch path), 4 loads per tile
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../scouter_b200/csrc/ptx.cuh"
using namespace scouter::ptx;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(256) probe(const __grid_constant__ CUtensorMap tm16, const __grid_constant__ CUtensorMap tm32,
                                             float* out, int C, int m_tiles, int mode, long long* clk) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    fence_proxy_async();
    __syncthreads();
    const long long t0 = clock64();
    const int n_tiles = C / 128;
    const int total = m_tiles * n_tiles;
    if (mode <= 1) {
        if (threadIdx.x == 0) {
            int buf = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int nt = t % n_tiles, mt = t / n_tiles;
                const int per = mode == 0 ? 8 : 4;
                for (int j = 0; j < per; ++j) {
                    bulk_wait_read<1>();
                    if (mode == 0) tma_store_2d(&tm16, smem + buf * 8192, nt * 128 + j * 16, mt * 128);
                    else tma_store_2d(&tm32, smem + buf * 16384, nt * 128 + j * 32, mt * 128);
                    bulk_commit();
                    buf ^= 1;
                }
            }
            bulk_wait<0>();
        }
    } else if (mode == 2) {
        const int row = threadIdx.x & 127, half = threadIdx.x >> 7;   // 2 groups x 64 columns like the BN=128 epilogue
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const int nt = t % n_tiles, mt = t / n_tiles;
            float4* p = reinterpret_cast<float4*>(out + ((long long)mt * 128 + row) * C + nt * 128 + half * 64);
            const float4 v = make_float4(row, t, 1.f, 2.f);
#pragma unroll
            for (int j = 0; j < 16; ++j) p[j] = v;
        }
    } else if (mode == 3) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const int nt = t % n_tiles, mt = t / n_tiles;
            const float4 v = make_float4(lane, t, 1.f, 2.f);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int row = warp * 16 + j;
                reinterpret_cast<float4*>(out + ((long long)mt * 128 + row) * C + nt * 128)[lane] = v;
            }
        }
    } else {
        if (threadIdx.x == 0) {
            uint32_t ph[2] = {0, 0};
            int buf = 0, inflight = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int nt = t % n_tiles, mt = t / n_tiles;
                for (int j = 0; j < 4; ++j) {
                    if (inflight == 2) { mbar_wait(&bar[buf], ph[buf]); ph[buf] ^= 1; --inflight; }
                    mbar_arrive_expect_tx(&bar[buf], 16384);
                    tma_load_2d(smem + buf * 16384, &tm32, &bar[buf], nt * 128 + j * 32, mt * 128);
                    ++inflight;
                    buf ^= 1;
                }
            }
            while (inflight) { mbar_wait(&bar[buf], ph[buf]); ph[buf] ^= 1; --inflight; buf ^= 1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* f; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    Enc enc = (Enc)f;
    const int C = 256, M = 256 * 56 * 56;          // layer1 conv3 output: 822 MB
    float* out; cudaMalloc(&out, (size_t)M * C * 4);
    long long* clk; cudaMalloc(&clk, 148 * 8);
    CUtensorMap tm16, tm32;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M}, strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t es[2] = {1, 1};
    cuuint32_t b16[2] = {16, 128}, b32[2] = {32, 128};
    enc(&tm16, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, b16, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&tm32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, b32, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int smem = 65536 + 1024 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[5] = {"TMA store 64-B rows", "TMA store 128-B rows", "st.v4 row-per-thread", "st.v4 coalesced", "TMA load 128-B rows"};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            probe<<<148, 256, smem>>>(tm16, tm32, out, C, M / 128, mode, clk);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("%-24s %8.1f us  %7.0f GB/s  (%.1f clk per 128x128 tile per SM at 1.9 GHz)\n", names[mode], ms * 1e3,
                            (double)M * C * 4 / ms / 1e6, ms * 1e-3 * 1.9e9 / ((double)M / 128 * (C / 128) / 148));
        }
    }
    return 0;
}
