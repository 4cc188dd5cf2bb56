// Probe: per-CTA latency and throughput of a TMA ring streaming a tall, thin slab (the fused head's feature stream):
// tensor (M, K=2048) fp32 row-major, box {32 channels, R rows}, NA-deep ring, one producer thread and one consumer thread
// (the consumer only waits for the data and hands the stage back).  Prints the average clocks per box and the
// issue-to-landed latency of a box in steady state, for grids of 16 and 128 CTAs and several ring depths.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_stream_probe tma_stream_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../scouter_b200/csrc/ptx.cuh"
using namespace scouter::ptx;

__global__ void __launch_bounds__(256) probe(const __grid_constant__ CUtensorMap tm, int R, int NA, int kblocks, int stage_bytes,
                                             int P, int blocked, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 16 * stage_bytes);
    uint64_t* empty = full + 16;
    __shared__ long long t_issue[64];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    const int warp = threadIdx.x / 32;
    if (warp >= 4 && warp < 4 + P && threadIdx.x % 32 == 0) {
        const int p = warp - 4;                       // producer p issues k-blocks kb = p (mod P)
        for (int kb = p; kb < kblocks; kb += P) {
            const int stage = kb % NA;
            if (kb >= NA) mbar_wait(&empty[stage], (uint32_t)((kb / NA) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[stage], (uint32_t)(R * 128));
            t_issue[kb] = clock64();
            if (blocked) tma_load_2d(smem + stage * stage_bytes, &tm, &full[stage], 0, (kb * gridDim.x + blockIdx.x) * R);
            else tma_load_2d(smem + stage * stage_bytes, &tm, &full[stage], kb * 32, blockIdx.x * R);
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        long long lat = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&full[stage], phase);
            if (kb >= kblocks / 2) lat += clock64() - t_issue[kb];
            mbar_arrive(&empty[stage]);
            if (++stage == NA) { stage = 0; phase ^= 1; }
        }
        out[blockIdx.x * 2] = clock64() - t0;
        out[blockIdx.x * 2 + 1] = lat / (kblocks - kblocks / 2);
    }
}

int main() {
    const int K = 2048, B = 256, n = 49, M = B * n, kblocks = 64;
    float* feat;
    cudaMalloc(&feat, (size_t)M * K * 4);
    cudaMemset(feat, 0, (size_t)M * K * 4);
    long long* out;
    cudaMalloc(&out, 256 * 2 * sizeof(long long));
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))f;
    char* flush;
    cudaMalloc(&flush, 256 << 20);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int blocked : {0, 1})
    for (int R : {98}) {
        CUtensorMap tm;
        // blocked = 1: the same bytes viewed as (32, M*64): every box is one contiguous 12.5 KB run
        cuuint64_t dims[2] = {(cuuint64_t)(blocked ? 32 : K), (cuuint64_t)(blocked ? (size_t)M * 64 : M)};
        cuuint64_t strides[1] = {(cuuint64_t)(blocked ? 32 : K) * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)R};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, feat, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int stage_bytes = ((R + 7) / 8 * 8) * 128;
        for (int grid : {16, 128}) {
            if (grid * R > M) continue;
            for (int P : {1, 2, 4})
            for (int NA : {4, 8, 12}) {
                cudaMemset(flush, 1, 256 << 20);
                probe<<<grid, 256, 16 * stage_bytes + 2048, 0>>>(tm, R, NA, kblocks, stage_bytes, P, blocked, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
                long long h[512];
                cudaMemcpy(h, out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
                double tot = 0, lat = 0;
                for (int i = 0; i < grid; ++i) { tot += h[2 * i]; lat += h[2 * i + 1]; }
                printf("blocked=%d P=%d R=%3d grid=%3d NA=%2d: %7.0f clk/box, steady-state issue->landed latency %6.0f clk, %.2f TB/s aggregate at 1.9 GHz\n", blocked, P, R,
                       grid, NA, tot / grid / kblocks, lat / grid, (double)grid * R * 128 * kblocks / (tot / grid / 1.9e9) / 1e12);
            }
        }
    }
    return 0;
}
