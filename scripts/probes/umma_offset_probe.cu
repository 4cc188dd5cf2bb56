// Probe: can a K-major SWIZZLE_128B UMMA operand start at an arbitrary 128-byte row inside a TMA-written patch?
// (needed to run the 9 taps of a 3x3 convolution out of ONE halo tile instead of 9 shifted TMA loads)
// For row offsets j = 0..20 and three choices of the descriptor's base_offset field, D = A[j:j+128,:] * B^T (K = 32)
// is compared with the host result.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../scouter_b200/csrc/ptx.cuh"
using namespace scouter::ptx;

constexpr int ROWS = 256, N = 32;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                             float* out, int joff, int mode) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                   // 256 rows x 128 B
    uint8_t* sB = smem + ROWS * 128;      // 32 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + N * 128);
    uint64_t* done = bar + 1;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(tptr, 32);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, ROWS * 128 + N * 128);
        tma_load_2d(sA, &tmA, bar, 0, 0);
        tma_load_2d(sA + 128 * 128, &tmA, bar, 0, 128);
        tma_load_2d(sB, &tmB, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA) + joff * 128;
        uint64_t da = smem_desc_sw128(a0);
        int bo = mode == 0 ? 0 : (mode == 1 ? (joff & 7) : ((8 - (joff & 7)) & 7));
        da |= (uint64_t)bo << 49;
        const uint64_t db = smem_desc_sw128(smem_u32(sB));
        const uint32_t idesc = idesc_tf32(128, N);
        for (int k = 0; k < 4; ++k) umma_tf32(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + j] = __uint_as_float(r[j]);
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 32);
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* f; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    Enc enc = (Enc)f;
    std::vector<float> A(ROWS * 32), B(N * 32);
    for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.f;   // tf32-exact values
    for (auto& v : B) v = (float)((rand() % 17) - 8) / 8.f;
    float *dA, *dB, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tA, tB;
    cuuint64_t dA2[2] = {32, ROWS}, sA2[1] = {128}; cuuint32_t bA[2] = {32, 128}, es[2] = {1, 1};
    cuuint64_t dB2[2] = {32, N}; cuuint32_t bB[2] = {32, N};
    enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dA2, sA2, bA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dB2, sA2, bB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int smem = ROWS * 128 + N * 128 + 1024 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> O(128 * N);
    for (int mode = 0; mode < 3; ++mode) {
        printf("mode %d (base_offset = %s):", mode, mode == 0 ? "0" : mode == 1 ? "j%8" : "(8-j%8)%8");
        for (int j = 0; j <= 20; ++j) {
            probe<<<1, 128, smem>>>(tA, tB, dO, j, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf(" j=%d CUDA error %s\n", j, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 32; ++k) ref += (double)A[(m + j) * 32 + k] * B[n * 32 + k];
                    maxerr = fmax(maxerr, fabs(ref - O[m * N + n]));
                }
            printf(" %d:%s", j, maxerr < 1e-4 ? "ok" : "BAD");
        }
        printf("\n");
    }
    return 0;
}
