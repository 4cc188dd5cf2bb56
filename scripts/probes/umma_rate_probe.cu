// Probe: cost of one tcgen05.mma.kind::tf32 (M=128, K=8, SS operands, SWIZZLE_128B) as a function of N, measured on one
// SM: one thread issues R MMAs back to back into the same accumulator, commits, waits; clock64 around the lot.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../../scouter_b200/csrc/ptx.cuh"
using namespace scouter::ptx;

template <int N>
__global__ void __launch_bounds__(128) rate(long long* out, int reps, int distinct) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* done = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x / 32;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(done, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 1) tmem_alloc(tptr, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        const uint64_t da = smem_desc_sw128(smem_u32(smem));
        const uint64_t db = smem_desc_sw128(smem_u32(smem + 16384));
        const uint32_t idesc = idesc_tf32(128, N);
        // warm-up
        for (int i = 0; i < 8; ++i) umma_tf32(tmem, da, db, idesc, i != 0);
        umma_commit(done); mbar_wait(done, 0);
        long long t0 = clock64();
        for (int i = 0; i < reps; ++i) umma_tf32(tmem + (distinct ? (i & 1) * N : 0), da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1);
        long long t1 = clock64();           // issue time only
        umma_commit(done); mbar_wait(done, 1);
        long long t2 = clock64();           // until all retired
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N>
void run(long long* d) {
    int smem = 16384 + 32768 + 1024 + 64;
    cudaFuncSetAttribute(rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int reps : {16, 64, 256}) {
        long long h[2];
        rate<N><<<1, 128, smem>>>(d, reps, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d error %s\n", N, cudaGetErrorString(e)); return; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N=%3d reps=%3d: issue %6.1f clk/MMA, retire %6.1f clk/MMA\n", N, reps, (double)h[0] / reps, (double)h[1] / reps);
    }
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    run<16>(d); run<32>(d); run<64>(d); run<128>(d); run<256>(d);
    return 0;
}
