// Probe: cost of one "tap" of the conv issue loop on one SM, isolated from data movement: per iteration the thread does
//   [W] a try_wait on an already-completed mbarrier, [F] tcgen05.fence::after_thread_sync, [M] 12 MMAs, [C] c commits.
// Variants switch the pieces on/off to see what the ~1000 clk per tap of the small-N conv kernels are made of.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../../scouter_b200/csrc/ptx.cuh"
using namespace scouter::ptx;

template <int N>
__global__ void __launch_bounds__(128) tap(long long* out, int iters, int do_wait, int do_fence, int nmma, int ncommit) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);   // [0] ready (completed once), [1..2] commit sinks, [3] final
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x / 32;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); mbar_arrive(&bars[0]); }
    fence_proxy_async();
    if (warp == 1) tmem_alloc(tptr, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        const uint32_t sa = smem_u32(smem);
        const uint32_t idesc = idesc_tf32(128, N);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (do_wait) mbar_wait(&bars[0], 0);
            if (do_fence) tc_fence_after();
            const uint64_t da = smem_desc_sw128(sa + (it & 3) * 128);
            const uint64_t db = smem_desc_sw128(sa + 16384);
            for (int k = 0; k < nmma; ++k) umma_tf32(tmem + (k % 3) * N, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
            for (int c = 0; c < ncommit; ++c) umma_commit(&bars[1 + c]);   // phases flip freely: nobody waits on them
        }
        long long t1 = clock64();
        umma_commit(&bars[3]); mbar_wait(&bars[3], 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N>
void run(long long* d, const char* label, int w, int f, int m, int c) {
    int smem = 16384 + 32768 + 1024 + 128;
    cudaFuncSetAttribute(tap<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 200;
    tap<N><<<1, 128, smem>>>(d, iters, w, f, m, c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: error %s\n", label, cudaGetErrorString(e)); return; }
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%3d %-34s issue %7.1f clk/tap   retire %7.1f clk/tap\n", N, label, (double)h[0] / iters, (double)h[1] / iters);
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    run<32>(d, "12 mma", 0, 0, 12, 0);
    run<32>(d, "12 mma + 1 commit", 0, 0, 12, 1);
    run<32>(d, "12 mma + 2 commits", 0, 0, 12, 2);
    run<32>(d, "wait + fence + 12 mma + 2 commits", 1, 1, 12, 2);
    run<32>(d, "wait + fence only", 1, 1, 0, 0);
    run<32>(d, "2 commits only", 0, 0, 0, 2);
    run<32>(d, "8 mma + 2 commits", 0, 0, 8, 2);
    run<128>(d, "12 mma + 2 commits", 0, 0, 12, 2);
    run<128>(d, "wait + fence + 12 mma + 2 commits", 1, 1, 12, 2);
    return 0;
}
