// Probe: chip-wide sustained peaks of the two arithmetic units the rooflines of this repo are quoted against and that
// MEASURED_PEAKS.json does not hold: tcgen05.mma.kind::tf32 (M128 N256 K8, SS operands) and FP32 FFMA.
// Both run on every SM for tens of milliseconds (so the clocks settle under the power cap) and are timed with CUDA events.
// Prints one JSON object; scripts/gpu_peaks.sh stores it as gpurun_out/unit_peaks.json.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o unit_peaks_probe unit_peaks_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../../scouter_b200/csrc/ptx.cuh"
using namespace scouter::ptx;

template <bool BF16>
__global__ void __launch_bounds__(128) umma_sustained(int reps) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* done = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 2);
    const int warp = threadIdx.x / 32;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (threadIdx.x == 0) { mbar_init(&done[0], 1); mbar_init(&done[1], 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 1) tmem_alloc(tptr, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        const uint64_t da = smem_desc_sw128(smem_u32(smem));
        const uint64_t db = smem_desc_sw128(smem_u32(smem + 16384));
        const uint32_t idesc = BF16 ? idesc_bf16(128, 256) : idesc_tf32(128, 256);
        uint32_t ph[2] = {0, 0};
        for (int r = 0; r < reps; ++r) {          // 64 MMAs per commit, two accumulators in flight
            const int b = r & 1;
            if (r >= 2) { mbar_wait(&done[b], ph[b]); ph[b] ^= 1; }
            for (int i = 0; i < 64; ++i) {
                if (BF16) umma_bf16(tmem + b * 256, da + 2 * (i & 3), db + 2 * (i & 3), idesc, i != 0);
                else umma_tf32(tmem + b * 256, da + 2 * (i & 3), db + 2 * (i & 3), idesc, i != 0);
            }
            umma_commit(&done[b]);
        }
        for (int b = 0; b < 2; ++b) if (reps > b) mbar_wait(&done[b], ph[b]);   // the last commit on each accumulator
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(512) ffma_sustained(float* out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s;
}

template <class F>
static double best_ms(F&& launch, int reps = 3) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); exit(1); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, 64);
    const int smem = 16384 + 32768 + 1024 + 64;
    cudaFuncSetAttribute(umma_sustained<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(umma_sustained<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // burst: ~2 ms; sustained: ~60 ms back to back
    const int reps_burst = 400, reps_long = 12000;
    auto tf = [&](double ms, int reps, int k) { return 2.0 * 128 * 256 * k * 64.0 * reps * sms / (ms * 1e-3) / 1e12; };
    double ms;
    ms = best_ms([&] { umma_sustained<false><<<sms, 128, smem>>>(reps_burst); });
    const double tf32_burst = tf(ms, reps_burst, 8);
    ms = best_ms([&] { umma_sustained<false><<<sms, 128, smem>>>(reps_long); }, 2);
    const double tf32_sus = tf(ms, reps_long, 8);
    ms = best_ms([&] { umma_sustained<true><<<sms, 128, smem>>>(reps_burst); });
    const double bf16_burst = tf(ms, reps_burst, 16);
    ms = best_ms([&] { umma_sustained<true><<<sms, 128, smem>>>(reps_long); }, 2);
    const double bf16_sus = tf(ms, reps_long, 16);
    const int it_b = 2000, it_l = 60000;
    auto ff = [&](double ms, int iters) { return 2.0 * 64.0 * iters * 512.0 * 4 * sms / (ms * 1e-3) / 1e12; };
    ms = best_ms([&] { ffma_sustained<<<4 * sms, 512>>>(out, it_b, 1.0001f, 0.5f); });
    const double ffma_burst = ff(ms, it_b);
    ms = best_ms([&] { ffma_sustained<<<4 * sms, 512>>>(out, it_l, 1.0001f, 0.5f); }, 2);
    const double ffma_sus = ff(ms, it_l);
    printf("{\"sms\": %d, \"tf32_umma_tflops\": %.1f, \"tf32_umma_tflops_sustained\": %.1f, \"bf16_umma_tflops\": %.1f, "
           "\"bf16_umma_tflops_sustained\": %.1f, \"fp32_ffma_tflops\": %.2f, \"fp32_ffma_tflops_sustained\": %.2f, "
           "\"how\": \"scripts/probes/unit_peaks_probe.cu: tcgen05.mma M128 N256 (K8 tf32 / K16 bf16, SS operands) on every SM, "
           "64 MMAs per commit, 2 accumulators; FFMA: 4 CTAs x 512 threads per SM, 8 independent chains; burst ~2 ms, sustained ~60 ms; CUDA events\"}\n",
           sms, tf32_burst, tf32_sus, bf16_burst, bf16_sus, ffma_burst, ffma_sus);
    return 0;
}
