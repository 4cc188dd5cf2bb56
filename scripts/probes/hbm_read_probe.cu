// Probe: what read-only bandwidth can one kernel pull out of HBM on this part, by access path?  The fused head streams a
// 102.8 MB feature tensor exactly once; its roofline denominator (MEASURED_PEAKS.json) is a torch COPY (read + write).
//   (1) ldg:   plain LDG.128 streaming sum, grid = SMs x {1,2,4,8} CTAs of 256 threads, 8 independent loads per thread
//   (2) tma2d: the head's feature stream -- tensor (M, 2048) fp32, box {32 channels, R rows}, NA-deep ring, P producers,
//              one consumer thread handing the stage back -- on 128 / 148 CTAs, R = 98 / 85 / 64
//   (3) bulk:  cp.async.bulk (1-D, contiguous CHUNK bytes) ring on 148 CTAs
// Every run is preceded by a 256 MB write so the 102.8 MB come from DRAM, and timed with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o hbm_read_probe hbm_read_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../scouter_b200/csrc/ptx.cuh"
using namespace scouter::ptx;

__global__ void __launch_bounds__(256) ldg_read(const float4* __restrict__ p, size_t n4, float* out) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
    }
    for (; i < n4; i += stride) { const float4 v = __ldcs(p + i); a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    if (a.x + a.y + a.z + a.w == 12345.678f) out[0] = a.x;
}

// contiguous per-CTA slabs instead of grid-strided (what a persistent tile kernel would do)
__global__ void __launch_bounds__(256) ldg_read_slab(const float4* __restrict__ p, size_t n4, float* out) {
    const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
    const size_t b = (size_t)blockIdx.x * per, e = b + per < n4 ? b + per : n4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    size_t i = b + threadIdx.x;
    for (; i + 7 * 256 < e; i += 8 * 256) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + i + u * 256);
#pragma unroll
        for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
    }
    for (; i < e; i += 256) { const float4 v = __ldcs(p + i); a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    if (a.x + a.y + a.z + a.w == 12345.678f) out[0] = a.x;
}

__global__ void __launch_bounds__(256) tma2d(const __grid_constant__ CUtensorMap tm, int R, int NA, int kblocks, int stage_bytes,
                                             int P, int row_stride_units) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 16 * stage_bytes);
    uint64_t* empty = full + 16;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int warp = threadIdx.x / 32;
    if (warp >= 4 && warp < 4 + P && threadIdx.x % 32 == 0) {
        const int p = warp - 4;
        for (int kb = p; kb < kblocks; kb += P) {
            const int stage = kb % NA;
            if (kb >= NA) mbar_wait(&empty[stage], (uint32_t)((kb / NA) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[stage], (uint32_t)(R * 128));
            tma_load_2d(smem + stage * stage_bytes, &tm, &full[stage], kb * 32, blockIdx.x * row_stride_units);
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&full[stage], phase);
            mbar_arrive(&empty[stage]);
            if (++stage == NA) { stage = 0; phase ^= 1; }
        }
    }
}

// 1-D bulk copies of CHUNK contiguous bytes, NA-deep ring, chunks interleaved over the grid
__global__ void __launch_bounds__(128) bulk1d(const uint8_t* __restrict__ src, size_t bytes, int chunk, int NA) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)16 * chunk);
    uint64_t* empty = full + 16;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const size_t nchunks = bytes / chunk;
    if (threadIdx.x == 0) {
        int it = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
            const int stage = it % NA;
            if (it >= NA) mbar_wait(&empty[stage], (uint32_t)((it / NA) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[stage], (uint32_t)chunk);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(smem + (size_t)stage * chunk)),
                         "l"(src + c * chunk), "r"(chunk), "r"(smem_u32(&full[stage]))
                         : "memory");
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            mbar_wait(&full[stage], phase);
            mbar_arrive(&empty[stage]);
            if (++stage == NA) { stage = 0; phase ^= 1; }
        }
    }
}

static char* g_flush;
static char* g_flush2;
static float* g_out;
// how the 102.8 MB are evicted from the 126 MB L2 before each timed launch:
//   0 = 256 MB memset (L2 left full of DIRTY lines: their write-back competes with the timed reads)
//   1 = memset, then a 256 MB READ of another buffer (L2 left full of clean lines)
static int g_flush_mode = 0;
template <class F>
static double timed_us(F&& launch, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        cudaMemsetAsync(g_flush, r, 256 << 20);
        if (g_flush_mode == 1) ldg_read<<<1184, 256>>>((const float4*)g_flush2, (size_t)(256 << 20) / 16, g_out);
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(1); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms * 1e3 < best) best = ms * 1e3;
    }
    return best;
}

int main() {
    const int K = 2048, B = 256, n = 49, M = B * n, kblocks = 64;
    const size_t bytes = (size_t)M * K * 4;
    float* feat;
    cudaMalloc(&feat, bytes + (1 << 20));
    cudaMemset(feat, 0, bytes + (1 << 20));
    float* out;
    cudaMalloc(&out, 1024);
    cudaMalloc(&g_flush, 256 << 20);
    cudaMalloc(&g_flush2, 256 << 20);
    cudaMemset(g_flush2, 0, 256 << 20);
    g_out = out;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d, buffer %.1f MB\n", sms, bytes / 1e6);

  for (int fm : {0, 1}) {
    g_flush_mode = fm;
    printf("---- flush mode %d (%s)\n", fm, fm ? "memset + 256 MB read: clean L2" : "256 MB memset: dirty L2");
    for (int mult : {2, 4, 8}) {
        for (int g0 : {128, sms}) {
            const int grid = g0 * mult;
            double us = timed_us([&] { ldg_read<<<grid, 256>>>((const float4*)feat, bytes / 16, out); });
            printf("ldg  grid-strided grid=%5d x256: %7.1f us  %.2f TB/s\n", grid, us, bytes / us / 1e6);
            us = timed_us([&] { ldg_read_slab<<<grid, 256>>>((const float4*)feat, bytes / 16, out); });
            printf("ldg  per-CTA slab grid=%5d x256: %7.1f us  %.2f TB/s\n", grid, us, bytes / us / 1e6);
        }
    }

    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))f;
    cudaFuncSetAttribute(tma2d, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    for (int promo : {3})
    for (int R : {98, 85}) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)R};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, feat, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int stage_bytes = ((R + 7) / 8 * 8) * 128;
        const int grid = (M + R - 1) / R;       // every row exactly once: R=98 -> 128 CTAs, 85 -> 148, 64 -> 196
        for (int P : {1, 2})
        for (int NA : {8, 16}) {
            if ((size_t)16 * stage_bytes + 2048 > 225 * 1024) continue;
            double us = timed_us([&] { tma2d<<<grid, 256, 16 * stage_bytes + 2048>>>(tm, R, NA, kblocks, stage_bytes, P, R); });
            printf("tma2d promo=%d R=%3d grid=%3d P=%d NA=%2d: %7.1f us  %.2f TB/s\n", promo == 2 ? 128 : 256, R, grid, P, NA, us,
                   bytes / us / 1e6);
        }
    }
    cudaFuncSetAttribute(bulk1d, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    for (int chunk : {8192, 12288})
        for (int NA : {8, 16})
            for (int grid : {sms}) {
                if (grid > sms && (size_t)16 * chunk + 2048 > 110 * 1024) continue;
                double us = timed_us([&] { bulk1d<<<grid, 128, (size_t)16 * chunk + 2048>>>((const uint8_t*)feat, bytes, chunk, NA); });
                printf("bulk1d chunk=%5d grid=%3d NA=%2d: %7.1f us  %.2f TB/s\n", chunk, grid, NA, us, bytes / us / 1e6);
            }
  }
    return 0;
}
