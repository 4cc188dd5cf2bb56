// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA + TMEM).
// Encodings follow the PTX ISA as used by CUTLASS's cute/arch/{mma_sm100_desc,mma_sm100_umma,copy_sm90_tma}.hpp.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace scouter {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// The hot loops keep 32-bit shared-window addresses of their barriers (one cvta at kernel start, not one per use).
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return mbar_try_wait_a(smem_u32(bar), parity); }
static __device__ __noinline__ void mbar_timeout(uint32_t parity) {
    printf("scouter_b200: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
    __trap();
}
// Bounded wait: a protocol bug must trap (and fail the launch), never hang the GPU.  The single-thread issue loops
// run at ~4.5 clk per instruction, so the fast path is one try_wait + branch and the diagnostics live out of line.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_a(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_a(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) mbar_timeout(parity);
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_a(smem_u32(bar), parity); }
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// ---- role-level stall accounting (only in the -DSCOUTER_PROF build made by scripts/prof_roles.py) -------
#ifdef SCOUTER_PROF
#define PROF_DECL(n) long long prof_##n = 0
#define PROF_T(n, stmt) do { const long long _t = clock64(); stmt; prof_##n += clock64() - _t; } while (0)
#define PROF_BEGIN(n) const long long prof_begin_##n = clock64()
#define PROF_END(n) prof_##n = clock64() - prof_begin_##n
#define PROF_STORE(buf, slot, n) (buf)[blockIdx.x * 32 + (slot)] = (unsigned long long)prof_##n
#else
#define PROF_DECL(n)
#define PROF_T(n, stmt) stmt
#define PROF_BEGIN(n)
#define PROF_END(n)
#define PROF_STORE(buf, slot, n)
#endif

// ---- TMA ------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// L2 prefetch of a box (no shared-memory destination, no completion tracking): decouples the DRAM latency from the
// depth of the shared-memory ring.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
                 : "memory");
}
// Multicast variant: the box lands at the same CTA-relative offset in every CTA of `mask`, and each of those CTAs'
// mbarrier (same CTA-relative offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}

// ---- thread-block clusters -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// Arrival without release semantics: "this CTA issues no more remote operations" (no memory of ours is read by the peer,
// so nothing has to be flushed -- the .release form stalls every warp on a membar).
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// TMA store (shared -> global), bulk async-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- tcgen05 / TMEM -------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: arrive (count 1) on an mbarrier when all previously issued MMAs of this thread are done.
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { umma_commit_a(smem_u32(bar)); }
// Same, arriving on the barrier at this CTA-relative offset in every CTA of `mask` (cluster multicast).
__device__ __forceinline__ void umma_commit_mc_a(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, tf32 operands, fp32 accumulate.  Issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands (K = 16 per instruction), fp32 accumulate.  Issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Same with the A operand in TENSOR MEMORY (lane = row, 32-bit columns along K: 8 columns per tf32 MMA, 8 columns of
// packed bf16x2 per bf16 MMA): no shared-memory bandwidth for A, floor M*N/256 clk instead of the SS-mode operand read.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Instruction descriptor, kind::f16 with bf16 operands, fp32 accumulator, both operands K-major.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4)                      // c_format  = F32
           | (1u << 7) | (1u << 10)       // a_format = b_format = BF16
           | ((uint32_t)(N >> 3) << 17)   // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// Instruction descriptor, kind::f16 with fp16 operands, fp32 accumulator, both operands K-major.  (The descriptor has one format
// field per operand, but a kind::f16 MMA whose A and B formats differ is an illegal instruction on sm_100a -- measured.)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // a_format = b_format = 0 (F16)
}

// Instruction descriptor, kind::tf32, fp32 accumulator, both operands K-major (mma_sm100_desc.hpp InstrDescriptor).
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4)                      // c_format  = F32
           | (2u << 7) | (2u << 10)       // a_format = b_format = TF32
           | ((uint32_t)(N >> 3) << 17)   // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// Shared-memory matrix descriptor for a K-major operand tile whose rows are 128 bytes (one SWIZZLE_128B
// atom along K): 8-row groups 1024 bytes apart (SBO), LBO unused (=1), version 1 (Blackwell), layout 2.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, 14 bits
    d |= (uint64_t)1 << 16;                    // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                    // descriptor version
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

// Same for 64-byte rows (32 bf16 along K, one SWIZZLE_64B atom): 8-row groups 512 bytes apart, layout 4.
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
    return d;
}

// Descriptors split into a constant high word and a low word {start address >> 4 | LBO}: the issue loops advance the
// low word with 32-bit adds (stage += bytes/16, k sub-step += 2, halo tap += rows*8) instead of rebuilding 64 bits.
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_HI_SW64 = (512u >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc_make(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// Operand split of one 8-float group for the error-compensated product: x = h + r with h = fp16(x) (round to nearest, 11-bit
// significand, saturating at +-65504) and r = x - h, which is EXACT in fp32 (the low 13 bits of x; below fp16's normal range
// h is a multiple of 2^-24 and r still exact) and ~2^-12 |x|.  The product is  a*w ~= a_h*w_h + a_b*w_r + a_r*w_h  with
//   ACTIVATIONS: a_h = fp16(a), a_b = bf16(a), a_r = fp16(a - a_h)   (three 16-bit forms), and
//   WEIGHTS:     w_h = fp16(w), w_r = bf16(w - w_h)                  (two forms),
// so the main product and one correction are fp16 x fp16 MMAs and the other correction is bf16 x bf16 (a kind::f16 MMA
// cannot mix the two formats).  The weight remainder needs bf16's exponent range (weights are ~1e-2, their remainders ~1e-6);
// the activation remainder in fp16 is exact to 2^-25 ABSOLUTE (fp16's subnormal spacing is 2^-24) and relative 2^-11 of
// itself above |a| = 0.25 -- for activations of scale >= 0.06 that is below the 2^-21 the bf16 operands leave anyway; a
// network whose activations are all below that degrades gracefully to an absolute error of 2^-25 |w| per product.
__device__ __forceinline__ void unpack_f16x2(uint32_t h2, float& lo, float& hi) {
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(h2));
}
__device__ __forceinline__ void split2_act(float lo, float hi, uint32_t& h2, uint32_t& b2, uint32_t& r2) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(hi), "f"(lo));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b2) : "f"(hi), "f"(lo));
    float fl, fh;
    unpack_f16x2(h2, fl, fh);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r2) : "f"(hi - fh), "f"(lo - fl));
}
__device__ __forceinline__ void split2_wgt(float lo, float hi, uint32_t& h2, uint32_t& r2) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(hi), "f"(lo));
    float fl, fh;
    unpack_f16x2(h2, fl, fh);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r2) : "f"(hi - fh), "f"(lo - fl));
}

// Converts rows of a SWIZZLE_128B fp32 tile (128-byte rows) into SWIZZLE_64B 16-bit tiles (64-byte rows) `tile` bytes apart:
// activations [fp16(x) | bf16(x) | fp16(x - fp16(x))], weights (WGT) [fp16(w) | bf16(w - fp16(w))].  128 threads; thread `tid`
// owns the 8-float group p = tid % 4 of rows tid/4 + 32*i, so its swizzled source / destination offsets are the same for every
// i up to a multiple of 4096 / 2048 bytes (row % 8 and (row / 2) % 4 do not change when the row advances by 32): three
// offsets per thread, computed once per kernel.
struct SplitLane {
    uint32_t src0, src1, dst;
    int row;
};
__device__ __forceinline__ SplitLane split_lane(int tid) {
    const int row = tid >> 2, p = tid & 3, s = row & 7;
    SplitLane l;
    l.src0 = row * 128 + (((2 * p) ^ s) << 4);
    l.src1 = row * 128 + (((2 * p + 1) ^ s) << 4);
    l.dst = row * 64 + ((p ^ ((row >> 1) & 3)) << 4);
    l.row = row;
    return l;
}
// U groups of 32 rows starting at row group `g0`; all loads are issued before the first conversion so that their
// shared-memory latency overlaps.  Rows >= rows are skipped (halo patches are not a multiple of 32 rows).
template <int U, bool GUARD, bool WGT = false>
__device__ __forceinline__ void split_groups(const uint8_t* src, uint8_t* dst, uint32_t tile, const SplitLane& l, int g0, int rows) {
    float4 a[U], b[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
        if (!GUARD || l.row + 32 * (g0 + i) < rows) {
            a[i] = *reinterpret_cast<const float4*>(src + l.src0 + 4096 * (g0 + i));
            b[i] = *reinterpret_cast<const float4*>(src + l.src1 + 4096 * (g0 + i));
        }
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
        if (!GUARD || l.row + 32 * (g0 + i) < rows) {
            const float v[8] = {a[i].x, a[i].y, a[i].z, a[i].w, b[i].x, b[i].y, b[i].z, b[i].w};
            uint32_t xh[4], xb[4], rb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if constexpr (WGT) split2_wgt(v[2 * j], v[2 * j + 1], xh[j], rb[j]);
                else split2_act(v[2 * j], v[2 * j + 1], xh[j], xb[j], rb[j]);
            }
            uint8_t* d = dst + l.dst + 2048 * (g0 + i);
            *reinterpret_cast<uint4*>(d) = make_uint4(xh[0], xh[1], xh[2], xh[3]);
            if constexpr (WGT) {
                *reinterpret_cast<uint4*>(d + tile) = make_uint4(rb[0], rb[1], rb[2], rb[3]);
            } else {
                *reinterpret_cast<uint4*>(d + tile) = make_uint4(xb[0], xb[1], xb[2], xb[3]);
                *reinterpret_cast<uint4*>(d + 2 * tile) = make_uint4(rb[0], rb[1], rb[2], rb[3]);
            }
        }
    }
}
// Whole tile of ROWS rows (a multiple of 32).
template <int ROWS, bool WGT = false>
__device__ __forceinline__ void split_tile_hbr(const uint8_t* src, uint8_t* dst, uint32_t tile, const SplitLane& l) {
    static_assert(ROWS % 32 == 0, "split_tile_hbr: whole 32-row groups");
    split_groups<ROWS / 32, false, WGT>(src, dst, tile, l, 0, ROWS);
}
// Any number of rows (halo patches).
__device__ __forceinline__ void split_rows_hbr(const uint8_t* src, uint8_t* dst, uint32_t tile, const SplitLane& l, int rows) {
    const int groups = (rows + 31) >> 5;
    int g = 0;
    for (; g + 4 <= groups; g += 4) split_groups<4, true>(src, dst, tile, l, g, rows);
    for (; g < groups; ++g) split_groups<1, true>(src, dst, tile, l, g, rows);
}

// 32 lanes x 32 columns of fp32 from TMEM: thread i of the warp gets lane (base_lane + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
// Registers -> TMEM: thread i of the warp writes lane (base_lane + i), columns [col, col+N).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ptx
}  // namespace scouter
