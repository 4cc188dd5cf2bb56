// The xSlot head as ONE kernel (SURVEY 8(d) "fused head"): conv1x1 + bias + ReLU (sloter/slot_model.py:108-109), + PE
// (:110-111) and SlotAttention.forward (sloter/utils/slot_attention.py:44-96) without the projected tokens ever
// leaving the SM.
//
//   unit  = G consecutive images whose R = G*n token rows fit one 128-row UMMA tile (n = 49 -> 2 images, 81 -> 1);
//           one CTA per unit, 512 threads, CTAs paired into clusters of two.
//   phase A (HBM-bound): the unit's (R x ch) fp32 feature slab streams through a deep TMA ring (NA stages of 32
//           channels).  Error-compensated tensor-core product as in umma_conv.cu (tf32 main product A_t*W_t + two bf16
//           correction products A*W_r, A_r*W; chunked TMEM accumulation merged in fp32 registers) -- but with the A
//           operands in TENSOR MEMORY: this GEMM is thin (N = 64), so in SS mode the MMAs' own operand reads (6 KB per
//           instruction, 48 KB per k-block) plus the split tiles saturate the 128 B/clk shared-memory pipe at 2.5x the
//           HBM time (measured: 1170 clk per k-block).  Here the splitter threads (thread = token row) read their row of
//           the TMA tile once, derive bf16(x) and bf16(x - trunc19(x)) in registers and tcgen05.st all three forms into
//           a TMEM operand buffer; the MMAs then read only the small weight tile from shared memory.
//           conv1x1.weight (fp32 + pre-split bf16 [W ; W_r]) streams through a 3-slot ring, multicast to both CTAs of
//           the cluster (each loads half).
//           warp 0 = feature producer, warp 3 = weight producer, warps 1 and 2 = MMA issuers on alternate k-blocks (four
//           TMEM accumulators: the single-thread issue cost is the floor of such a thin GEMM), warps 4-7 = accumulator
//           owners (thread = token row), warps 8-15 = two groups of splitters on alternate k-blocks.
//           ONE multicast tcgen05.commit per k-block releases the weight slot and the operand buffer in both CTAs.
//           The to_k weights are prefetched with cp.async meanwhile.
//   hand-over: the accumulator owners add the bias, apply ReLU and write X and X + PE straight into shared memory.
//   phase B (FP32 FMA, all 16 warps): to_k MLP, 3 x {QK^T, sum-normalise, sigmoid, attn.X / d, GRU}, logits -- the
//           arithmetic of xslot_fast.cu (fp32 with IEEE division, fixed-order reductions: bit-reproducible).
//
// Algorithmic HBM bytes = B*n*ch*4 (features) + weights + outputs; nothing is written back in between.
#include <cuda_bf16.h>

#include <cstdlib>

#include "ptx.cuh"
#include "umma.cuh"
#include "xslot.cuh"

namespace scouter {
using namespace ptx;

namespace {

constexpr int HT = 512, HW_ = HT / 32;
constexpr int LDX = XD + 4;
constexpr int W_FLOATS = 2 * XD * XG + 2 * XG;   // GRU block: WihT[64][192], WhhT[64][192], b_ih[192], b_hh[192]

// phase-A shared memory: [NA feature stages | NW weight slots]; tensor memory: [4 accumulators | NO operand buffers]
constexpr int W_F32 = XD * 128;                  // 64 rows x 32 tf32
constexpr int W_SLOT = 2 * W_F32;                // [W fp32 | bf16 W | bf16 W_r]
constexpr int MAX_NW = 8;                        // weight slots (a.nw of them)
constexpr int NO = 4;                            // TMEM operand buffers of 64 columns: [fp32 A (32) | bf16 A (16) | bf16 A_r (16)]
constexpr int ND = 8;                            // "k-block retired" barriers (>= MAX_NW, NO); power of two (index = kb & 7)
constexpr int OP_COL0 = 4 * XD;                  // operand buffers start after the four 64-column accumulators
// Loop phase, tensor-core GRU gates: weight tiles t = matrix*2 + gate-half as [fp32 (64) | bf16 remainder (32)] and the four
// 32-column accumulators.  Tiles 0-2 sit in columns the to_k MLP does not touch (it uses 0-63 and 256-383), so they can be
// filled while the MLP runs; tile 3 and the accumulators take the MLP's columns afterwards.
__device__ __forceinline__ uint32_t gw_col(int t) { return t == 0 ? 64u : (t == 1 ? 160u : (t == 2 ? 384u : 256u)); }
__device__ __forceinline__ uint32_t gd_col(int t) { return t == 0 ? 0u : (t == 1 ? 32u : (t == 2 ? 352u : 480u)); }
constexpr int MAX_NA = 8;
constexpr int CHUNK = 4;                         // k-blocks per TMEM accumulation chunk (per issuer; see umma_conv.cu)

struct FusedArgs {
    const float* conv_b;
    const float* gru_w_ih;   // (192, 64) row-major, PyTorch layout: K-major rows for the tensor-core gate GEMM
    const float* gru_w_hh;
    const float* gru_b_ih;
    const float* gru_b_hh;
    const float* packed;
    const float* pe;
    float* x_out;
    float* logits;
    float* attn;
    float* attn_sum;
    int B, n, G, S, C, spc, L, iters, loss_status;
    int kblocks;
    // shared-memory map (bytes from the 1024-aligned base); the phase-B regions alias the phase-A rings
    int na, nw, a_stage, off_w;
    int off_gru, off_small, off_bar;
    int tc_gates;      // GRU gate GEMMs on the tensor cores (W_ih / W_hh resident in TMEM); needs G*S <= 32
    int gw_early;      // byte offset of a 52 KB staging area inside the dead feature ring: the idle splitter warps move
                       // W_ih / W_hh into TMEM while the to_k MLP runs (0 = no room: staged inside the first iteration)
};

__host__ __device__ inline int kb_floats(int img, int n, int S) {
    const int a = img * n * LDX, b = img * S * 2 * XG + img * n * ((S + 3) & ~3);
    return ((a > b ? a : b) + 3) & ~3;
}

__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// updates: partial sum over the tokens j = j0, j0+jstep, ... of attn[j][i] * X[j][e..e+3] for C4*4 consecutive slots
template <int C4>
__device__ __forceinline__ void update_partial(const float* __restrict__ xp /* X[img][0][e] */, const float* __restrict__ ap /* attnT[img][0][i0] */,
                                               int n, int SP, int j0, int jstep, float* __restrict__ out /* upart[..][i0][e] */, int nslots) {
    float4 acc[4 * C4];
#pragma unroll
    for (int i = 0; i < 4 * C4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = j0; j < n; j += jstep) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + j * LDX);
#pragma unroll
        for (int c = 0; c < C4; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(ap + j * SP + 4 * c);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float4& o = acc[4 * c + t];
                o.x = fmaf(av[t], xv.x, o.x); o.y = fmaf(av[t], xv.y, o.y); o.z = fmaf(av[t], xv.z, o.z); o.w = fmaf(av[t], xv.w, o.w);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4 * C4; ++i)
        if (i < nslots) *reinterpret_cast<float4*>(out + i * XD) = acc[i];
}

#ifdef SCOUTER_PROF
__device__ unsigned long long g_prof_head[256 * 32];
__device__ unsigned long long g_trace_head[128 * 8];   // CTA 0: [k-block][event] clock64 timestamps
#define TRACE(kb, ev) do { if (blockIdx.x == 0 && (kb) < 128) g_trace_head[(kb) * 8 + (ev)] = clock64(); } while (0)
#else
#define TRACE(kb, ev)
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HT, 1)
head_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmT,
                  const __grid_constant__ CUtensorMap tmT2, const FusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int n = a.n, S = a.S, G = a.G;
    const int R = G * n;            // token rows of this unit
    const int SR = G * S;           // slot rows
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int b0 = blockIdx.x * G;
    const int nimg = min(G, a.B - b0);   // <= 0 for the padding CTA of an odd unit count (it streams zeros, writes nothing)
    const int NA = a.na, NW = a.nw;
    const uint32_t rank = cluster_ctarank();
    const XSlotPacked pk{S, a.L};

    uint64_t* fullA = reinterpret_cast<uint64_t*>(smem + a.off_bar);   // [MAX_NA] feature stage landed
    uint64_t* emptyA = fullA + MAX_NA;      // [MAX_NA] stage read into registers by its splitter group (4 warps)
    uint64_t* done = emptyA + MAX_NA;       // [ND] MMAs of the k-block retired in BOTH CTAs (2 multicast commits)
    uint64_t* fullW = done + ND;            // [NW] weight slot landed (own half + the peer's multicast half)
    uint64_t* opfull = fullW + MAX_NW;      // [NO] operands of the k-block are in TMEM
    uint64_t* cfull = opfull + NO;          // [2 issuers][2] accumulator chunk complete
    uint64_t* cempty = cfull + 4;           // [2][2] chunk drained by the 128 accumulator owners
    uint64_t* mlp_in = cempty + 4;          // to_k layer input is in TMEM (128 accumulator owners)
    uint64_t* mlp_out = mlp_in + 1;         // to_k layer MMAs retired
    uint64_t* gbar = mlp_out + 1;           // gate GEMMs of a GRU step retired (4 issuing threads)
    uint64_t* acc_done = gbar + 1;          // every conv accumulator has been drained (128 accumulator owners)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_done + 1);
    float* tokb = reinterpret_cast<float*>(smem + a.off_bar + 512);   // [L][64] to_k biases (read by every accumulator owner)
    float* slots0 = tokb + a.L * XD;                                   // [S][64] initial slots, staged during phase A

    // phase-B views
    const int SP = (S + 3) & ~3;
    float* Xs = reinterpret_cast<float*>(smem);   // [R][LDX]
    float* Ka = Xs + R * LDX;                      // [R][LDX]
    float* Kb = Ka + R * LDX;                      // MLP ping-pong; afterwards gates / attention scratch
    float* Wsm = reinterpret_cast<float*>(smem + a.off_gru);     // GRU block (aliases rings / to_k weights; loaded after the MLP)
    float* slots = reinterpret_cast<float*>(smem + a.off_small); // [SR][64]
    float* upd = slots + SR * XD;
    float* rsum = upd + SR * XD;
    float* usum = rsum + SR;
    float* gates = Kb;
    float* attnT = Kb + SR * 2 * XG;

    if (warp == 0 && elect_one()) {
        // the feature stream starts before anything else is set up: its first boxes need ~1.5k clk to arrive
        for (int i = 0; i < MAX_NA; ++i) {
            mbar_init(&fullA[i], 1);
            mbar_init(&emptyA[i], 4);
        }
        fence_barrier_init();
        for (int kb = 0; kb < min(NA, a.kblocks); ++kb) {
            mbar_arrive_expect_tx(&fullA[kb], (uint32_t)(R * 128));
            tma_load_2d(smem + kb * a.a_stage, &tmA, &fullA[kb], kb * 32, blockIdx.x * R);
        }
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmB2);
        prefetch_tmap(&tmT);
        prefetch_tmap(&tmT2);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < ND; ++i) mbar_init(&done[i], 2);
        for (int i = 0; i < MAX_NW; ++i) mbar_init(&fullW[i], 1);
        for (int i = 0; i < NO; ++i) mbar_init(&opfull[i], 4);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&cfull[i], 1);
            mbar_init(&cempty[i], 128);
        }
        mbar_init(mlp_in, 128);
        mbar_init(mlp_out, 1);
        mbar_init(gbar, 4);
        mbar_init(acc_done, 128);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_arrive();     // the peer multicasts into this CTA's weight slots and arrives on its barriers
    cluster_wait();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    PROF_DECL(phaseA); PROF_DECL(mlp); PROF_DECL(loop); PROF_BEGIN(phaseA);

    const int kblocks = a.kblocks;
    uint8_t* const w_ring = smem + a.off_w;

    // =========================================== phase A ===========================================================
    if (warp == 0) {
        if (elect_one()) {
            // ----- feature producer: runs NA k-blocks ahead of the MMAs (the first NA boxes went out during set-up) -----
            int stage = 0;
            uint32_t phase = 1;
            const int row0 = blockIdx.x * R;
            PROF_DECL(pa_wait);
            for (int kb = NA; kb < kblocks; ++kb) {
                PROF_T(pa_wait, mbar_wait(&emptyA[stage], phase ^ 1));
                mbar_arrive_expect_tx(&fullA[stage], (uint32_t)(R * 128));
                tma_load_2d(smem + stage * a.a_stage, &tmA, &fullA[stage], kb * 32, row0);
                TRACE(kb, 0);
                if (++stage == NA) { stage = 0; phase ^= 1; }
            }
            PROF_STORE(g_prof_head, 3, pa_wait);
        }
    } else if (warp == 3) {
        if (elect_one()) {
            // ----- weight producer: rank 0 multicasts the fp32 tile, rank 1 the bf16 [W ; W_r] pair -----
            PROF_DECL(pw_wait);
            int ws = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                if (kb >= NW) {
                    const int j = kb - NW;
                    PROF_T(pw_wait, mbar_wait(&done[j & (ND - 1)], (uint32_t)(j >> 3) & 1u));   // slot drained in both CTAs
                }
                mbar_arrive_expect_tx(&fullW[ws], (uint32_t)W_SLOT);
                if (rank == 0) tma_load_2d_mc(w_ring + ws * W_SLOT, &tmB, &fullW[ws], kb * 32, 0, (uint16_t)3);
                else tma_load_2d_mc(w_ring + ws * W_SLOT + W_F32, &tmB2, &fullW[ws], kb * 32, 0, (uint16_t)3);
                TRACE(kb, 1);
                if (++ws == NW) ws = 0;
            }
            // the to_k layers ride the same weight stream: layer l = two more "k-blocks" (its two 32-channel halves)
            for (int v = 0; v < 2 * a.L; ++v) {
                const int kb = kblocks + v, j = kb - NW;
                if (j >= 0) mbar_wait(&done[j & (ND - 1)], (uint32_t)(j >> 3) & 1u);
                mbar_arrive_expect_tx(&fullW[ws], (uint32_t)W_SLOT);
                if (rank == 0) tma_load_2d_mc(w_ring + ws * W_SLOT, &tmT, &fullW[ws], (v & 1) * 32, (v >> 1) * XD, (uint16_t)3);
                else tma_load_2d_mc(w_ring + ws * W_SLOT + W_F32, &tmT2, &fullW[ws], (v & 1) * 32, (v >> 1) * 2 * XD, (uint16_t)3);
                if (++ws == NW) ws = 0;
            }
            PROF_STORE(g_prof_head, 4, pw_wait);
        }
    } else if (warp == 1 || warp == 2) {
        if (elect_one()) {
            // ----- MMA issuers: issuer w takes k-blocks kb = w (mod 2) into its own pair of TMEM accumulators -----
            const int w = warp - 1;
            constexpr uint32_t idesc = idesc_tf32(128, XD);
            constexpr uint32_t idesc_b = idesc_bf16(128, XD);
            const uint32_t w_lo0 = desc_lo(smem_u32(w_ring));
            const uint32_t ready_a = smem_u32(opfull), done_a = smem_u32(done), cfull_a = smem_u32(cfull + 2 * w),
                           cempty_a = smem_u32(cempty + 2 * w), fullw_a = smem_u32(fullW);
            int ws = w % NW, ob = w, dn = w, in_chunk = 0;     // kb % NW, kb % NO, kb % ND (kb advances by 2)
            uint32_t ophase = 0, wphase = 0, cc = 0;
            PROF_DECL(is_cempty); PROF_DECL(is_op);
            for (int kb = w; kb < kblocks; kb += 2) {
                const uint32_t buf = cc & 1;
                if (in_chunk == 0) PROF_T(is_cempty, mbar_wait_a(cempty_a + 8 * buf, ((cc >> 1) & 1) ^ 1));
                PROF_T(is_op, mbar_wait_a(ready_a + 8 * ob, ophase));
                TRACE(kb, 5);
                mbar_wait_a(fullw_a + 8 * ws, wphase);
                TRACE(kb, 6);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(2 * w + buf) * XD;
                const uint32_t acc = in_chunk != 0;
                const uint32_t a_tm = tmem_base + OP_COL0 + (uint32_t)ob * 64;   // [fp32 A | bf16 A | bf16 A_r]
                const uint32_t w_lo = w_lo0 + (uint32_t)ws * (W_SLOT >> 4);      // [W fp32 | bf16 W | bf16 W_r]
#pragma unroll
                for (uint32_t k = 0; k < 2; ++k)   // A * W_r
                    umma_bf16_ts(d_tmem, a_tm + 32 + 8 * k, desc_make(DESC_HI_SW64, w_lo + ((W_F32 + W_F32 / 2) >> 4) + 2 * k), idesc_b, acc | k);
#pragma unroll
                for (uint32_t k = 0; k < 2; ++k)   // A_r * W
                    umma_bf16_ts(d_tmem, a_tm + 48 + 8 * k, desc_make(DESC_HI_SW64, w_lo + (W_F32 >> 4) + 2 * k), idesc_b, 1);
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)   // A_t * W_t
                    umma_tf32_ts(d_tmem, a_tm + 8 * k, desc_make(DESC_HI_SW128, w_lo + 2 * k), idesc, 1);
                umma_commit_mc_a(done_a + 8 * dn, (uint16_t)3);   // frees the weight slot and the operand buffer in both CTAs
                TRACE(kb, 7);
                dn = (dn + 2) & (ND - 1);
                ws += 2; if (ws >= NW) { ws -= NW; wphase ^= 1; }
                ob += 2; if (ob >= NO) { ob -= NO; ophase ^= 1; }
                if (++in_chunk == CHUNK || kb + 2 >= kblocks) {
                    umma_commit_a(cfull_a + 8 * buf);
                    ++cc;
                    in_chunk = 0;
                }
            }
            if (w == 0) { PROF_STORE(g_prof_head, 5, is_cempty); PROF_STORE(g_prof_head, 6, is_op); }
            if (w == 0) {
                // ----- to_k MLP: D[token, out] = A[token, in] * W_l^T, A = the previous layer's activations, re-split and
                //       stored to TMEM by the accumulator owners; same error-compensated product, 16 MMAs per layer -----
                const uint32_t a_tm = tmem_base + OP_COL0;     // [fp32 (64) | bf16 (32) | bf16 remainder (32)]
                for (int l = 0; l < a.L; ++l) {
                    mbar_wait(mlp_in, (uint32_t)l & 1u);
                    tc_fence_after();
                    for (int kk = 0; kk < 2; ++kk) {
                        const int kb = kblocks + 2 * l + kk;
                        const int ws2 = kb % NW;
                        mbar_wait_a(fullw_a + 8 * ws2, (uint32_t)(kb / NW) & 1u);
                        tc_fence_after();
                        const uint32_t w_lo = w_lo0 + (uint32_t)ws2 * (W_SLOT >> 4);
#pragma unroll
                        for (uint32_t k = 0; k < 2; ++k)   // A * W_r
                            umma_bf16_ts(tmem_base, a_tm + 64 + 16 * kk + 8 * k,
                                         desc_make(DESC_HI_SW64, w_lo + ((W_F32 + W_F32 / 2) >> 4) + 2 * k), idesc_b, (uint32_t)kk | k);
#pragma unroll
                        for (uint32_t k = 0; k < 2; ++k)   // A_r * W
                            umma_bf16_ts(tmem_base, a_tm + 96 + 16 * kk + 8 * k, desc_make(DESC_HI_SW64, w_lo + (W_F32 >> 4) + 2 * k), idesc_b, 1);
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k)   // A_t * W_t
                            umma_tf32_ts(tmem_base, a_tm + 32 * kk + 8 * k, desc_make(DESC_HI_SW128, w_lo + 2 * k), idesc, 1);
                        umma_commit_mc_a(done_a + 8 * (kb & (ND - 1)), (uint16_t)3);
                    }
                    umma_commit(mlp_out);
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ----- accumulator owners: thread = token row, 64 running fp32 sums starting from the bias -----
        const int q = warp & 3;
        const int row = q * 32 + lane;
        for (int i = tid - 128; i < a.L * XD; i += 128) tokb[i] = __ldg(a.packed + pk.tok_b(i / XD) + (i % XD));
        for (int i = tid - 128; i < S * XD / 4; i += 128) cp_async16(slots0 + i * 4, a.packed + pk.slots() + i * 4);
        named_bar_sync(1, 128);   // the four accumulator warps only
        float acc[XD];
#pragma unroll
        for (int j = 0; j < XD / 4; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(a.conv_b + 4 * j));
            acc[4 * j] = v.x; acc[4 * j + 1] = v.y; acc[4 * j + 2] = v.z; acc[4 * j + 3] = v.w;
        }
        const int nch0 = ((kblocks + 1) / 2 + CHUNK - 1) / CHUNK, nch1 = (kblocks / 2 + CHUNK - 1) / CHUNK;
        for (int ch = 0; ch < nch0; ++ch) {
            const int buf = ch & 1;
#pragma unroll
            for (int w = 0; w < 2; ++w) {          // fixed merge order: issuer 0's chunk, then issuer 1's
                if (w == 1 && ch >= nch1) break;
                mbar_wait(&cfull[2 * w + buf], (ch >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < XD / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * w + buf) * XD + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
                }
                tc_fence_before();
                mbar_arrive(&cempty[2 * w + buf]);
            }
        }
        // every conv MMA has retired (the last commits cover them all): the feature ring is dead, X takes its place
        mbar_arrive(acc_done);
        const bool rowv = row < R;
        const int img = rowv ? row / n : 0, j = rowv ? row - img * n : 0;
        const bool live = rowv && img < nimg;
        const uint32_t t_op = tmem_base + ((uint32_t)(q * 32) << 16) + OP_COL0;
        auto to_tmem = [&](const float (&v)[XD]) {        // fp32 | bf16 | bf16 remainder forms of this row's 64 activations
            uint32_t f[32], xb[16], rb[16];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float x0 = v[32 * h + 2 * i], x1 = v[32 * h + 2 * i + 1];
                    f[2 * i] = __float_as_uint(x0); f[2 * i + 1] = __float_as_uint(x1);
                    const float r0 = x0 - __uint_as_float(f[2 * i] & 0xFFFFE000u), r1 = x1 - __uint_as_float(f[2 * i + 1] & 0xFFFFE000u);
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xb[i]) : "f"(x1), "f"(x0));
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(rb[i]) : "f"(r1), "f"(r0));
                }
                tmem_st_32x32(t_op + 32 * h, f);
                tmem_st_32x16(t_op + 64 + 16 * h, xb);
                tmem_st_32x16(t_op + 96 + 16 * h, rb);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(mlp_in);
        };
        {
            float* xo = (a.x_out && live) ? a.x_out + ((size_t)(b0 + img) * n + j) * XD : nullptr;
#pragma unroll
            for (int e4 = 0; e4 < XD / 4; ++e4) {
                float4 xv = make_float4(fmaxf(acc[4 * e4], 0.f), fmaxf(acc[4 * e4 + 1], 0.f), fmaxf(acc[4 * e4 + 2], 0.f),
                                        fmaxf(acc[4 * e4 + 3], 0.f));
                if (!live) xv = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 pv = __ldg(reinterpret_cast<const float4*>(a.pe + j * XD + e4 * 4));
                if (rowv) *reinterpret_cast<float4*>(Xs + row * LDX + e4 * 4) = xv;
                if (xo) *reinterpret_cast<float4*>(xo + e4 * 4) = xv;
                acc[4 * e4] = xv.x + pv.x; acc[4 * e4 + 1] = xv.y + pv.y; acc[4 * e4 + 2] = xv.z + pv.z; acc[4 * e4 + 3] = xv.w + pv.w;
            }
        }
        to_tmem(acc);
        for (int l = 0; l < a.L; ++l) {
            const bool lastl = l + 1 == a.L;
            const float* bp = tokb + l * XD;
            mbar_wait(mlp_out, (uint32_t)l & 1u);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < XD / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 bv = *reinterpret_cast<const float4*>(bp + c * 32 + 4 * i4);
                    float v0 = __uint_as_float(r[4 * i4]) + bv.x, v1 = __uint_as_float(r[4 * i4 + 1]) + bv.y;
                    float v2 = __uint_as_float(r[4 * i4 + 2]) + bv.z, v3 = __uint_as_float(r[4 * i4 + 3]) + bv.w;
                    if (!lastl) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
                    acc[c * 32 + 4 * i4] = v0; acc[c * 32 + 4 * i4 + 1] = v1; acc[c * 32 + 4 * i4 + 2] = v2; acc[c * 32 + 4 * i4 + 3] = v3;
                }
            }
            if (!lastl) {
                tc_fence_before();
                to_tmem(acc);
            } else if (rowv) {
#pragma unroll
                for (int e4 = 0; e4 < XD / 4; ++e4)
                    *reinterpret_cast<float4*>(Ka + row * LDX + e4 * 4) = make_float4(acc[4 * e4], acc[4 * e4 + 1], acc[4 * e4 + 2], acc[4 * e4 + 3]);
            }
        }
        cp_async_wait_all();   // the staged initial slots (issued long ago)
    } else if (warp >= 8) {
        // ----- splitters: two groups of four warps on alternate k-blocks; thread = token row -----
        const int sg = (warp - 8) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool valid = row < R;
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + OP_COL0;
        PROF_DECL(sp_fullA); PROF_DECL(sp_done); PROF_DECL(sp_fullW); PROF_DECL(sp_st);
        int stage = sg % NA;
        uint32_t aphase = 0;
        for (int kb = sg; kb < kblocks; kb += 2) {
            const int ob = kb & (NO - 1);
            PROF_T(sp_fullA, mbar_wait(&fullA[stage], aphase));
            if (tid == 256 || tid == 384) TRACE(kb, 2);
            uint32_t f[32];
            {
                const uint8_t* src = smem + stage * a.a_stage + row * 128;
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (valid) v = *reinterpret_cast<const uint4*>(src + ((c ^ sw) << 4));   // SWIZZLE_128B: 16-byte chunk ^= row % 8
                    f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w;
                }
            }
            uint32_t xb[16], rb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float x0 = __uint_as_float(f[2 * i]), x1 = __uint_as_float(f[2 * i + 1]);
                const float r0 = x0 - __uint_as_float(f[2 * i] & 0xFFFFE000u), r1 = x1 - __uint_as_float(f[2 * i + 1] & 0xFFFFE000u);
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xb[i]) : "f"(x1), "f"(x0));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(rb[i]) : "f"(r1), "f"(r0));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyA[stage]);               // the stage lives in registers now
            if (kb >= NO) {
                const int j = kb - NO;
                PROF_T(sp_done, mbar_wait(&done[j & (ND - 1)], (uint32_t)(j >> 3) & 1u));   // the MMAs that read this operand buffer have retired
                tc_fence_after();
            }
            if (tid == 256 || tid == 384) TRACE(kb, 3);
            const uint32_t t0 = t_lane + (uint32_t)ob * 64;
            PROF_T(sp_st, tmem_st_32x32(t0, f); tmem_st_32x16(t0 + 32, xb); tmem_st_32x16(t0 + 48, rb); tmem_st_wait());
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&opfull[ob]);
            if (tid == 256 || tid == 384) TRACE(kb, 4);
            stage += 2; if (stage >= NA) { stage -= NA; aphase ^= 1; }
        }
        if (tid == 256) { PROF_STORE(g_prof_head, 7, sp_fullA); PROF_STORE(g_prof_head, 8, sp_done); PROF_STORE(g_prof_head, 9, sp_fullW); PROF_STORE(g_prof_head, 10, sp_st); }
        if (a.tc_gates && a.gw_early && a.iters > 1) {
            // ----- these eight warps are idle until the loop: W_ih, then W_hh -> staging (inside the dead feature ring,
            //       above X and K) -> TMEM, while the accumulator owners and issuer 0 run the to_k MLP -----
            constexpr int GW_LD = XD * 4 + 16;
            uint8_t* wst = smem + a.gw_early;
            const int t8 = tid - 256, w8 = warp - 8;
            const int grow = (w8 >> 2) * 128 + (w8 & 3) * 32 + lane;          // gate row of this thread within a matrix
            const bool gv = grow < XG;
            const uint32_t tm = tmem_base + ((uint32_t)((w8 & 3) * 32) << 16);
            auto stage = [&](const float* w) {
                for (int i = t8; i < XG * (XD / 4); i += 256) cp_async16(wst + (i >> 4) * GW_LD + (i & 15) * 16, w + (size_t)(i >> 4) * XD + (i & 15) * 4);
                cp_async_wait_all();
                named_bar_sync(2, 256);
            };
            auto load_row = [&](uint32_t (&f)[64], uint32_t (&rb)[32]) {
                const uint8_t* wrow = wst + min(grow, XG - 1) * GW_LD;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (gv) v = *reinterpret_cast<const uint4*>(wrow + 16 * c);
                    f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float r0 = __uint_as_float(f[2 * i]) - __uint_as_float(f[2 * i] & 0xFFFFE000u);
                    const float r1 = __uint_as_float(f[2 * i + 1]) - __uint_as_float(f[2 * i + 1] & 0xFFFFE000u);
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(rb[i]) : "f"(r1), "f"(r0));
                }
            };
            auto store_tile = [&](int t, const uint32_t (&f)[64], const uint32_t (&rb)[32]) {
                const uint32_t c0 = tm + gw_col(t);
                uint32_t x[32], y[16];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) x[i] = f[32 * h + i];
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] = rb[16 * h + i];
                    tmem_st_32x32(c0 + 32 * h, x);
                    tmem_st_32x16(c0 + 64 + 16 * h, y);
                }
            };
            uint32_t f[64], rb[32];
            named_bar_sync(2, 256);              // both splitter groups have read their last feature stage
            stage(a.gru_w_ih);
            load_row(f, rb);
            mbar_wait(acc_done, 0);              // conv accumulators (columns 64-255) and operand buffers are dead
            tc_fence_after();
            store_tile(w8 >> 2, f, rb);          // tiles 0, 1
            tmem_st_wait();
            named_bar_sync(2, 256);              // every row of W_ih has been read
            stage(a.gru_w_hh);
            load_row(f, rb);
            if ((w8 >> 2) == 0) {                // tile 2 now; tile 3 (columns of the MLP operands) after the MLP
                store_tile(2, f, rb);
                tmem_st_wait();
            }
            tc_fence_before();
            named_bar_sync(0, HT);               // the phase-A barrier (bar.sync 0 from this branch, same barrier as below)
            if ((w8 >> 2) == 1) {
                tc_fence_after();
                store_tile(3, f, rb);
                tmem_st_wait();
                tc_fence_before();
            }
        } else {
            tc_fence_before();
            named_bar_sync(0, HT);
        }
    }
    if (warp < 8) {
        tc_fence_before();
        named_bar_sync(0, HT);                   // end of phase A + MLP (the splitter warps arrive from their branch)
    }
    cluster_arrive_relaxed();   // no more multicasts / remote arrivals from this CTA; matched by cluster_wait() before exit
    tc_fence_after();
    PROF_END(phaseA); PROF_BEGIN(mlp);

    // =========================================== phase B ===========================================================
    for (int img = 0; img < G; ++img)
        for (int idx = tid; idx < S * XD; idx += HT) slots[img * S * XD + idx] = slots0[idx];

    PROF_END(mlp); PROF_BEGIN(loop);
    const float* K = Ka;
    // ---- GRU gate GEMMs on the tensor cores (SR <= 32) ---------------------------------------------------------------
    // gates^T[g, sr] = W[g, :] . x[sr, :]: the 192 gate rows are the M side (two 128-lane tiles per matrix), the slot rows
    // the N side (32 columns).  W_ih and W_hh live in TMEM for the whole loop as [fp32 (64 cols) | bf16 remainder (32)]
    // per tile (4 x 96 = 384 columns) -- thread = gate row, straight from the PyTorch (3d, d) layout; the per-step
    // operands (updates, slots) are small K-major tiles in shared memory in three forms: fp32 (read as tf32), fp32
    // remainder x - trunc19(x), bf16.  Same compensated product: W_t x_t + W_t x_r + W_r x.  80 MMAs of N = 32 per
    // step (16-clk floor) issued by four threads, against 491k FMAs on the CUDA cores.
    const bool tcg = a.tc_gates != 0 && a.iters > 1;
    constexpr int GB_TILE32 = 2 * 32 * 128, GB_TILE16 = 2 * 32 * 64;     // two 32-channel k-blocks of 32 rows
    uint8_t* gb = reinterpret_cast<uint8_t*>(Wsm);                        // [x | h] x [fp32 | fp32 rem | bf16]
    constexpr int GB_OP = 2 * GB_TILE32 + GB_TILE16;
    const int g_t = warp >> 2, g_row = (g_t & 1) * 128 + (warp & 3) * 32 + lane;   // this thread's (tile, gate row) for W / D
    const bool g_valid = g_row < XG;
    const uint32_t g_tm = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float g_bias = 0.f;
    auto put_operand = [&](int o, int sr, int e, float v) {              // element (sr, e) of operand o into its three tiles
        uint8_t* base = gb + o * GB_OP;
        const int kb = e >> 5, c = e & 31;
        const int o32 = kb * (32 * 128) + sr * 128 + (((c >> 2) ^ (sr & 7)) << 4) + ((c & 3) << 2);
        *reinterpret_cast<float*>(base + o32) = v;
        *reinterpret_cast<float*>(base + GB_TILE32 + o32) = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        const int o16 = kb * (32 * 64) + sr * 64 + (((c >> 3) ^ ((sr >> 1) & 3)) << 4) + ((c & 7) << 1);
        *reinterpret_cast<__nv_bfloat16*>(base + 2 * GB_TILE32 + o16) = __float2bfloat16_rn(v);
    };
    // W_ih, then W_hh: global -> shared memory with coalesced cp.async (row stride 272 bytes: the row-per-thread reads are
    // conflict-free; reading the rows straight from global costs 32 L1 lines per load instruction) -> TMEM.  Each transfer
    // (48 KB) hides behind a phase of the first iteration: W_ih behind the dots, W_hh behind normalise + update.
    constexpr int GW_LD = XD * 4 + 16;
    uint8_t* wst = gb + 2 * GB_OP;                                       // staging for one matrix: 192 rows x 272 B
    auto stage_w = [&](const float* w) {
        for (int i = tid; i < XG * (XD / 4); i += HT) cp_async16(wst + (i >> 4) * GW_LD + (i & 15) * 16, w + (size_t)(i >> 4) * XD + (i & 15) * 4);
    };
    auto w_to_tmem = [&](int mat) {                                     // all threads call; warps of matrix `mat` convert
        cp_async_wait_all();
        __syncthreads();
        if ((g_t >> 1) == mat) {
            uint32_t f[32], rb[16];
            const uint8_t* wrow = wst + min(g_row, XG - 1) * GW_LD;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (g_valid) v = *reinterpret_cast<const uint4*>(wrow + 128 * h + 16 * c);
                    f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float r0 = __uint_as_float(f[2 * i]) - __uint_as_float(f[2 * i] & 0xFFFFE000u);
                    const float r1 = __uint_as_float(f[2 * i + 1]) - __uint_as_float(f[2 * i + 1] & 0xFFFFE000u);
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(rb[i]) : "f"(r1), "f"(r0));
                }
                tmem_st_32x32(g_tm + gw_col(g_t) + 32 * h, f);
                tmem_st_32x16(g_tm + gw_col(g_t) + 64 + 16 * h, rb);
            }
            tmem_st_wait();
            tc_fence_before();
        }
        __syncthreads();      // the staging area is free again; W of this matrix is visible to the issuers
    };
    const bool gw_late = tcg && a.gw_early == 0;     // no room for the early staging: W moves to TMEM inside iteration 0
    if (tcg) {
        if (gw_late) stage_w(a.gru_w_ih);
        g_bias = g_valid ? __ldg((g_t >> 1 ? a.gru_b_hh : a.gru_b_ih) + g_row) : 0.f;
        for (int i = tid; i < 2 * GB_OP / 16; i += HT) reinterpret_cast<uint4*>(gb)[i] = make_uint4(0u, 0u, 0u, 0u);   // rows >= SR stay zero
        __syncthreads();
        for (int idx = tid; idx < SR * XD; idx += HT) put_operand(1, idx / XD, idx % XD, slots[idx]);
    } else if (a.iters > 1) {
        // the to_k weights (and the ring) are dead: bring in the GRU block while the first attention pass runs
        for (int i = tid; i < W_FLOATS / 4; i += HT) cp_async16(Wsm + i * 4, a.packed + pk.gru_wih_t() + i * 4);
    }

    // Thread maps of the loop, computed once (no runtime integer division inside the iterations).  The shared-memory
    // pipe delivers one wavefront per clock per SM against four FMA issue slots: every map below keeps its operands in
    // registers across as many FMAs as possible.
    //   dots / sigmoid: thread = (slot subset dq, token row dr): key row in registers, the subset's slot rows broadcast
    //   update:         thread = (token subset uq, image, slot chunk, feature quad): <= 16 slots x 4 features in registers;
    //                   the token subsets are merged in fixed order
    //   gates:          thread = (row tile, {ih,hh}, gate quad): 4 slot rows x 4 gate columns in registers; 15 warps busy,
    //                   i.e. the 491k MACs of a GRU step spread evenly over the four FMA issue ports
    const int nq = min(HT / R, S);                   // slot subsets
    const int sps = (S + nq - 1) / nq;               // slots per subset
    const int dq = tid / R, dr = tid - dq * R;
    const int dimg = dr / n, dj = dr - dimg * n;
    const bool d_on = dq < nq && dq * sps < S;
    const int di0 = dq * sps, di1 = min(S, di0 + sps);
    const bool d_live = dimg < nimg;
    const int SP4 = SP / 4, nsc = (SP4 + 3) / 4;     // slot quads per image, slot chunks of <= 4 quads
    const int upt = G * nsc * 16;                    // update threads per token subset
    const int nuq = min(6, HT / upt);                // 6 * 64 floats per slot row: the partials fit the gate scratch they alias
    const int uq = tid / upt, ur = tid - uq * upt;
    const int uimg = ur / (nsc * 16), usc = (ur / 16) % nsc, ue4 = ur % 16;
    const int uc4 = min(4, SP4 - 4 * usc), ui0 = 16 * usc;
    float* upart = gates;                            // [nuq][SR][64] partial updates (dead before the gates are written)
    const int gct = tid % 96, grt = tid / 96;        // gates: column tile (48 quads x {ih,hh}), row-tile group (5 groups)
    const int gwhich = gct / 48, gg0 = (gct - gwhich * 48) * 4;
    constexpr int GR = 4;                            // slot rows per gate tile
    const int n_rt = (SR + GR - 1) / GR;
    const int rs_img0 = warp / S, rs_i0 = warp - rs_img0 * S;                  // row sums: this warp's first two slot rows
    const int rs_img1 = (warp + HW_) / S, rs_i1 = warp + HW_ - rs_img1 * S;

#ifdef SCOUTER_PROF
    if (tid == 0) g_prof_head[blockIdx.x * 32 + 18] = (unsigned long long)(clock64() - prof_begin_loop);   // loop set-up
    long long lp_t = clock64(), lp_acc[6] = {0, 0, 0, 0, 0, 0};
#define LP_STAMP(k) do { const long long _n = clock64(); lp_acc[k] += _n - lp_t; if (tid == 0 && blockIdx.x == 0) g_trace_head[64 * 8 + it * 8 + (k)] = (unsigned long long)(_n - lp_t); lp_t = _n; } while (0)
#else
#define LP_STAMP(k)
#endif
    for (int it = 0; it < a.iters; ++it) {
        const bool last = it == a.iters - 1;
        if (d_on) {                                  // dots[img][i][j] = scale * <slot, key>, kept as attnT[img][j][i]
            float4 kreg[XD / 4];
            const float4* kp = reinterpret_cast<const float4*>(K + dr * LDX);
#pragma unroll
            for (int e4 = 0; e4 < XD / 4; ++e4) kreg[e4] = kp[e4];
            for (int i = di0; i < di1; ++i) {
                const float4* sp = reinterpret_cast<const float4*>(slots + (dimg * S + i) * XD);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int e4 = 0; e4 < XD / 4; ++e4) {
                    const float4 s4 = sp[e4];
                    a0 = fmaf(s4.x, kreg[e4].x, a0); a1 = fmaf(s4.y, kreg[e4].y, a1);
                    a2 = fmaf(s4.z, kreg[e4].z, a2); a3 = fmaf(s4.w, kreg[e4].w, a3);
                }
                attnT[dr * SP + i] = ((a0 + a1) + (a2 + a3)) * 0.125f;
            }
        }
        __syncthreads();
        if (gw_late && it == 0) {
            w_to_tmem(0);
            stage_w(a.gru_w_hh);
        }
        LP_STAMP(0);
        for (int sr = warp, k = 0; sr < SR; sr += HW_, ++k) {   // row sums r_bi, lane-strided then a fixed shuffle tree
            int img, i;
            if (k == 0) { img = rs_img0; i = rs_i0; } else if (k == 1) { img = rs_img1; i = rs_i1; } else { img = sr / S; i = sr - img * S; }
            float s_ = 0.f;
            for (int j = lane; j < n; j += 32) s_ += attnT[(img * n + j) * SP + i];
            s_ = warp_sum(s_);
            if (lane == 0) rsum[sr] = s_;
        }
        __syncthreads();
        if (d_on) {                                  // attn = sigmoid(D / r * t), in place; t_b summed in fixed order by every thread
            const float* rs = rsum + dimg * S;
            float t_ = 0.f;
            for (int i = 0; i < S; ++i) t_ += rs[i];
            for (int i = di0; i < di1; ++i) {
                float* p = attnT + dr * SP + i;
                const float at = sigm(*p / rs[i] * t_);
                *p = at;
                if (last && d_live && a.attn) a.attn[((size_t)(b0 + dimg) * S + i) * n + dj] = at;
            }
        }
        __syncthreads();
        LP_STAMP(1);
        if (uq < nuq) {                              // updates[sr][e] = sum_j attn * X / d
            const float* xp = Xs + uimg * n * LDX + ue4 * 4;
            const float* ap = attnT + uimg * n * SP + ui0;
            float* out = upart + ((size_t)(uq * SR + uimg * S + ui0)) * XD + ue4 * 4;
            const int ns = S - ui0;
            switch (uc4) {
                case 1: update_partial<1>(xp, ap, n, SP, uq, nuq, out, ns); break;
                case 2: update_partial<2>(xp, ap, n, SP, uq, nuq, out, ns); break;
                case 3: update_partial<3>(xp, ap, n, SP, uq, nuq, out, ns); break;
                default: update_partial<4>(xp, ap, n, SP, uq, nuq, out, ns); break;
            }
        }
        __syncthreads();
        for (int idx = tid; idx < SR * XD; idx += HT) {     // merge the token subsets in fixed order
            float s_ = upart[idx];
            for (int q = 1; q < nuq; ++q) s_ += upart[q * SR * XD + idx];
            upd[idx] = s_ * (1.0f / XD);
            if (tcg && !last) put_operand(0, idx / XD, idx % XD, s_ * (1.0f / XD));
        }
        if (tcg && !last) fence_proxy_async();   // operand tiles (updates here, slots in the cell pass) -> tensor-core reads
        __syncthreads();
        LP_STAMP(2);
        if (last) {
            for (int sr = warp; sr < SR; sr += HW_) {
                float s_ = upd[sr * XD + lane] + upd[sr * XD + 32 + lane];
                s_ = warp_sum(s_);
                if (lane == 0) usum[sr] = s_;
            }
        } else {
            if (it == 0) {
                if (gw_late) {
                    w_to_tmem(1);
                } else if (!tcg) {
                    cp_async_wait_all();
                    __syncthreads();
                }
            }
            LP_STAMP(3);
            if (tcg) {
                if (warp < 4 && lane == 0) {                 // issuer of tile `warp`: (matrix, gate half)
                    constexpr uint32_t idesc = idesc_tf32(128, 32), idesc_b = idesc_bf16(128, 32);
                    const uint32_t t = (uint32_t)warp, o = t >> 1;
                    const uint32_t d_t = tmem_base + gd_col((int)t), w_t = tmem_base + gw_col((int)t);
                    const uint32_t b32 = desc_lo(smem_u32(gb + o * GB_OP)), br32 = b32 + (GB_TILE32 >> 4), b16 = b32 + (2 * GB_TILE32 >> 4);
                    tc_fence_after();
#pragma unroll
                    for (uint32_t j = 0; j < 8; ++j)     // W_t x_t
                        umma_tf32_ts(d_t, w_t + 8 * j, desc_make(DESC_HI_SW128, b32 + (j >> 2) * (32 * 128 >> 4) + 2 * (j & 3)), idesc, j != 0);
#pragma unroll
                    for (uint32_t j = 0; j < 8; ++j)     // W_t x_r
                        umma_tf32_ts(d_t, w_t + 8 * j, desc_make(DESC_HI_SW128, br32 + (j >> 2) * (32 * 128 >> 4) + 2 * (j & 3)), idesc, 1);
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j)     // W_r x (bf16)
                        umma_bf16_ts(d_t, w_t + 64 + 8 * j, desc_make(DESC_HI_SW64, b16 + (j >> 1) * (32 * 64 >> 4) + 2 * (j & 1)), idesc_b, 1);
                    umma_commit(gbar);
                }
                mbar_wait(gbar, (uint32_t)it & 1u);
                tc_fence_after();
                uint32_t r[32];
                tmem_ld_32x32(g_tm + gd_col(g_t), r);
                tmem_ld_wait();
                if (g_valid) {
                    float* gp = gates + (g_t >> 1) * XG + g_row;         // [sr][ih | hh][gate]
#pragma unroll
                    for (int sr = 0; sr < 32; ++sr)
                        if (sr < SR) gp[sr * 2 * XG] = __uint_as_float(r[sr]) + g_bias;
                }
                tc_fence_before();
            } else if (grt < 5) {                            // gate pre-activations gi = W_ih u + b_ih, gh = W_hh h + b_hh
                const float* WT = Wsm + gwhich * XD * XG + gg0;                    // [e][192], this thread's gate quad
                const float4 b4 = *reinterpret_cast<const float4*>(Wsm + 2 * XD * XG + gwhich * XG + gg0);
                const float* src = gwhich ? slots : upd;
                for (int rt = grt; rt < n_rt; rt += 5) {
                    const int r0 = rt * GR;
                    float4 acc[GR];
#pragma unroll
                    for (int r = 0; r < GR; ++r) acc[r] = b4;
#pragma unroll 4
                    for (int e4 = 0; e4 < XD / 4; ++e4) {
                        const float4 w0 = *reinterpret_cast<const float4*>(WT + (e4 * 4 + 0) * XG);
                        const float4 w1 = *reinterpret_cast<const float4*>(WT + (e4 * 4 + 1) * XG);
                        const float4 w2 = *reinterpret_cast<const float4*>(WT + (e4 * 4 + 2) * XG);
                        const float4 w3 = *reinterpret_cast<const float4*>(WT + (e4 * 4 + 3) * XG);
#pragma unroll
                        for (int r = 0; r < GR; ++r) {
                            const float4 v = *reinterpret_cast<const float4*>(src + min(r0 + r, SR - 1) * XD + e4 * 4);
                            float4& o = acc[r];
                            o.x = fmaf(v.x, w0.x, o.x); o.y = fmaf(v.x, w0.y, o.y); o.z = fmaf(v.x, w0.z, o.z); o.w = fmaf(v.x, w0.w, o.w);
                            o.x = fmaf(v.y, w1.x, o.x); o.y = fmaf(v.y, w1.y, o.y); o.z = fmaf(v.y, w1.z, o.z); o.w = fmaf(v.y, w1.w, o.w);
                            o.x = fmaf(v.z, w2.x, o.x); o.y = fmaf(v.z, w2.y, o.y); o.z = fmaf(v.z, w2.z, o.z); o.w = fmaf(v.z, w2.w, o.w);
                            o.x = fmaf(v.w, w3.x, o.x); o.y = fmaf(v.w, w3.y, o.y); o.z = fmaf(v.w, w3.z, o.z); o.w = fmaf(v.w, w3.w, o.w);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < GR; ++r)
                        if (r0 + r < SR) *reinterpret_cast<float4*>(gates + (r0 + r) * 2 * XG + gwhich * XG + gg0) = acc[r];
                }
            }
            __syncthreads();
            LP_STAMP(4);
            for (int idx = tid; idx < SR * XD; idx += HT) {   // GRU cell, gate order [r|z|n]
                const int sr = idx / XD, e = idx - sr * XD;
                const float* gi = gates + sr * 2 * XG;
                const float* gh = gi + XG;
                const float rg_ = sigm(gi[e] + gh[e]);
                const float zg = sigm(gi[XD + e] + gh[XD + e]);
                const float ng = tanhf(gi[2 * XD + e] + rg_ * gh[2 * XD + e]);
                const float hp = slots[idx];
                const float hn = (hp - ng) * zg + ng;          // ATen's form of (1-z)*n + z*h
                slots[idx] = hn;
                if (tcg) put_operand(1, sr, e, hn);
            }
        }
        __syncthreads();
        LP_STAMP(5);
    }
#ifdef SCOUTER_PROF
    if (tid == 0) for (int k = 0; k < 6; ++k) g_prof_head[blockIdx.x * 32 + 12 + k] = (unsigned long long)lp_acc[k];
#endif

    for (int idx = tid; idx < nimg * a.C; idx += HT) {
        const int img = idx / a.C, c = idx - img * a.C;
        float s = 0.f;
        for (int m = 0; m < a.spc; ++m) s += usum[img * S + c * a.spc + m];
        a.logits[(size_t)(b0 + img) * a.C + c] = (float)a.loss_status * s;
    }
    if (a.attn_sum) {
        for (int img = warp; img < nimg; img += HW_) {        // per-image sum of the final attention (area loss), fixed order
            float s = 0.f;
            for (int k = lane; k < n * S; k += 32) {
                const int j = k / S, i = k - j * S;
                s += attnT[(img * n + j) * SP + i];
            }
            s = warp_sum(s);
            if (lane == 0) a.attn_sum[b0 + img] = s;
        }
    }
    PROF_END(loop);
    if (tid == 0) { PROF_STORE(g_prof_head, 0, phaseA); PROF_STORE(g_prof_head, 1, mlp); PROF_STORE(g_prof_head, 2, loop); }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    cluster_wait();       // the peer may still be arriving on this CTA's barriers until it has left phase A
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

// Shared-memory map for (G, n, S, L); returns total dynamic bytes (incl. alignment slack) or 0 when it does not fit.
//   phase A + to_k MLP: [NA feature stages][NW weight slots]
//   loop:               [X][K][scratch][slots, updates, sums][GRU block] alias the rings (the GRU block is fetched after the MLP)
size_t layout(int G, int n, int S, int L, FusedArgs* out) {
    const int R = G * n, SR = G * S;
    const int R8 = (R + 7) & ~7;
    const size_t a_stage = (size_t)R8 * 128;
    const size_t tok_end = align_up((size_t)(2 * R * LDX + kb_floats(G, n, S)) * 4, 16);
    const size_t small = align_up((size_t)(2 * SR * XD + 2 * SR + 64) * 4, 16);
    const size_t off_small = tok_end;
    const size_t off_gru = align_up(off_small + small, 1024);   // GRU block / tensor-core gate operand tiles (swizzled: 1024-aligned)
    static int nw_env = [] { const char* e = getenv("SCOUTER_HEAD_NW"); int v = e ? atoi(e) : 6; return v < 2 ? 2 : (v > MAX_NW ? MAX_NW : v); }();
    // the weight tile comes from L2 through a multicast TMA whose latency is ~2-3 k-blocks: prefer a deep weight ring, then
    // as many feature stages as fit
    static int na_env = [] { const char* e = getenv("SCOUTER_HEAD_NA"); int v = e ? atoi(e) : MAX_NA; return v < 2 ? 2 : (v > MAX_NA ? MAX_NA : v); }();
    for (int nw = nw_env; nw >= 2; --nw)
        for (int na = na_env; na >= 3; --na) {
            if ((size_t)na * a_stage < (size_t)R * LDX * 4) continue;   // X is written while the weight ring still feeds the MLP
            const size_t off_w = na * a_stage;
            const size_t ring_end = off_w + (size_t)nw * W_SLOT;
            // GRU block of the FMA path (99.8 KB) or the [W_ih ; W_hh] staging of the tensor-core path (384 rows x 272 B)
            // GRU block of the FMA path (99.8 KB), or operand tiles (40 KB) + one-matrix staging (51 KB) of the tensor-core path
            const size_t body = std::max(ring_end, off_gru + (size_t)W_FLOATS * 4);
            const size_t off_bar = align_up(body, 16);
            const size_t total = off_bar + 512 + (size_t)(L + S) * XD * 4 + 1024;   // barriers, to_k biases, initial slots, slack
            if (total <= 227 * 1024) {
                if (out) {
                    out->na = na; out->nw = nw; out->a_stage = (int)a_stage; out->off_w = (int)off_w;
                    out->off_gru = (int)off_gru; out->off_small = (int)off_small; out->off_bar = (int)off_bar;
                }
                return total;
            }
        }
    return 0;
}

__global__ void split_w_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float v = w[i];
    out[i] = __float2bfloat16_rn(v);
    out[count + i] = __float2bfloat16_rn(v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u));
}

}  // namespace

// Images per unit: as many as fit a 128-row tile AND the shared-memory map (many slots need more scratch per image).
static int pick_group(int batch, int n, int S, int L) {
    for (int G = std::max(1, std::min(128 / n, batch)); G >= 1; --G)
        if (S * G <= 128 && layout(G, n, S, L, nullptr) != 0) return G;
    return 0;
}

bool head_fused_supported(const scouter_xslot_desc_t* d, int batch, int n, int channel) {
    static bool off = getenv("SCOUTER_NO_FUSED_HEAD") != nullptr;
    const int S = d->num_classes * d->slots_per_class;
    if (off || n > 128 || S > 32 || channel % 32 || d->to_k_layers < 1) return false;
    return pick_group(batch, n, S, d->to_k_layers) != 0 && encode_fn() != nullptr;
}

size_t head_fused_workspace_bytes(int channel) { return align_up((size_t)2 * XD * channel * 2, 1024); }

int head_fused_launch(const scouter_xslot_desc_t* d, const void* packed, const scouter_head_io_t* io, const float* feat_nhwc,
                      void* workspace, cudaStream_t s) {
    const int n = io->h * io->w;
    FusedArgs a;
    a.conv_b = io->conv_b; a.packed = (const float*)packed; a.pe = io->pe;
    a.gru_w_ih = d->gru_w_ih; a.gru_w_hh = d->gru_w_hh; a.gru_b_ih = d->gru_b_ih; a.gru_b_hh = d->gru_b_hh;
    a.x_out = io->x_out; a.logits = io->logits; a.attn = io->attn; a.attn_sum = io->attn_sum;
    a.B = io->batch; a.n = n;
    a.S = d->num_classes * d->slots_per_class; a.C = d->num_classes; a.spc = d->slots_per_class; a.L = d->to_k_layers;
    a.G = pick_group(io->batch, n, a.S, a.L);
    SC_CHECK_ARG(a.G > 0, SCOUTER_E_UNSUPPORTED, "head_fused: no unit size fits");
    static bool no_tc_gates = getenv("SCOUTER_HEAD_FMA_GATES") != nullptr;
    a.tc_gates = (!no_tc_gates && a.G * a.S <= 32) ? 1 : 0;
    a.kblocks = io->channel / 32;
    const size_t smem = layout(a.G, n, a.S, a.L, &a);
    SC_CHECK_ARG(smem, SCOUTER_E_UNSUPPORTED, "head_fused: shared-memory layout does not fit");
    {   // staging for one GRU matrix (192 rows x 272 B) above X and K, inside the feature ring that is dead by then
        const size_t lo = align_up((size_t)2 * a.G * n * LDX * 4, 1024), need = (size_t)XG * (XD * 4 + 16);
        static bool no_early = getenv("SCOUTER_HEAD_GW_LATE") != nullptr;
        a.gw_early = (a.tc_gates && !no_early && lo + need <= (size_t)a.off_w) ? (int)lo : 0;
    }
    a.iters = d->iters; a.loss_status = d->loss_status;
    EncodeTiledFn enc = encode_fn();
    SC_CHECK_ARG(enc, SCOUTER_E_UNSUPPORTED, "head_fused: cuTensorMapEncodeTiled is not available from the driver");
    // bf16 [W ; W - trunc19(W)] of conv1x1.weight: given by the caller (packed once per parameter version) or derived here
    const void* w_split = io->conv_w_split;
    if (!w_split) {
        const int count = XD * io->channel;
        split_w_bf16_kernel<<<cdiv(count, 256), 256, 0, s>>>(io->conv_w, (__nv_bfloat16*)workspace, count);
        SC_LAUNCH_CHECK();
        w_split = workspace;
    }
    const int R = a.G * n;
    const long long M = (long long)io->batch * n;
    CUtensorMap tmA, tmB, tmB2, tmT, tmT2;
    {
        cuuint64_t dims[2] = {(cuuint64_t)io->channel, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)io->channel * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)R};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)feat_nhwc, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "head_fused: cuTensorMapEncodeTiled(features) failed with %d", (int)r);
        cuuint64_t dimsB[2] = {(cuuint64_t)io->channel, (cuuint64_t)XD};
        cuuint32_t boxB[2] = {32, (cuuint32_t)XD};
        r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)io->conv_w, dimsB, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "head_fused: cuTensorMapEncodeTiled(conv1x1.weight) failed with %d", (int)r);
        cuuint64_t dimsB2[2] = {(cuuint64_t)io->channel, (cuuint64_t)2 * XD};
        cuuint64_t stridesB2[1] = {(cuuint64_t)io->channel * 2};
        cuuint32_t boxB2[2] = {32, (cuuint32_t)2 * XD};
        r = enc(&tmB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w_split, dimsB2, stridesB2, boxB2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "head_fused: cuTensorMapEncodeTiled(bf16 weight pair) failed with %d", (int)r);
        // to_k layers from the packed parameter block: (L*64, 64) fp32 and (L*128, 64) bf16 [W ; W_r]
        const XSlotPacked pk{a.S, a.L};
        cuuint64_t dimsT[2] = {(cuuint64_t)XD, (cuuint64_t)a.L * XD};
        cuuint64_t stridesT[1] = {(cuuint64_t)XD * 4};
        r = enc(&tmT, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)(a.packed + pk.tok_w_raw(0)), dimsT, stridesT, boxB, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "head_fused: cuTensorMapEncodeTiled(to_k weights) failed with %d", (int)r);
        cuuint64_t dimsT2[2] = {(cuuint64_t)XD, (cuuint64_t)a.L * 2 * XD};
        cuuint64_t stridesT2[1] = {(cuuint64_t)XD * 2};
        r = enc(&tmT2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(a.packed + pk.tok_w_pair(0)), dimsT2, stridesT2, boxB2, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "head_fused: cuTensorMapEncodeTiled(to_k bf16 pairs) failed with %d", (int)r);
    }
    SC_CUDA(cudaFuncSetAttribute(head_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int units = cdiv(io->batch, a.G);
    head_fused_kernel<<<2 * cdiv(units, 2), HT, smem, s>>>(tmA, tmB, tmB2, tmT, tmT2, a);   // clusters of two (padding CTA when odd)
    SC_LAUNCH_CHECK();
    return 0;
}

}  // namespace scouter

#ifdef SCOUTER_PROF
extern "C" int scouter_prof_read_head(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, scouter::g_prof_head, sizeof(unsigned long long) * n);
}
extern "C" int scouter_trace_read_head(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, scouter::g_trace_head, sizeof(unsigned long long) * n);
}
#endif
