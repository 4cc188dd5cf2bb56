// The xSlot head as ONE kernel (SURVEY 8(d) "fused head"): conv1x1 + bias + ReLU (sloter/slot_model.py:108-109), + PE
// (:110-111) and SlotAttention.forward (sloter/utils/slot_attention.py:44-96) without the projected tokens ever
// leaving the SM.
//
//   unit  = G <= 2 consecutive images whose R = G*n token rows (+ G extra rows, see below) fit one 128-row UMMA tile
//           (n = 49 -> 2 images, 81 -> 1); one CTA per unit, 512 threads, CTAs paired into clusters of two.
//   phase A (HBM-bound): the unit's (R x ch) fp32 feature slab streams through a deep TMA ring (NA stages of 32
//           channels).  Error-compensated tensor-core product as in umma_conv.cu (tf32 main product A_t*W_t + two bf16
//           correction products A*W_r, A_r*W; chunked TMEM accumulation merged in fp32 registers) -- but with the A
//           operands in TENSOR MEMORY: this GEMM is thin (N = 64), so in SS mode the MMAs' own operand reads (6 KB per
//           instruction, 48 KB per k-block) plus the split tiles saturate the 128 B/clk shared-memory pipe at 2.5x the
//           HBM time (measured: 1170 clk per k-block).  Here the splitter threads (thread = token row) read their row of
//           the TMA tile once, derive bf16(x) and bf16(x - trunc19(x)) in registers and tcgen05.st all three forms into
//           a TMEM operand buffer; the MMAs then read only the small weight tile from shared memory.
//           conv1x1.weight (fp32 + pre-split bf16 [W ; W_r]) streams through a ring, multicast to both CTAs of
//           the cluster (each loads half).
//           warp 0 = feature producers (two threads on alternate k-blocks), warp 3 = weight producer, warps 1 and 2 =
//           MMA issuers on alternate k-blocks (four TMEM accumulators), warps 4-7 = accumulator owners (thread = token
//           row), warps 8-15 = two groups of splitters on alternate k-blocks.
//   to_k MLP: on the tensor cores (TS mode, the weights ride the multicast weight ring).
//   loop (tensor cores; fixed-order arithmetic, bit-reproducible, every image independent of its unit mate):
//           everything that is constant over the iterations sits in TENSOR MEMORY as an A operand (lane = row):
//             keys K (R token rows + one row per image holding ksum = sum_j K_j), 96 columns [fp32 | bf16 remainder];
//             GRU weights, 288 columns: tile A = rows r|z with K = [W_ih | W_hh] (one accumulator pair gives gi + gh),
//             tile B = lanes 0-63 W_ih n-rows, lanes 64-127 W_hh n-rows;
//           what changes per iteration is a small K-major tile in shared memory (B operand, N = 32 slot rows):
//             slots h, updates u (three forms each: fp32, fp32 remainder, bf16), attention (tf32 hi / lo).
//           per iteration:  dots^T[token, slot] = K . h      (20 TS-mode MMAs; the ksum row delivers the row sums r_i =
//                                                              sum_j dots_ij for free: sum_j <h_i, K_j> = <h_i, ksum>)
//                           attn = sigmoid(dots / r * t)      (token threads, IEEE division: SURVEY D9)
//                           u^T[image x e, slot] = X^T . attn  (SS-mode MMAs; X^T is a static tile whose rows 0-63 / 64-127 are
//                                                              the two images, so one chain serves both -- each image reads
//                                                              only its own 16 columns; compensated like everything else:
//                                                              X_t a_hi + X_t a_lo + X_r a, X_t = tf32(X), X_r = bf16(X - X_t))
//                           gates^T = W . [u | h]              (80 TS-mode MMAs of N = 32 from four threads)
//                           GRU cell                           (all warps)
//           The last iteration needs only logits_i = sum_j attn_ij * rowsum(X_j) / d.
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

// phase-A shared memory: [NA feature stages | NW weight slots]; tensor memory: [4 accumulators | NO operand buffers]
constexpr int W_F32 = XD * 128;                  // 64 rows x 32 tf32
constexpr int W_SLOT = 2 * W_F32;                // [W fp32 | bf16 W | bf16 W_r]
constexpr int MAX_NW = 8;                        // weight slots (a.nw of them)
constexpr int NO = 4;                            // TMEM operand buffers of 64 columns: [fp32 A (32) | bf16 A (16) | bf16 A_r (16)]
constexpr int ND = 8;                            // "k-block retired" barriers (>= MAX_NW, NO); power of two (index = kb & 7)
constexpr int OP_COL0 = 4 * XD;                  // operand buffers start after the four 64-column accumulators
constexpr int MAX_NA = 8;
constexpr int CHUNK = 4;                         // k-blocks per TMEM accumulation chunk (per issuer; see umma_conv.cu)

// Tensor-memory map of the loop (512 columns).  The to_k MLP uses columns 0-63 (D) and 256-383 (operands), so the GRU
// weights (64-255, 384-479) can be filled while it runs; the keys take the MLP operand columns afterwards.
constexpr uint32_t C_ACC_RZ_U = 0;      // 32: gates r|z, W_ih part (also the accumulator of the update MMAs)
constexpr uint32_t C_ACC_N_U = 32;      // 32: lanes 0-63 = gi_n
constexpr uint32_t C_GA = 64;           // 192: [W_ih fp32 (64) | W_hh fp32 (64) | W_ih bf16 remainder (32) | W_hh bf16 remainder (32)], lanes = gate rows 0-127
constexpr uint32_t C_KEY = 256;         // 64: K fp32, lanes = token rows, rows R.. = ksum per image  (= the MLP operand layout [fp32 | bf16 | bf16
constexpr uint32_t C_DOT = 320;         // 32: dots^T (main product); during the gate GEMMs: gates r|z, W_hh part       remainder] at 256-383:
constexpr uint32_t C_KEY_REM = 352;     // 32: bf16 remainder of K                                                      one store path)
constexpr uint32_t C_GB = 384;          // 96: [fp32 (64) | bf16 remainder (32)], lanes 0-63 = W_ih rows 128.., lanes 64-127 = W_hh rows 128..
constexpr uint32_t C_ACC_N_H = 480;     // 32: lanes 64-127 = gh_n
constexpr int GW_LD = XD * 4 + 16;      // staging row stride of one GRU matrix (272 B: row-per-thread reads are conflict-free)
constexpr int GB_TILE32 = 2 * 32 * 128, GB_TILE16 = 2 * 32 * 64;     // two 32-channel k-blocks of 32 rows
constexpr int GB_OP = 2 * GB_TILE32 + GB_TILE16;                     // one operand: [fp32 | fp32 remainder | bf16]

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
    // shared-memory map (bytes from the 1024-aligned base); the loop regions alias the phase-A rings
    int na, nw, a_stage, off_w;
    int off_bar;
    int off_xt, xt_kb;         // X^T tile (tf32): kbt k-blocks of (64 G rows = image x feature) x 32 tokens, xt_kb bytes each; with one
                               // image per unit the M = 128 MMAs read 64 rows (8 KB) past the tile: slack behind it
    int off_xl, xl_kb;         // bf16(X - tf32(X)) in the same arrangement (64-byte rows): the correction operand of the update MMAs
    int off_atb;               // bf16(attn) tile for that correction product: kbt k-blocks of 32 slot rows x 32 tokens (2 KB each)
    int off_misc;              // xsum[128] | ksum[2][64] | rs[2][32] | usum[32]
    int off_wst;               // staging of one GRU matrix (192 x 272 B), live while the to_k MLP runs
    int off_kp;                // plain keys [R][LDX] for the ksum column sums (inside the dead weight ring)
    int off_gb;                // GRU / dots B operands: [u | h] x [fp32 | fp32 remainder | bf16]
    int off_at, at_form;       // attention tile [hi | lo], each kbt k-blocks of 32 slot rows x 32 tokens (at_form bytes per form)
    int off_grz, off_gn;       // gate exchange: sigmoid(r|z) [32][128], n pre-activations [32][128]
    int off_attnp;             // plain attention of the last iteration [R][SPP]
    int off_slots;             // slots [32][64]
};

// Loop transcendentals: ex2.approx / rcp.approx forms (~2 ulp).  Their error is far below the tf32-class terms the loop
// already carries (X^T) and enters linearly; the sum-normalisation itself keeps its IEEE division.
#ifdef SCOUTER_HEAD_PRECISE_MATH
__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float tanh_fast(float v) { return tanhf(v); }
#else
__device__ __forceinline__ float sigm(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float tanh_fast(float v) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * v)); }
#endif
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float trunc19_rem(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

#ifdef SCOUTER_PROF
__device__ unsigned long long g_prof_head[256 * 32];
__device__ unsigned long long g_trace_head[128 * 8];   // CTA 0: [k-block][event] clock64 timestamps
#define TRACE(kb, ev) do { if (blockIdx.x == 0 && (kb) < 128) g_trace_head[(kb) * 8 + (ev)] = clock64(); } while (0)
#else
#define TRACE(kb, ev)
#endif

template <int G>     // images per unit (1 or 2)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HT, 1)
head_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmT,
                  const __grid_constant__ CUtensorMap tmT2, const FusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
#ifdef SCOUTER_PROF
    const long long prof_entry = clock64();
    if (threadIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_prof_head[blockIdx.x * 32 + 28] = gt; }
#endif
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int n = a.n, S = a.S;
    const int R = G * n;            // token rows of this unit; rows R .. R+G-1 of the key tile hold ksum per image
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
    uint64_t* dbar = acc_done + 1;          // dots MMAs of an iteration retired (2 issuing threads)
    uint64_t* ubar = dbar + 1;              // update MMAs of an iteration retired (3 issuing threads)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ubar + 1);
    float* tokb = reinterpret_cast<float*>(smem + a.off_bar + 512);   // [L][64] to_k biases (read by every accumulator owner)
    float* slots0 = tokb + a.L * XD;                                   // [S][64] initial slots, staged during phase A

    // loop views (alias the phase-A rings; see layout())
    const int ksteps = (n + 7) >> 3;                // K = 8 steps of the update MMAs
    constexpr int SS = G == 2 ? 16 : 0;             // slot row of (image g, slot i) = g * SS + i
    constexpr int SPP = G == 2 ? 17 : 33;           // row stride of the plain attention (odd: conflict-free column reads)
    uint8_t* const xt = smem + a.off_xt;
    uint8_t* const xl = smem + a.off_xl;
    uint8_t* const atb = smem + a.off_atb;
    const int ksteps16 = (n + 15) >> 4;             // K = 16 steps of the bf16 correction MMAs
    float* const xsum = reinterpret_cast<float*>(smem + a.off_misc);       // [128] row sums of X
    float* const ksum = xsum + 128;                                         // [2][64]
    float* const rs = ksum + 128;                                           // [2][32] row sums r_i of the dots
    float* const usum = rs + 64;                                            // [32] sum_e updates of the last iteration
    uint8_t* const gb = smem + a.off_gb;
    uint8_t* const at = smem + a.off_at;
    float* const grz = reinterpret_cast<float*>(smem + a.off_grz);
    float* const gn = reinterpret_cast<float*>(smem + a.off_gn);
    float* const attnP = reinterpret_cast<float*>(smem + a.off_attnp);
    float* const slots = reinterpret_cast<float*>(smem + a.off_slots);

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
        mbar_init(dbar, 2);
        mbar_init(ubar, 3);
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
#ifdef SCOUTER_PROF
    if (threadIdx.x == 0) g_prof_head[blockIdx.x * 32 + 26] = (unsigned long long)(prof_begin_phaseA - prof_entry);
#endif

    const int kblocks = a.kblocks;
    uint8_t* const w_ring = smem + a.off_w;
    // small parameters that the serial tail reads once (PE rows, GRU matrices and biases): pull them into L2 now, while the
    // feature stream runs -- inside a forward they were evicted by gigabytes of activations since the last call
    if (warp >= 4 && warp < 8) {
        const int t = tid - 128;
        auto pf = [](const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
        for (int i = t; i < (n * XD * 4 + 127) / 128; i += 128) pf(reinterpret_cast<const char*>(a.pe) + i * 128);
        for (int i = t; i < XG * XD * 4 / 128; i += 128) { pf(reinterpret_cast<const char*>(a.gru_w_ih) + i * 128); pf(reinterpret_cast<const char*>(a.gru_w_hh) + i * 128); }
        if (t < 6) { pf(a.gru_b_ih + 32 * t); pf(a.gru_b_hh + 32 * t); }
    }

    // =========================================== phase A ===========================================================
    if (warp == 0) {
        if (lane < 2) {
            // ----- feature producers: lanes 0 and 1 issue alternate k-blocks (one thread sustains ~1 box per 345 clk, a full-rate
            //       stream needs more: profiles/r02_hbm_read_probe.txt); they run NA k-blocks ahead of the MMAs (the first NA
            //       boxes went out during set-up) -----
            int stage = lane;                       // (NA + lane) % NA
            uint32_t phase = 1;
            const int row0 = blockIdx.x * R;
            PROF_DECL(pa_wait);
            for (int kb = NA + lane; kb < kblocks; kb += 2) {
                PROF_T(pa_wait, mbar_wait(&emptyA[stage], phase ^ 1));
                mbar_arrive_expect_tx(&fullA[stage], (uint32_t)(R * 128));
                tma_load_2d(smem + stage * a.a_stage, &tmA, &fullA[stage], kb * 32, row0);
                TRACE(kb, 0);
                stage += 2; if (stage >= NA) { stage -= NA; phase ^= 1; }
            }
            if (lane == 0) PROF_STORE(g_prof_head, 3, pa_wait);
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
        // every conv MMA has retired (the last commits cover them all): the feature ring is dead, X^T takes its place
        mbar_arrive(acc_done);
#ifdef SCOUTER_PROF
        const long long t_own0 = clock64();
#define OWN_STAMP(slot) do { if (tid == 128) g_prof_head[blockIdx.x * 32 + (slot)] = (unsigned long long)(clock64() - t_own0); } while (0)
        if (tid == 128) g_prof_head[blockIdx.x * 32 + 19] = (unsigned long long)(t_own0 - prof_begin_phaseA);
#else
#define OWN_STAMP(slot)
#endif
        const bool rowv = row < R;
        const int img = rowv ? row / n : 0, j = rowv ? row - img * n : 0;
        const bool live = rowv && img < nimg;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t t_op = t_lane + OP_COL0;
        float* Kp = reinterpret_cast<float*>(smem + a.off_kp);
        // One pass per to_k layer plus one: pass 0 hands X over, pass l > 0 finishes layer l-1, the last pass turns the keys
        // into the A operand of the dots MMAs.  Every pass ends in the same store of this row as [fp32 | bf16 | bf16
        // remainder] into TMEM columns 256-383 -- the MLP operand layout IS the key layout (C_KEY / C_KEY_REM), and one
        // call site keeps the straight-line conversion code (cold instruction fetches: ~5 clk per instruction) single.
        for (int l = 0; l <= a.L; ++l) {
            if (l == 0) {
                // hand-over: X = relu(acc) -> global (optional) and, through 64 scratch columns of tensor memory (the dead conv
                // accumulator 0: a rolled loop cannot index the register array), into the X^T tiles; acc <- X + PE
                float* xo = (a.x_out && live) ? a.x_out + ((size_t)(b0 + img) * n + j) * XD : nullptr;
                float xs = 0.f;
                uint32_t f[32];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int e4 = 8 * h; e4 < 8 * h + 8; ++e4) {
                        float4 xv = make_float4(fmaxf(acc[4 * e4], 0.f), fmaxf(acc[4 * e4 + 1], 0.f), fmaxf(acc[4 * e4 + 2], 0.f),
                                                fmaxf(acc[4 * e4 + 3], 0.f));
                        if (!live) xv = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 pv = __ldg(reinterpret_cast<const float4*>(a.pe + j * XD + e4 * 4));
                        if (xo) *reinterpret_cast<float4*>(xo + e4 * 4) = xv;
                        xs += xv.x; xs += xv.y; xs += xv.z; xs += xv.w;
                        f[4 * (e4 - 8 * h)] = __float_as_uint(xv.x); f[4 * (e4 - 8 * h) + 1] = __float_as_uint(xv.y);
                        f[4 * (e4 - 8 * h) + 2] = __float_as_uint(xv.z); f[4 * (e4 - 8 * h) + 3] = __float_as_uint(xv.w);
                        acc[4 * e4] = xv.x + pv.x; acc[4 * e4 + 1] = xv.y + pv.y; acc[4 * e4 + 2] = xv.z + pv.z; acc[4 * e4 + 3] = xv.w + pv.w;
                    }
                    tmem_st_32x32(t_lane + 32 * h, f);
                }
                tmem_st_wait();
                OWN_STAMP(24);
                if (rowv) xsum[row] = xs;
                // X^T tile (tf32, the A operand of the update MMAs: row = 64 image + feature, K = token) and its bf16 remainder
                uint8_t* xcol = xt + (j >> 5) * a.xt_kb + img * (64 * 128) + ((j & 3) << 2);
                uint8_t* lcol = xl + (j >> 5) * a.xl_kb + img * (64 * 64) + ((j & 7) << 1);
                const uint32_t jc = (uint32_t)(j & 31) >> 2, jc8 = (uint32_t)(j & 31) >> 3;
#pragma unroll 1
                for (uint32_t e4 = 0; e4 < XD / 4; ++e4) {
                    uint32_t xr[4];
                    tmem_ld_32x4(t_lane + 4 * e4, xr);
                    tmem_ld_wait();
                    if (rowv) {
#pragma unroll
                        for (uint32_t t = 0; t < 4; ++t) {
                            const uint32_t e = 4 * e4 + t;
                            const float x = __uint_as_float(xr[t]), xh = to_tf32(x);
                            *reinterpret_cast<float*>(xcol + e * 128 + ((jc ^ (e & 7)) << 4)) = xh;
                            *reinterpret_cast<__nv_bfloat16*>(lcol + e * 64 + ((jc8 ^ ((e >> 1) & 3)) << 4)) = __float2bfloat16_rn(x - xh);
                        }
                    }
                }
                OWN_STAMP(25);
                // the K padding of the tiles (tokens n .. 16*ksteps16-1) must be exact zeros: 0 * garbage could be NaN
                {
                    const int t = tid - 128, g = t >> 6;
                    const uint32_t e = (uint32_t)t & 63u;
                    if (g < G) {
#pragma unroll 1
                        for (int jj = n; jj < 16 * ksteps16; ++jj) {
                            if (jj < 8 * ksteps)
                                *reinterpret_cast<float*>(xt + (jj >> 5) * a.xt_kb + (g * 64 + e) * 128 + (((((uint32_t)jj & 31) >> 2) ^ (e & 7)) << 4) +
                                                          ((jj & 3) << 2)) = 0.f;
                            *reinterpret_cast<__nv_bfloat16*>(xl + (jj >> 5) * a.xl_kb + (g * 64 + e) * 64 +
                                                              (((((uint32_t)jj & 31) >> 3) ^ ((e >> 1) & 3)) << 4) + ((jj & 7) << 1)) = __float2bfloat16_rn(0.f);
                        }
                    }
                }
                tc_fence_before();      // the scratch columns become the accumulator of the first MLP layer
                OWN_STAMP(20);
            } else {
                const bool lastl = l == a.L;
                const float* bp = tokb + (l - 1) * XD;
                mbar_wait(mlp_out, (uint32_t)(l - 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < XD / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_lane + c * 32, r);
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
                tc_fence_before();
                if (lastl) {
                    OWN_STAMP(22);
                    // keys: besides the token rows, one row per image holds ksum = sum_j K_j, so that the dots MMA delivers
                    // r_i = sum_j <h_i, K_j> = <h_i, ksum> in that lane.  Every MLP MMA has retired: the weight ring is dead and
                    // takes the plain key rows for the column sums (fixed order over j: position-independent).
                    if (rowv) {
#pragma unroll
                        for (int e4 = 0; e4 < XD / 4; ++e4)
                            *reinterpret_cast<float4*>(Kp + row * LDX + e4 * 4) = make_float4(acc[4 * e4], acc[4 * e4 + 1], acc[4 * e4 + 2], acc[4 * e4 + 3]);
                    }
                    named_bar_sync(1, 128);
                    {
                        const int t = tid - 128, g = t >> 6, e = t & 63;
                        if (g < G) {
                            float s_ = 0.f;
                            const float* kp = Kp + (size_t)g * n * LDX + e;
                            int jj = 0;
#pragma unroll 1
                            for (; jj + 8 <= n; jj += 8) {            // loads first (their latencies overlap), then the adds in token order
                                float kv[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) kv[u] = kp[(jj + u) * LDX];
#pragma unroll
                                for (int u = 0; u < 8; ++u) s_ += kv[u];
                            }
                            for (; jj < n; ++jj) s_ += kp[jj * LDX];
                            ksum[g * XD + e] = s_;
                        }
                    }
                    named_bar_sync(1, 128);
                    if (!rowv) {
                        const bool ks = row < R + G;
                        const float4* kq = reinterpret_cast<const float4*>(ksum + (ks ? row - R : 0) * XD);
#pragma unroll
                        for (int e4 = 0; e4 < XD / 4; ++e4) {
                            const float4 kv = ks ? kq[e4] : make_float4(0.f, 0.f, 0.f, 0.f);
                            acc[4 * e4] = kv.x; acc[4 * e4 + 1] = kv.y; acc[4 * e4 + 2] = kv.z; acc[4 * e4 + 3] = kv.w;
                        }
                    }
                }
            }
            {   // fp32 | bf16 | bf16 remainder forms of this row's 64 values -> TMEM columns 256-383
                uint32_t f[32], xb[16], rb[16];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float x0 = acc[32 * h + 2 * i], x1 = acc[32 * h + 2 * i + 1];
                        f[2 * i] = __float_as_uint(x0); f[2 * i + 1] = __float_as_uint(x1);
                        const float r0 = trunc19_rem(x0), r1 = trunc19_rem(x1);
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xb[i]) : "f"(x1), "f"(x0));
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(rb[i]) : "f"(r1), "f"(r0));
                    }
                    tmem_st_32x32(t_op + 32 * h, f);
                    tmem_st_32x16(t_op + 64 + 16 * h, xb);
                    tmem_st_32x16(t_op + 96 + 16 * h, rb);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(mlp_in);      // (after the last pass nobody waits on it any more)
            }
            if (l == 0) OWN_STAMP(21);
        }
        OWN_STAMP(23);
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
        if (a.iters > 1) {
            // ----- these eight warps are idle until the loop: W_ih, then W_hh -> staging (inside the dead feature ring) ->
            //       TMEM, while the accumulator owners and issuer 0 run the to_k MLP.  Thread = TMEM lane L:
            //       tile A (warps 8-11): gate row L of r|z, columns [W_ih | W_hh | remainders]; tile B (warps 12-15): lanes 0-63
            //       = W_ih row 128+L, lanes 64-127 = W_hh row 128+(L-64) -----
            uint8_t* wst = smem + a.off_wst;
            const int t8 = tid - 256, w8 = warp - 8;
            const bool tileA = w8 < 4;
            const int L = q * 32 + lane;
            const uint32_t tm = tmem_base + ((uint32_t)(q * 32) << 16);
            auto stage_w = [&](const float* w) {
                for (int i = t8; i < XG * (XD / 4); i += 256) cp_async16(wst + (i >> 4) * GW_LD + (i & 15) * 16, w + (size_t)(i >> 4) * XD + (i & 15) * 4);
                cp_async_wait_all();
                named_bar_sync(2, 256);
            };
            auto load_store = [&](int grow, uint32_t c_f32, uint32_t c_rem, bool on) {    // `on` is warp-uniform
                if (!on) return;
                const uint8_t* wrow = wst + grow * GW_LD;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t f[32], rb[16];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 v = *reinterpret_cast<const uint4*>(wrow + 128 * h + 16 * c);
                        f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float r0 = trunc19_rem(__uint_as_float(f[2 * i])), r1 = trunc19_rem(__uint_as_float(f[2 * i + 1]));
                        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(rb[i]) : "f"(r1), "f"(r0));
                    }
                    tmem_st_32x32(tm + c_f32 + 32 * h, f);
                    tmem_st_32x16(tm + c_rem + 16 * h, rb);
                }
            };
            named_bar_sync(2, 256);              // both splitter groups have read their last feature stage
            stage_w(a.gru_w_ih);
            mbar_wait(acc_done, 0);              // conv accumulators (columns 0-255) and operand buffers (256-511) are dead
            tc_fence_after();
            if (tileA) load_store(L, C_GA, C_GA + 128, true);
            else load_store(128 + (L & 63), C_GB, C_GB + 64, q < 2);
            tmem_st_wait();
            named_bar_sync(2, 256);              // every row of W_ih has been read
            stage_w(a.gru_w_hh);
            if (tileA) load_store(L, C_GA + 64, C_GA + 160, true);
            else load_store(128 + (L & 63), C_GB, C_GB + 64, q >= 2);
            tmem_st_wait();
        }
        tc_fence_before();
        named_bar_sync(0, HT);                   // the phase-A barrier (bar.sync 0 from this branch, same barrier as below)
    }
    if (warp < 8) {
        tc_fence_before();
        named_bar_sync(0, HT);                   // end of phase A + MLP (the splitter warps arrive from their branch)
    }
    cluster_arrive_relaxed();   // no more multicasts / remote arrivals from this CTA; matched by cluster_wait() before exit
    tc_fence_after();
    PROF_END(phaseA); PROF_BEGIN(mlp);

    // =========================================== the loop ==========================================================
    // Thread = (TMEM lane prow = 32 * (warp % 4) + lane, column group cg = warp / 4): the four warps that can reach a TMEM
    // lane quarter split the 32 accumulator columns (slot rows) between them, so every element-wise stage runs on all 512
    // threads with at most 8 independent chains per thread.  Thread 0 issues the dots and update MMAs, lane 0 of warps
    // 0-3 the gate GEMMs; only lane 0 of a warp waits on an mbarrier (the rest of the warp parks at __syncwarp and does
    // not compete with the issuing lane for issue slots).
    const bool tcg = a.iters > 1;
    const int q4 = warp & 3, cg = warp >> 2, prow = q4 * 32 + lane;
    const uint32_t t_mine = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const bool o_tok = prow < R;                               // token row
    const bool o_ks = prow >= R && prow < R + G;               // ksum row of image prow - R
    const int o_g = prow < R ? prow / n : min(prow - R, G - 1);
    const int o_j = prow - o_g * n;                            // token index inside the image (token rows only)
    const bool o_live = o_tok && o_g < nimg;
    auto put_operand = [&](int o, int sr, int e, float v) {              // element (sr, e) of operand o into its three tiles
        uint8_t* base = gb + o * GB_OP;
        const int kb = e >> 5, c = e & 31;
        const int o32 = kb * (32 * 128) + sr * 128 + (((c >> 2) ^ (sr & 7)) << 4) + ((c & 3) << 2);
        *reinterpret_cast<float*>(base + o32) = v;
        *reinterpret_cast<float*>(base + GB_TILE32 + o32) = trunc19_rem(v);
        const int o16 = kb * (32 * 64) + sr * 64 + (((c >> 3) ^ ((sr >> 1) & 3)) << 4) + ((c & 7) << 1);
        *reinterpret_cast<__nv_bfloat16*>(base + 2 * GB_TILE32 + o16) = __float2bfloat16_rn(v);
    };
    auto sr_valid = [&](int sr) { return G == 2 ? (sr & 15) < S : sr < S; };    // G is a template parameter
    auto wait_bar = [&](uint64_t* bar, uint32_t parity) {
        if (lane == 0) mbar_wait(bar, parity);
        __syncwarp();
        tc_fence_after();
    };
    // GRU cell elements of this thread: element index k*HT + tid over (image, slot, feature); slot row of pair sl
    const int c_e = tid & 63, c_sl0 = tid >> 6, n_cell = (G * S * XD - tid + HT - 1) / HT;     // pairs sl = c_sl0 + 8 k, k < n_cell
    auto sl_row = [&](int sl) { return (G == 2 && sl >= S) ? SS + sl - S : sl; };
    // gate-row biases of this TMEM lane: r|z rows (lane = gate row), n rows (lanes 0-63 W_ih, 64-127 W_hh)
    float g_bi = 0.f, g_bh = 0.f, g_bn = 0.f;
    if (tcg) {
        g_bi = __ldg(a.gru_b_ih + prow);
        g_bh = __ldg(a.gru_b_hh + prow);
        g_bn = q4 < 2 ? __ldg(a.gru_b_ih + 128 + prow) : __ldg(a.gru_b_hh + 128 + prow - 64);
    }
    // Rows of absent slots in the operand / attention tiles only reach accumulator columns nobody reads; what must be
    // exact zeros is the K padding of the attention tiles (tokens n .. 8*ksteps-1 meet the zero padding of X^T: 0 * NaN)
    if (tid < 64) rs[tid] = 1.0f;
    {
        const uint32_t sr = (uint32_t)tid & 31u;
        const int jj = n + (tid >> 5);                       // 16 pad tokens at most, one per warp
        if (jj < 8 * ksteps) {
            const uint32_t o = (jj >> 5) * (32 * 128) + sr * 128 + (((((uint32_t)jj & 31) >> 2) ^ (sr & 7)) << 4) + ((jj & 3) << 2);
            *reinterpret_cast<float*>(at + o) = 0.f;
            *reinterpret_cast<float*>(at + a.at_form + o) = 0.f;
        }
        if (jj < 16 * ksteps16)
            *reinterpret_cast<__nv_bfloat16*>(atb + (jj >> 5) * (32 * 64) + sr * 64 + (((((uint32_t)jj & 31) >> 3) ^ ((sr >> 1) & 3)) << 4) +
                                              ((jj & 7) << 1)) = __float2bfloat16_rn(0.f);
    }
#pragma unroll 1
    for (int k = 0; k < n_cell; ++k) {
        const int sl = c_sl0 + 8 * k, sr = sl_row(sl);
        const float v = slots0[(sl >= S ? sl - S : sl) * XD + c_e];
        slots[sr * XD + c_e] = v;
        put_operand(1, sr, c_e, v);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    PROF_END(mlp); PROF_BEGIN(loop);
#ifdef SCOUTER_PROF
#define LT(ev) do { if (blockIdx.x == 0) g_trace_head[(70 + it * 2) * 8 + (ev)] = (unsigned long long)(clock64() - prof_begin_loop); } while (0)
#else
#define LT(ev)
#endif
    constexpr uint32_t idesc32 = idesc_tf32(128, 32), idesc32b = idesc_bf16(128, 32);
    const uint32_t gb_lo = desc_lo(smem_u32(gb));
    // column chunk of this thread in the dots accumulator: G = 2 -> 4 slots of its image (columns 16 g + 4 cg ..),
    // G = 1 -> 8 slots (columns 8 cg ..)
    constexpr int NV = G == 2 ? 4 : 8;
    const int i0 = NV * cg;
    const bool chunk_on = i0 < S;                            // warp-uniform: this warp's slot chunk exists
    for (int it = 0; it < a.iters; ++it) {
        const bool last = it == a.iters - 1;
        // ---- dots^T[token, slot row] = K . h  (compensated: K_t h_t + K_t h_r + K_r h); two issuing threads with their own
        //      accumulators (main product / corrections), summed by the readers in a fixed order ---------------------------
        if (lane == 0 && warp < 2) {
            const uint32_t k_t = tmem_base + C_KEY;
            const uint32_t b32 = gb_lo + (GB_OP >> 4), br32 = b32 + (GB_TILE32 >> 4), b16 = b32 + (2 * GB_TILE32 >> 4);
            tc_fence_after();
            if (tid == 0) LT(0);
            if (warp == 0) {
#pragma unroll
                for (uint32_t j = 0; j < 8; ++j)
                    umma_tf32_ts(tmem_base + C_DOT, k_t + 8 * j, desc_make(DESC_HI_SW128, b32 + (j >> 2) * (32 * 128 >> 4) + 2 * (j & 3)), idesc32, j != 0);
            } else {
#pragma unroll
                for (uint32_t j = 0; j < 8; ++j)
                    umma_tf32_ts(tmem_base + C_ACC_N_U, k_t + 8 * j, desc_make(DESC_HI_SW128, br32 + (j >> 2) * (32 * 128 >> 4) + 2 * (j & 3)), idesc32, j != 0);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    umma_bf16_ts(tmem_base + C_ACC_N_U, tmem_base + C_KEY_REM + 8 * j, desc_make(DESC_HI_SW64, b16 + (j >> 1) * (32 * 64 >> 4) + 2 * (j & 1)), idesc32b, 1);
            }
            umma_commit(dbar);
            if (tid == 0) LT(1);
        }
        wait_bar(dbar, (uint32_t)it & 1u);
        if (tid == 0) LT(2);
        // ---- attn = sigmoid(dots / r_i * t), r_i from the ksum lane, t = sum_i r_i (the d^-1/2 scale is a power of two:
        //      dots / r is unchanged by it and t carries it).  Thread = (token row, chunk of NV slots) ---------------------
        {
            float v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = 0.f;
            if (chunk_on) {
                if constexpr (G == 2) {        // two images: this row's image owns columns 16 g + ..
                    uint32_t m0[4], m1[4], c0[4], c1[4];
                    tmem_ld_32x4(t_mine + C_DOT + (uint32_t)i0, m0);
                    tmem_ld_32x4(t_mine + C_DOT + 16u + (uint32_t)i0, m1);
                    tmem_ld_32x4(t_mine + C_ACC_N_U + (uint32_t)i0, c0);
                    tmem_ld_32x4(t_mine + C_ACC_N_U + 16u + (uint32_t)i0, c1);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = __uint_as_float(o_g ? m1[k] : m0[k]) + __uint_as_float(o_g ? c1[k] : c0[k]);
                } else {
                    uint32_t m0[8], c0[8];
                    tmem_ld_32x8(t_mine + C_DOT + (uint32_t)i0, m0);
                    tmem_ld_32x8(t_mine + C_ACC_N_U + (uint32_t)i0, c0);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < NV; ++k) v[k] = __uint_as_float(m0[k]) + __uint_as_float(c0[k]);
                }
                if (o_ks) {
#pragma unroll
                    for (int k = 0; k < NV; ++k)
                        if (i0 + k < S) rs[o_g * 32 + i0 + k] = v[k];
                }
            }
            if (tid == 128) LT(3);
            tc_fence_before();
            __syncthreads();
            if (tid == 128) LT(4);
            if (chunk_on && o_tok) {
                const float* rp = rs + o_g * 32;
                float t_ = 0.f;
                for (int i = 0; i < S; ++i) t_ += rp[i];
                t_ *= 0.125f;
                float at_[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) at_[k] = sigm(v[k] / rp[i0 + k] * t_);      // rs is 1 beyond S: no NaN factory
                uint8_t* acol = at + (o_j >> 5) * (32 * 128) + ((o_j & 3) << 2);
                uint8_t* bcol = atb + (o_j >> 5) * (32 * 64) + ((o_j & 7) << 1);
                const uint32_t jc = (uint32_t)(o_j & 31) >> 2, jc8 = (uint32_t)(o_j & 31) >> 3;
                float* gout = (last && o_live && a.attn) ? a.attn + (size_t)(b0 + o_g) * S * n + o_j : nullptr;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const int i = i0 + k;
                    if (i < S) {
                        if (!last) {
                            const uint32_t sr = (uint32_t)(o_g * SS + i);
                            const float hi = to_tf32(at_[k]);
                            float* p = reinterpret_cast<float*>(acol + sr * 128 + ((jc ^ (sr & 7)) << 4));
                            *p = hi;
                            *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p) + a.at_form) = at_[k] - hi;
                            *reinterpret_cast<__nv_bfloat16*>(bcol + sr * 64 + ((jc8 ^ ((sr >> 1) & 3)) << 4)) = __float2bfloat16_rn(at_[k]);
                        } else {
                            attnP[prow * SPP + i] = at_[k];
                            if (gout) gout[(size_t)i * n] = at_[k];
                        }
                    }
                }
            }
            if (tid == 128) LT(8);
            if (!last) fence_proxy_async();
        }
        __syncthreads();
        if (last) break;
        // ---- u^T[64 image + feature, slot row] = X^T . attn (hi, then lo): one chain; image g reads only its own 16 columns,
        //      and both images see the same K order (position-independent arithmetic) -----------------------------------------
        if (lane == 0 && warp < 3) {                  // warp 0: X_t . attn_hi, warp 1: X_t . attn_lo, warp 2: X_r . attn (bf16); own accumulators
            tc_fence_after();
            if (tid == 0) LT(5);
            if (warp < 2) {
                const uint32_t d_t = tmem_base + (warp == 0 ? C_ACC_RZ_U : C_ACC_N_H);
                uint32_t ad = desc_lo(smem_u32(xt)), bd = desc_lo(smem_u32(at)) + (warp == 0 ? 0u : (uint32_t)a.at_form >> 4);
                for (int s = 0; s < ksteps; ++s) {
                    umma_tf32(d_t, desc_make(DESC_HI_SW128, ad), desc_make(DESC_HI_SW128, bd), idesc32, (uint32_t)s != 0u);
                    if ((s & 3) == 3) { ad += ((uint32_t)a.xt_kb >> 4) - 6; bd += ((32 * 128) >> 4) - 6; }
                    else { ad += 2; bd += 2; }
                }
            } else {
                uint32_t ad = desc_lo(smem_u32(xl)), bd = desc_lo(smem_u32(atb));
                for (int s = 0; s < ksteps16; ++s) {
                    umma_bf16(tmem_base + C_ACC_N_U, desc_make(DESC_HI_SW64, ad), desc_make(DESC_HI_SW64, bd), idesc32b, (uint32_t)s != 0u);
                    if (s & 1) { ad += ((uint32_t)a.xl_kb >> 4) - 2; bd += ((32 * 64) >> 4) - 2; }
                    else { ad += 2; bd += 2; }
                }
            }
            umma_commit(ubar);
            if (tid == 0) LT(6);
        }
        wait_bar(ubar, (uint32_t)it & 1u);
        if (tid == 0) LT(7);
        if constexpr (G == 2) {                       // lane = 64 image + feature: updates / d -> operand tiles of the gate GEMMs
            if (chunk_on) {
                uint32_t r[4], rl[4], rx[4];
                const int c0 = 16 * (q4 >> 1) + 4 * cg;
                tmem_ld_32x4(t_mine + C_ACC_RZ_U + (uint32_t)c0, r);
                tmem_ld_32x4(t_mine + C_ACC_N_H + (uint32_t)c0, rl);
                tmem_ld_32x4(t_mine + C_ACC_N_U + (uint32_t)c0, rx);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (4 * cg + k < S)
                        put_operand(0, c0 + k, prow & 63, ((__uint_as_float(r[k]) + __uint_as_float(rl[k])) + __uint_as_float(rx[k])) * (1.0f / XD));
                fence_proxy_async();
            }
        } else if (q4 < 2 && chunk_on) {
            uint32_t r[8], rl[8], rx[8];
            tmem_ld_32x8(t_mine + C_ACC_RZ_U + 8u * (uint32_t)cg, r);
            tmem_ld_32x8(t_mine + C_ACC_N_H + 8u * (uint32_t)cg, rl);
            tmem_ld_32x8(t_mine + C_ACC_N_U + 8u * (uint32_t)cg, rx);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (8 * cg + k < S)
                    put_operand(0, 8 * cg + k, prow, ((__uint_as_float(r[k]) + __uint_as_float(rl[k])) + __uint_as_float(rx[k])) * (1.0f / XD));
            fence_proxy_async();
        }
        if (tid == 128) LT(9);
        tc_fence_before();
        __syncthreads();
        // ---- gates^T = W . [u | h]: four accumulators, four issuing threads, 20 MMAs each ------------------------------
        if (warp < 4 && lane == 0) {
            const uint32_t o = (uint32_t)warp & 1u;                                    // operand: 0 = updates, 1 = slots
            const uint32_t d_t = tmem_base + (warp == 0 ? C_ACC_RZ_U : warp == 1 ? C_DOT : warp == 2 ? C_ACC_N_U : C_ACC_N_H);
            const uint32_t w_f = tmem_base + (warp < 2 ? C_GA + 64 * o : C_GB), w_r = tmem_base + (warp < 2 ? C_GA + 128 + 32 * o : C_GB + 64);
            const uint32_t b32 = gb_lo + o * (GB_OP >> 4), br32 = b32 + (GB_TILE32 >> 4), b16 = b32 + (2 * GB_TILE32 >> 4);
            tc_fence_after();
            if (tid == 0) LT(10);
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j)     // W_t x_t
                umma_tf32_ts(d_t, w_f + 8 * j, desc_make(DESC_HI_SW128, b32 + (j >> 2) * (32 * 128 >> 4) + 2 * (j & 3)), idesc32, j != 0);
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j)     // W_t x_r
                umma_tf32_ts(d_t, w_f + 8 * j, desc_make(DESC_HI_SW128, br32 + (j >> 2) * (32 * 128 >> 4) + 2 * (j & 3)), idesc32, 1);
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)     // W_r x (bf16)
                umma_bf16_ts(d_t, w_r + 8 * j, desc_make(DESC_HI_SW64, b16 + (j >> 1) * (32 * 64 >> 4) + 2 * (j & 1)), idesc32b, 1);
            umma_commit(gbar);
            if (tid == 0) LT(11);
        }
        wait_bar(gbar, (uint32_t)it & 1u);
        if (tid == 0) LT(12);
        {
            // gate read-out, slot rows 8 cg .. 8 cg + 7 of this lane's gate row: r|z -> sigmoid((gi + b_ih) + (gh + b_hh));
            // n rows: lanes 0-63 gi_n + b_ih, lanes 64-127 gh_n + b_hh
            uint32_t au[8], ah[8], av[8];
            tmem_ld_32x8(t_mine + C_ACC_RZ_U + 8u * (uint32_t)cg, au);
            tmem_ld_32x8(t_mine + C_DOT + 8u * (uint32_t)cg, ah);
            tmem_ld_32x8(t_mine + (q4 < 2 ? C_ACC_N_U : C_ACC_N_H) + 8u * (uint32_t)cg, av);
            tmem_ld_wait();
            float sg[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) sg[k] = sigm((__uint_as_float(au[k]) + g_bi) + (__uint_as_float(ah[k]) + g_bh));
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (sr_valid(8 * cg + k)) {
                    grz[(8 * cg + k) * 128 + prow] = sg[k];
                    gn[(8 * cg + k) * 128 + prow] = __uint_as_float(av[k]) + g_bn;
                }
            if (tid == 256) LT(13);
        }
        tc_fence_before();
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < n_cell; ++k) {            // GRU cell, gate order [r|z|n]
            const int sr = sl_row(c_sl0 + 8 * k);
            const float rg_ = grz[sr * 128 + c_e], zg = grz[sr * 128 + 64 + c_e];
            const float ng = tanh_fast(gn[sr * 128 + c_e] + rg_ * gn[sr * 128 + 64 + c_e]);
            const float hp = slots[sr * XD + c_e];
            const float hn = (hp - ng) * zg + ng;          // ATen's form of (1-z)*n + z*h
            slots[sr * XD + c_e] = hn;
            put_operand(1, sr, c_e, hn);
        }
        if (tid == 0) LT(14);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) LT(15);
    }

#ifdef SCOUTER_PROF
#define ET(slot) do { if (tid == 0) g_prof_head[blockIdx.x * 32 + (slot)] = (unsigned long long)(clock64() - prof_begin_loop); } while (0)
#else
#define ET(slot)
#endif
    ET(12);
    // ---- logits_i = loss_status / d * sum_j attn_ij * rowsum(X_j): one warp per (image, slot), fixed shuffle tree -------
    {   // half-warp per (image, slot): all pairs at once
        const int sl = tid >> 4, hl = lane & 15;
        const int g = (G == 2 && sl >= S) ? 1 : 0, i = sl - g * S;
        float s_ = 0.f;
        if (sl < G * S)
            for (int jj = hl; jj < n; jj += 16) s_ = fmaf(attnP[(g * n + jj) * SPP + i], xsum[g * n + jj], s_);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s_ += __shfl_xor_sync(0xffffffffu, s_, o);
        if (hl == 0 && sl < G * S) usum[sl] = s_ * (1.0f / XD);
    }
    ET(13);
    __syncthreads();
    ET(14);
    for (int idx = tid; idx < nimg * a.C; idx += HT) {
        const int img = idx / a.C, c = idx - img * a.C;
        float s = 0.f;
        for (int m = 0; m < a.spc; ++m) s += usum[img * S + c * a.spc + m];
        a.logits[(size_t)(b0 + img) * a.C + c] = (float)a.loss_status * s;
    }
    if (a.attn_sum) {
        for (int img = warp; img < nimg; img += HW_) {        // per-image sum of the final attention (area loss), fixed order
            float s = 0.f;
            for (int j = lane; j < n; j += 32)
                for (int i = 0; i < S; ++i) s += attnP[(img * n + j) * SPP + i];
            s = warp_sum(s);
            if (lane == 0) a.attn_sum[b0 + img] = s;
        }
    }
    ET(15);
    PROF_END(loop);
    if (tid == 0) { PROF_STORE(g_prof_head, 0, phaseA); PROF_STORE(g_prof_head, 1, mlp); PROF_STORE(g_prof_head, 2, loop); }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    cluster_wait();       // the peer may still be arriving on this CTA's barriers until it has left phase A
#ifdef SCOUTER_PROF
    if (tid == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_prof_head[blockIdx.x * 32 + 29] = gt;
        g_prof_head[blockIdx.x * 32 + 27] = (unsigned long long)(clock64() - prof_entry);
    }
#endif
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
//   phase A + to_k MLP:  [NA feature stages][NW weight slots]
//   from the hand-over:  [X^T tiles + 8 KB slack][xsum, ksum, rs, usum] at the bottom of the (dead) feature ring, then the
//                        staging of one GRU matrix; the plain keys for ksum at the bottom of the (dead) weight ring
//   loop:                GRU / dots operand tiles, attention tiles, gate exchange, plain attention, slots -- from the staging
//                        area upwards (staging, plain keys and the weight ring are dead by then)
size_t layout(int G, int n, int S, int L, FusedArgs* out) {
    const int R = G * n;
    const int R8 = (R + 7) & ~7;
    const int kbt = (n + 31) / 32, SPP = G == 2 ? 17 : 33;
    const size_t a_stage = (size_t)R8 * 128;
    const size_t xt_kb = (size_t)G * 64 * 128;                               // k-block of the X^T tile: (64 G) rows x 32 tokens
    const size_t off_xt = 0, off_xl = off_xt + kbt * xt_kb + (G == 1 ? 8192 : 0);     // G = 1: the M = 128 MMAs read 64 rows past the tile
    const size_t xl_kb = xt_kb / 2;
    const size_t off_misc = off_xl + kbt * xl_kb + (G == 1 ? 4096 : 0);
    const size_t off_stage = align_up(off_misc + 2048, 1024);
    const size_t wst_bytes = (size_t)XG * GW_LD;
    const size_t at_form = (size_t)kbt * 32 * 128;
    // loop regions
    const size_t off_gb = off_stage;
    const size_t off_at = off_gb + 2 * GB_OP;                                 // 1024-aligned: GB_OP = 20480
    const size_t off_atb = off_at + 2 * at_form;                              // 1024-aligned (at_form = kbt * 4096)
    const size_t off_grz = align_up(off_atb + (size_t)kbt * 32 * 64, 1024);
    const size_t off_gn = off_grz + 32 * 128 * 4;
    const size_t off_attnp = off_gn + 32 * 128 * 4;
    const size_t off_slots = off_attnp + align_up((size_t)R * SPP * 4, 16);
    const size_t loop_end = off_slots + 32 * XD * 4;
    static int nw_env = [] { const char* e = getenv("SCOUTER_HEAD_NW"); int v = e ? atoi(e) : 6; return v < 2 ? 2 : (v > MAX_NW ? MAX_NW : v); }();
    // the weight tile comes from L2 through a multicast TMA whose latency is ~2-3 k-blocks: prefer a deep weight ring, then
    // as many feature stages as fit
    static int na_env = [] { const char* e = getenv("SCOUTER_HEAD_NA"); int v = e ? atoi(e) : MAX_NA; return v < 2 ? 2 : (v > MAX_NA ? MAX_NA : v); }();
    for (int nw = nw_env; nw >= 2; --nw)
        for (int na = na_env; na >= 3; --na) {
            const size_t off_w = align_up(std::max((size_t)na * a_stage, off_stage), 1024);   // X^T / sums are written while the weight ring feeds the MLP
            const size_t ring_end = off_w + (size_t)nw * W_SLOT;
            // staging of one GRU matrix: inside the feature ring when it fits below the weight ring, else behind the rings
            const size_t off_wst = off_stage + wst_bytes <= off_w ? off_stage : align_up(ring_end, 1024);
            // plain keys: inside the weight ring (dead when they are written) if they fit, else behind everything
            const size_t kp_bytes = (size_t)R * LDX * 4;
            const size_t after = std::max(ring_end, off_wst + wst_bytes);
            const size_t off_kp = kp_bytes <= (size_t)nw * W_SLOT ? off_w : align_up(after, 16);
            const size_t body = std::max(std::max(after, loop_end), off_kp + kp_bytes);
            const size_t off_bar = align_up(body, 16);
            const size_t total = off_bar + 512 + (size_t)(L + S) * XD * 4 + 1024;   // barriers, to_k biases, initial slots, slack
            if (total <= 227 * 1024) {
                if (out) {
                    out->na = na; out->nw = nw; out->a_stage = (int)a_stage; out->off_w = (int)off_w; out->off_bar = (int)off_bar;
                    out->off_xt = (int)off_xt; out->xt_kb = (int)xt_kb; out->off_misc = (int)off_misc;
                    out->off_xl = (int)off_xl; out->xl_kb = (int)xl_kb; out->off_atb = (int)off_atb;
                    out->off_wst = (int)off_wst; out->off_kp = (int)off_kp;
                    out->off_gb = (int)off_gb; out->off_at = (int)off_at; out->at_form = (int)at_form;
                    out->off_grz = (int)off_grz; out->off_gn = (int)off_gn; out->off_attnp = (int)off_attnp; out->off_slots = (int)off_slots;
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

// Images per unit: two when both fit the 128-row tile together with their ksum rows and share the 32 slot rows of the
// gate GEMMs (16 each), else one.
static int pick_group(int batch, int n, int S, int L) {
    for (int G = std::min(2, batch); G >= 1; --G) {
        if (G * n + G > 128 || S > 32 / G) continue;
        if (layout(G, n, S, L, nullptr) != 0) return G;
    }
    return 0;
}

bool head_fused_supported(const scouter_xslot_desc_t* d, int batch, int n, int channel) {
    static bool off = getenv("SCOUTER_NO_FUSED_HEAD") != nullptr;
    const int S = d->num_classes * d->slots_per_class;
    if (off || n > 127 || S > 32 || channel % 32 || d->to_k_layers < 1 || d->iters < 1) return false;
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
    a.kblocks = io->channel / 32;
    const size_t smem = layout(a.G, n, a.S, a.L, &a);
    SC_CHECK_ARG(smem, SCOUTER_E_UNSUPPORTED, "head_fused: shared-memory layout does not fit");
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
    const int units = cdiv(io->batch, a.G);
    if (a.G == 2) {
        SC_CUDA(cudaFuncSetAttribute(head_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        head_fused_kernel<2><<<2 * cdiv(units, 2), HT, smem, s>>>(tmA, tmB, tmB2, tmT, tmT2, a);   // clusters of two (padding CTA when odd)
    } else {
        SC_CUDA(cudaFuncSetAttribute(head_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        head_fused_kernel<1><<<2 * cdiv(units, 2), HT, smem, s>>>(tmA, tmB, tmB2, tmT, tmT2, a);
    }
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
