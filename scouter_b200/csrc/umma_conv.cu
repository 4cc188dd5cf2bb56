// tcgen05 + TMA implicit-GEMM convolution for sm_100a (SCOUTER_MATH_TC).
//
//   D[M = B*H*W, N = Cout/groups] = A[M, K = kh*kw*Cin/groups] * W[N, K]^T     (stride 1, 1x1 or 3x3 pad 1)
//
// * Two precisions of the same kernel (template SPLIT):
//     SPLIT=true  (SCOUTER_MATH_TC, default): error-compensated 16-bit product (ptx.cuh split2_act / split2_wgt).  x = h + r
//       with h = fp16(x) (rounded, 11-bit significand, saturating) and r = x - h, exact in fp32 and ~2^-12 |x|.  The issuer
//       accumulates, in the fp32 TMEM accumulator,
//           A_b*W_r (bf16 x bf16)  +  A_r*W_h  +  A_h*W_h (fp16 x fp16)         2 + 2 + 2 kind::f16 MMAs (K = 16) per k-block
//       with A_b = bf16(a), A_r = fp16(a - a_h), W_r = bf16(w - w_h).  Four splitter warps derive A_h / A_b / A_r from the
//       TMA-written fp32 tile; weights come pre-split from the host ([fp16 W_h ; bf16 W_r]) or are split in the kernel.
//       Operands stay plain fp32 in HBM.  Round 1 ran the main product as kind::tf32 on trunc19(x) (4 MMAs of K = 8 + 2 + 2
//       bf16 corrections): 8 MMA times per k-block against 6 now.
//     SPLIT=false (SCOUTER_MATH_TC_FAST): one tf32 MMA; producers round stored activations with cvt.rna and
//       weights are pre-rounded on the host, so the MMA multiplies exactly the stored values (cuDNN-TF32 class).
// * No im2col buffer: for a 3x3 conv the K loop walks the 9 taps and each tap is ONE 4-D TMA box load
//   {32 channels, Wb, Hb, Nb} of the input shifted by (r-1, s-1); TMA's out-of-bound zero fill is the padding.
//   A 1x1 conv uses a flat 2-D map {32 channels, 128 rows}.  Both land as 128-byte rows in SWIZZLE_128B
//   layout, i.e. a K-major UMMA operand tile of 128 rows x 32 tf32.
// * Persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected
//   thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers -> bias/residual/ReLU -> global),
//   warps 8-11 = operand splitters (SPLIT only).  smem ring of STAGES {A 16 KB, W [, the 16-bit tiles]} (TS: see Cfg); two TMEM
//   accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Reference ops replaced: the nn.Conv2d+BatchNorm2d(+ReLU)(+residual) chains of timm/models/resnest.py:111-143,
// split_attn.py:43-45,56-60 and resnet.py:403-408 (eval mode, BN folded).
#include "ptx.cuh"
#include "umma.cuh"

#pragma nv_diag_suppress 177   // Cfg members that only some instantiations reference

namespace scouter {
using namespace ptx;

namespace {

struct UmmaArgs {
    const float* bias;
    const float* res;
    float* out;
    int mode;  // 0 = flat rows (1x1), 1 = spatial boxes
    int M;     // flat: B*H*W
    int B, H, W;
    int Wb, Hb, Nb, tw, th;
    int m_tiles, n_tiles, groups;
    int cin_g, cout_g, Cout;
    int kw, pad, kblocks, cblocks;
    int relu, round_out;
    int a_bytes;
    int chunk;  // k-blocks per accumulation chunk
    int tma_store;  // flat mode: epilogue stages 128x16 blocks in shared memory and writes them with TMA stores
    int rem_rows;  // SPLIT: > 0 = weights are pre-split, remainder rows start at this row of the weight map
    int ksplit;    // > 1: split-K; unit (tile, sp) covers k-blocks [sp*kblocks, (sp+1)*kblocks) of the full K and writes raw
                   // partial sums to out + sp*slab (no bias / ReLU); `kblocks` is then the per-split count
    long long slab;  // floats between partial slabs
};

// TS: the activation operands live in TENSOR MEMORY.  In SS mode every MMA reads its whole A slice from shared memory
// (M128 x N128 x K8: 4 KB + 4 KB = 64 clk at 128 B/clk, exactly its tensor-pipe floor), so together with the TMA writes and
// the split tiles a BN=128 k-block moves 160 KB through the shared-memory pipe: 1250 clk for 512 clk of MMAs (measured:
// 1200 clk per k-block on the layer-4 1x1 convs).  With TS the splitter threads (thread = row) read their row of the TMA
// tile once, derive the three 16-bit forms in registers and tcgen05.st [fp16 | bf16 | bf16 remainder] into the 48-column TMEM
// operand buffer of the stage; the MMAs read only the weight tiles from shared memory.
template <int BN, bool SPLIT, bool RES = false, bool TS = false>
struct Cfg {
    static_assert(!TS || SPLIT, "TS is a variant of the error-compensated kernel");
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int RAW = A_BYTES + B_BYTES;            // what TMA writes per stage
    // SPLIT: [A fp32 | W fp32 (only when split in the kernel) | fp16 A_h | bf16 A_b | fp16 A_r | fp16 W_h | bf16 W_r]
    static constexpr int A_TILE = A_BYTES / 2, B_TILE = B_BYTES / 2;   // 16-bit tiles: 64-byte rows
    static constexpr int STAGE = SPLIT ? RAW + 3 * A_TILE + 2 * B_TILE : RAW;
    static constexpr int OFF_A16 = RAW, OFF_W16 = RAW + 3 * A_TILE;
    static constexpr int OP_COL0 = 2 * BN;                   // TS: operand buffer b = 48 columns at OP_COL0 + 48*b
    static constexpr int OP_COLS = 48;                       //     [fp16 A_h | bf16 A_b | fp16 A_r], 16 columns each

    static constexpr int TMEM_COLS = TS ? 512 : (2 * BN < 32 ? 32 : 2 * BN);
    // SPLIT: two epilogue warpgroups (warps 4-7 and 12-15) share the columns of a BN=128 tile so that the running
    // fp32 sums of the chunked accumulation fit in registers (64 per thread).
    static constexpr int EPI_GROUPS = (SPLIT && BN == 128) ? 2 : 1;
    static constexpr int NC = BN / EPI_GROUPS;               // columns per epilogue thread
    static constexpr int THREADS = SPLIT ? 512 : 256;
    static constexpr int OUT_STAGE = 128 * 64;              // TMA-store staging: 128 rows x 16 fp32, SWIZZLE_64B
    // RES: the whole 128 x BN residual tile is TMA-loaded into BN/16 slabs; the epilogue adds it in place and the
    // same slabs are the source of the TMA stores (no staging double buffer, no row-per-thread residual loads).
    // SPLIT: one slab per 16 output columns of the tile (RES: they first receive the residual), so a tile leaves with ONE proxy
    // fence + barrier per epilogue group; the single-pass kernel keeps the two-buffer ping-pong.
    static constexpr int OUT_BYTES = SPLIT ? (BN / 16) * OUT_STAGE : EPI_GROUPS * 2 * OUT_STAGE;
    static constexpr int FIT = (224 * 1024 - OUT_BYTES) / STAGE;
    static constexpr int STAGES = TS ? 1 : (FIT > 8 ? 8 : FIT);
    // TS: two rings instead of stages.  The activation tile is dead as soon as the splitters hold it in registers, so its
    // ring (NA x 16 KB, fed from HBM) runs far ahead; the weight tiles (fp32 + bf16 pair, from L2) need NW slots; NO TMEM
    // operand buffers sit between the splitters and the MMAs.
    static constexpr int W_SLOT = 2 * B_TILE;                // [fp16 W_h | bf16 W_r], BN x 64 bytes each
    static constexpr int BUDGET = 224 * 1024 - OUT_BYTES;
    static constexpr int NW = (BUDGET - 4 * W_SLOT) / A_BYTES >= 5 ? 4 : 3;
    static constexpr int NA_FIT = (BUDGET - NW * W_SLOT) / A_BYTES;
    static constexpr int NA = NA_FIT > 8 ? 8 : NA_FIT;
    static constexpr int NO = 4;
    static constexpr int RING_BYTES = TS ? NA * A_BYTES + NW * W_SLOT : STAGES * STAGE;
    static constexpr int SMEM = RING_BYTES + 1024 /*align slack*/ + 1024 /*barriers*/ + OUT_BYTES;
};

#ifdef SCOUTER_PROF
__device__ unsigned long long g_prof_flat[256 * 32];
#endif

template <int BN, bool SPLIT, bool RES, bool TS>
__global__ void __launch_bounds__(Cfg<BN, SPLIT, RES, TS>::THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ CUtensorMap tmR, const UmmaArgs p) {
    using C = Cfg<BN, SPLIT, RES, TS>;
    static_assert(!RES || SPLIT, "the in-place residual epilogue exists for the SPLIT kernel only");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* cfull = empty + C::STAGES;   // [2] accumulator chunk complete (tcgen05.commit)
    uint64_t* cempty = cfull + 2;          // [2] accumulator chunk drained by every epilogue thread
    uint64_t* split_done = cempty + 2;     // [STAGES], SPLIT only: remainders written, stage ready for the issuer
    uint64_t* rfull = split_done + C::STAGES;   // RES: residual tile landed (TMA)
    uint64_t* rempty = rfull + 1;               // RES: every store of the tile has read its slab (one arrival per epilogue group)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rempty + 1);
    uint8_t* out_stage = smem + C::RING_BYTES + 1024;   // [EPI_GROUPS][2][OUT_STAGE], 1024-aligned
    // TS rings
    uint64_t* fullA = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES + 512);   // [8] activation tile landed
    uint64_t* emptyA = fullA + 8;           // [8] tile read into registers by the four splitter warps
    uint64_t* fullW = emptyA + 8;           // [4] weight slot landed
    uint64_t* done = fullW + 4;             // [4] MMAs of k-block (index gk & 3) retired: frees weight slot and operand buffer
    uint64_t* opfull = done + 4;            // [4] operands of the k-block are in TMEM
    uint8_t* const w_ring = smem + C::NA * C::A_BYTES;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            mbar_init(&split_done[i], 4);   // one arrival per splitter warp
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&cfull[i], 1);
            mbar_init(&cempty[i], 128 * C::EPI_GROUPS);
        }
        mbar_init(rfull, 1);
        mbar_init(rempty, C::EPI_GROUPS);
        if constexpr (TS) {
            for (int i = 0; i < 8; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 4); }
            for (int i = 0; i < 4; ++i) { mbar_init(&fullW[i], 1); mbar_init(&done[i], 1); mbar_init(&opfull[i], 4); }
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int total = p.m_tiles * p.n_tiles * p.groups * p.ksplit;
    // The K loop of a tile is cut into chunks of p.chunk k-blocks; each chunk accumulates in its own TMEM buffer
    // (the two buffers alternate across chunks AND tiles) and the epilogue threads merge the chunks in fp32
    // round-to-nearest registers.  The TMEM accumulator truncates on every MMA, so this bounds the truncation
    // chain to chunk*4*(SPLIT?3:1) steps; with one chunk per tile it degenerates to plain double buffering.
    const int nchunks = (p.kblocks + p.chunk - 1) / p.chunk;

    if (TS && warp == 0) {
        if (elect_one()) {
            // ===== activation producer (TS): runs NA k-blocks ahead; also fetches the residual tile (RES) =====
            int sa_i = 0;
            uint32_t aphase = 0;
            int gk = 0;
            uint32_t rphase = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int sp = t % p.ksplit;
                const int tt = t / p.ksplit;
                const int nt = tt % p.n_tiles;
                const int mt = (tt / p.n_tiles) % p.m_tiles;
                const int g = tt / (p.n_tiles * p.m_tiles);
                int w0 = 0, h0 = 0, b0 = 0;
                if (p.mode) {
                    w0 = (mt % p.tw) * p.Wb;
                    h0 = ((mt / p.tw) % p.th) * p.Hb;
                    b0 = (mt / (p.tw * p.th)) * p.Nb;
                }
                bool res_pending = RES;
                auto issue_residual = [&]() {
                    mbar_arrive_expect_tx(rfull, (uint32_t)(128 * BN * 4));
#pragma unroll
                    for (int j = 0; j < BN / 16; ++j)
                        tma_load_2d(out_stage + j * C::OUT_STAGE, &tmR, rfull, g * p.cout_g + nt * BN + j * 16, mt * 128);
                    res_pending = false;
                    rphase ^= 1;
                };
                for (int kb = 0; kb < p.kblocks; ++kb, ++gk) {
                    if (RES && res_pending && mbar_try_wait(rempty, rphase ^ 1)) issue_residual();
                    if (gk >= C::NA) mbar_wait(&emptyA[sa_i], aphase ^ 1);
                    uint8_t* sa = smem + sa_i * C::A_BYTES;
                    mbar_arrive_expect_tx(&fullA[sa_i], (uint32_t)p.a_bytes);
                    const int kbg = sp * p.kblocks + kb;
                    const int tap = kbg / p.cblocks;
                    const int c0 = p.cin_g * g + (kbg - tap * p.cblocks) * 32;
                    if (p.mode) {
                        const int r = tap / p.kw, s_ = tap - r * p.kw;
                        tma_load_4d(sa, &tmA, &fullA[sa_i], c0, w0 + s_ - p.pad, h0 + r - p.pad, b0);
                    } else {
                        tma_load_2d(sa, &tmA, &fullA[sa_i], c0, mt * 128);
                    }
                    if (++sa_i == C::NA) { sa_i = 0; aphase ^= 1; }
                }
                if (RES && res_pending) {
                    mbar_wait(rempty, rphase ^ 1);
                    issue_residual();
                }
            }
        }
    } else if (TS && warp == 3) {
        if (elect_one()) {
            // ===== weight producer (TS): pre-split [fp16 W_h ; bf16 W_r] tiles, NW slots =====
            int ws = 0, gk = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int sp = t % p.ksplit;
                const int tt = t / p.ksplit;
                const int nt = tt % p.n_tiles;
                const int g = tt / (p.n_tiles * p.m_tiles);
                for (int kb = 0; kb < p.kblocks; ++kb, ++gk) {
                    if (gk >= C::NW) {
                        const int j = gk - C::NW;
                        mbar_wait(&done[j & 3], (uint32_t)(j >> 2) & 1u);
                    }
                    uint8_t* sw = w_ring + ws * C::W_SLOT;
                    mbar_arrive_expect_tx(&fullW[ws], (uint32_t)C::W_SLOT);
                    const int kbg = sp * p.kblocks + kb;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        tma_load_2d(sw + j * C::B_TILE, &tmB2, &fullW[ws], kbg * 32, j * p.rem_rows + g * p.cout_g + nt * BN);
                    if (++ws == C::NW) ws = 0;
                }
            }
        }
    } else if (TS && warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer (TS): A operands from the TMEM operand buffers, weights from the weight ring =====
            constexpr uint32_t id_h = idesc_f16(128, BN), id_b = idesc_bf16(128, BN);
            const uint32_t w_lo0 = desc_lo(smem_u32(w_ring));
            const uint32_t opfull_a = smem_u32(opfull), fullw_a = smem_u32(fullW), done_a = smem_u32(done), cfull_a = smem_u32(cfull),
                           cempty_a = smem_u32(cempty);
            const int kblocks = p.kblocks, chunk = p.chunk;
            int ws = 0, in_chunk = 0;
            uint32_t wphase = 0, gk = 0, cc = 0;
            PROF_DECL(cempty); PROF_DECL(full); PROF_DECL(opf); PROF_DECL(iss); PROF_DECL(nkb); PROF_BEGIN(iss);
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
#ifdef SCOUTER_PROF
                prof_nkb += kblocks;
#endif
                for (int kb = 0; kb < kblocks; ++kb, ++gk) {
                    const uint32_t buf = cc & 1, ob = gk & 3;
                    if (in_chunk == 0) PROF_T(cempty, mbar_wait_a(cempty_a + 8 * buf, ((cc >> 1) & 1) ^ 1));
                    PROF_T(opf, mbar_wait_a(opfull_a + 8 * ob, (gk >> 2) & 1));
                    PROF_T(full, mbar_wait_a(fullw_a + 8 * ws, wphase));
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * BN;
                    const uint32_t acc = in_chunk != 0;
                    const uint32_t a_tm = tmem_base + C::OP_COL0 + (uint32_t)C::OP_COLS * ob;   // [fp16 A_h | bf16 A_b | fp16 A_r]
                    const uint32_t w_lo = w_lo0 + (uint32_t)ws * (C::W_SLOT >> 4);              // [fp16 W_h | bf16 W_r]
#pragma unroll
                    for (uint32_t k = 0; k < 2; ++k)   // A_b * W_r   (bf16)
                        umma_bf16_ts(d_tmem, a_tm + 16 + 8 * k, desc_make(DESC_HI_SW64, w_lo + (C::B_TILE >> 4) + 2 * k), id_b, acc | k);
#pragma unroll
                    for (uint32_t k = 0; k < 2; ++k)   // A_r * W_h   (fp16)
                        umma_bf16_ts(d_tmem, a_tm + 32 + 8 * k, desc_make(DESC_HI_SW64, w_lo + 2 * k), id_h, 1);
#pragma unroll
                    for (uint32_t k = 0; k < 2; ++k)   // A_h * W_h   (fp16)
                        umma_bf16_ts(d_tmem, a_tm + 8 * k, desc_make(DESC_HI_SW64, w_lo + 2 * k), id_h, 1);
                    umma_commit_a(done_a + 8 * ob);    // frees the weight slot and the operand buffer
                    if (++ws == C::NW) { ws = 0; wphase ^= 1; }
                    if (++in_chunk == chunk || kb + 1 == kblocks) {
                        umma_commit_a(cfull_a + 8 * buf);
                        ++cc;
                        in_chunk = 0;
                    }
                }
            }
            PROF_END(iss);
            PROF_STORE(g_prof_flat, 4, iss); PROF_STORE(g_prof_flat, 6, cempty); PROF_STORE(g_prof_flat, 7, full);
            PROF_STORE(g_prof_flat, 5, opf); PROF_STORE(g_prof_flat, 8, nkb);
        }
    } else if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            PROF_DECL(empty); PROF_DECL(prod); PROF_BEGIN(prod);
            uint32_t rphase = 0;   // RES: parity of the residual buffer cycle
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int sp = t % p.ksplit;
                const int tt = t / p.ksplit;
                const int nt = tt % p.n_tiles;
                const int mt = (tt / p.n_tiles) % p.m_tiles;
                const int g = tt / (p.n_tiles * p.m_tiles);
                int w0 = 0, h0 = 0, b0 = 0;
                if (p.mode) {
                    w0 = (mt % p.tw) * p.Wb;
                    h0 = ((mt / p.tw) % p.th) * p.Hb;
                    b0 = (mt / (p.tw * p.th)) * p.Nb;
                }
                // RES: the residual tile goes out as soon as the previous tile's stores have drained the slabs -- polled
                // between k-blocks so that a busy epilogue never holds back the operand loads of this tile.
                bool res_pending = RES;
                auto issue_residual = [&]() {
                    mbar_arrive_expect_tx(rfull, (uint32_t)(128 * BN * 4));
#pragma unroll
                    for (int j = 0; j < BN / 16; ++j)
                        tma_load_2d(out_stage + j * C::OUT_STAGE, &tmR, rfull, g * p.cout_g + nt * BN + j * 16, mt * 128);
                    res_pending = false;
                    rphase ^= 1;
                };
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    if (RES && res_pending && mbar_try_wait(rempty, rphase ^ 1)) issue_residual();
                    PROF_T(empty, mbar_wait(&empty[stage], phase ^ 1));
                    uint8_t* sa = smem + stage * C::STAGE;
                    uint8_t* sb = sa + C::A_BYTES;
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)(p.a_bytes + C::B_BYTES));   // fp32 W, or the two 16-bit tiles
                    const int kbg = sp * p.kblocks + kb;   // k-block index in the full K
                    const int tap = kbg / p.cblocks;
                    const int c0 = p.cin_g * g + (kbg - tap * p.cblocks) * 32;
                    if (p.mode) {
                        const int r = tap / p.kw, s = tap - r * p.kw;
                        tma_load_4d(sa, &tmA, &full[stage], c0, w0 + s - p.pad, h0 + r - p.pad, b0);
                    } else {
                        tma_load_2d(sa, &tmA, &full[stage], c0, mt * 128);
                    }
                    if (SPLIT && p.rem_rows) {  // host-pre-split weights: fp16 W_h and bf16 W_r tiles
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            tma_load_2d(sa + C::OFF_W16 + j * C::B_TILE, &tmB2, &full[stage], kbg * 32, j * p.rem_rows + g * p.cout_g + nt * BN);
                    } else {
                        tma_load_2d(sb, &tmB, &full[stage], kbg * 32, g * p.cout_g + nt * BN);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                if (RES && res_pending) {
                    mbar_wait(rempty, rphase ^ 1);
                    issue_residual();
                }
            }
            PROF_END(prod);
            PROF_STORE(g_prof_flat, 0, prod); PROF_STORE(g_prof_flat, 1, empty);
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            // One thread, ~4.5 clk per dependent instruction (profiles/r01_role_stalls_*.txt): descriptors are a constant
            // high word plus a low word advanced by one 32-bit add per stage, barriers are 32-bit shared addresses.
            constexpr uint32_t idesc = idesc_tf32(128, BN);
            constexpr uint32_t SSTEP = C::STAGE >> 4;
            const uint32_t s_lo0 = desc_lo(smem_u32(smem));
            const uint32_t ready_a = smem_u32(SPLIT ? split_done : full), empty_a = smem_u32(empty), cfull_a = smem_u32(cfull),
                           cempty_a = smem_u32(cempty);
            const int kblocks = p.kblocks, chunk = p.chunk;
            int stage = 0, in_chunk = 0;
            uint32_t phase = 0, s_lo = s_lo0;
            uint32_t cc = 0;  // running chunk counter
            PROF_DECL(cempty); PROF_DECL(full); PROF_DECL(iss); PROF_DECL(nkb); PROF_BEGIN(iss);
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
#ifdef SCOUTER_PROF
                prof_nkb += p.kblocks;
#endif
                for (int kb = 0; kb < kblocks; ++kb) {
                    const uint32_t buf = cc & 1;
                    if (in_chunk == 0) PROF_T(cempty, mbar_wait_a(cempty_a + 8 * buf, ((cc >> 1) & 1) ^ 1));
                    PROF_T(full, mbar_wait_a(ready_a + 8 * stage, phase));
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * BN;
                    const uint32_t acc = in_chunk != 0;
                    // UMMA_K = 8 tf32 / 16 halves = 32 bytes: advance the start address inside the swizzle atom (+2 x 16 B)
                    if constexpr (SPLIT) {
                        constexpr uint32_t id_h = idesc_f16(128, BN), id_b = idesc_bf16(128, BN);
                        constexpr uint32_t AH = C::OFF_A16 >> 4, AB = (C::OFF_A16 + C::A_TILE) >> 4, AR = (C::OFF_A16 + 2 * C::A_TILE) >> 4;
                        constexpr uint32_t WH = C::OFF_W16 >> 4, WR = (C::OFF_W16 + C::B_TILE) >> 4;
#pragma unroll
                        for (uint32_t k = 0; k < 2; ++k)   // A_b * W_r   (bf16)
                            umma_bf16(d_tmem, desc_make(DESC_HI_SW64, s_lo + AB + 2 * k), desc_make(DESC_HI_SW64, s_lo + WR + 2 * k), id_b, acc | k);
#pragma unroll
                        for (uint32_t k = 0; k < 2; ++k)   // A_r * W_h   (fp16)
                            umma_bf16(d_tmem, desc_make(DESC_HI_SW64, s_lo + AR + 2 * k), desc_make(DESC_HI_SW64, s_lo + WH + 2 * k), id_h, 1);
#pragma unroll
                        for (uint32_t k = 0; k < 2; ++k)   // A_h * W_h   (fp16)
                            umma_bf16(d_tmem, desc_make(DESC_HI_SW64, s_lo + AH + 2 * k), desc_make(DESC_HI_SW64, s_lo + WH + 2 * k), id_h, 1);
                    } else {
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k)
                            umma_tf32(d_tmem, desc_make(DESC_HI_SW128, s_lo + 2 * k),
                                      desc_make(DESC_HI_SW128, s_lo + (C::A_BYTES >> 4) + 2 * k), idesc, acc | k);
                    }
                    umma_commit_a(empty_a + 8 * stage);  // frees the smem slot when these MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; s_lo = s_lo0; } else { s_lo += SSTEP; }
                    if (++in_chunk == chunk || kb + 1 == kblocks) {
                        umma_commit_a(cfull_a + 8 * buf);
                        ++cc;
                        in_chunk = 0;
                    }
                }
            }
            PROF_END(iss);
            PROF_STORE(g_prof_flat, 4, iss); PROF_STORE(g_prof_flat, 6, cempty); PROF_STORE(g_prof_flat, 7, full);
            PROF_STORE(g_prof_flat, 8, nkb);
        }
    } else if ((warp >= 4 && warp < 8) || (C::EPI_GROUPS == 2 && warp >= 12)) {
        // ===== epilogue: thread = one output row (32 TMEM lanes per warp) x NC columns =====
        const int q = warp & 3;
        const int grp = warp >= 12 ? 1 : 0;
        const int row = q * 32 + lane;
        const int col0 = grp * C::NC;
        uint32_t cc = 0;
        uint32_t rphase = 0;
        PROF_DECL(cfull); PROF_DECL(store); PROF_DECL(merge); PROF_DECL(epi); PROF_DECL(e_wait); PROF_DECL(e_bar1); PROF_DECL(e_math);
        PROF_DECL(e_fence); PROF_DECL(e_bar2); PROF_DECL(e_tma); PROF_BEGIN(epi);
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const int sp = t % p.ksplit;
            const int tt = t / p.ksplit;
            const int nt = tt % p.n_tiles;
            const int mt = (tt / p.n_tiles) % p.m_tiles;
            const int g = tt / (p.n_tiles * p.m_tiles);
            bool valid;
            long long orow;
            if (p.mode) {
                const int hw = p.Hb * p.Wb;
                const int nb = row / hw, rem = row - nb * hw;
                const int hb = rem / p.Wb, wb = rem - hb * p.Wb;
                const int b = (mt / (p.tw * p.th)) * p.Nb + nb;
                const int h = ((mt / p.tw) % p.th) * p.Hb + hb;
                const int w = (mt % p.tw) * p.Wb + wb;
                valid = nb < p.Nb && b < p.B && h < p.H && w < p.W;
                orow = ((long long)b * p.H + h) * p.W + w;
            } else {
                orow = (long long)mt * 128 + row;
                valid = orow < p.M;
            }
            const int ch0 = g * p.cout_g + nt * BN + col0;
            float* op = p.out + (long long)sp * p.slab + orow * p.Cout + ch0;
            const float* rp = p.res ? p.res + orow * p.Cout + ch0 : nullptr;
            // SPLIT folds bias and residual into the initial value of the running sums (loads issued before the MMAs of
            // the tile are waited for); a bias load inside the store phase costs an exposed L2 round trip per 16 columns
            const float* bias_e = SPLIT ? nullptr : p.bias;

            auto finish = [&](const uint32_t (&r)[32], int c) {   // bias / residual / ReLU / store of 32 columns
                if (!valid) return;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    if (bias_e) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias_e + ch0 + c * 32 + 4 * j));
                        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                    }
                    if (rp) {
                        const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + c * 32 + 4 * j));
                        v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
                    }
                    if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    if (p.round_out) { v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w); }
                    *reinterpret_cast<float4*>(op + c * 32 + 4 * j) = v;
                }
            };

            // TMA-store variant (flat mode): 16 columns of every row of the tile go through a swizzled staging buffer
            // and leave as ONE bulk tensor store -- no LSU wavefronts for the output, rows >= M are clipped by TMA.
            int sbuf = 0;
            auto emit_tma16 = [&](const float* v, int c16) {
                uint8_t* stg = out_stage + (grp * 2 + sbuf) * C::OUT_STAGE;
                PROF_T(e_wait, if (row == 0) bulk_wait_read<1>());   // the store that used this buffer two chunks ago is done
                PROF_T(e_bar1, named_bar_sync(1 + grp, 128));
                float x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = v[j];
#ifdef SCOUTER_PROF
                const long long _tb = clock64();
#endif
                if (bias_e) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias_e + ch0 + c16 * 16 + 4 * j));
                        x[4 * j] += bv.x; x[4 * j + 1] += bv.y; x[4 * j + 2] += bv.z; x[4 * j + 3] += bv.w;
                    }
                }
                if (rp && valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + c16 * 16 + 4 * j));
                        x[4 * j] += rv.x; x[4 * j + 1] += rv.y; x[4 * j + 2] += rv.z; x[4 * j + 3] += rv.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (p.relu) x[j] = fmaxf(x[j], 0.f);
                    if (p.round_out) x[j] = to_tf32(x[j]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)   // SWIZZLE_64B: 16-byte chunk index ^= (row / 2) % 4
                    *reinterpret_cast<float4*>(stg + row * 64 + ((j ^ ((row >> 1) & 3)) << 4)) =
                        make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
#ifdef SCOUTER_PROF
                prof_e_math += clock64() - _tb;
#endif
                PROF_T(e_fence, fence_proxy_async());
                PROF_T(e_bar2, named_bar_sync(1 + grp, 128));
                if (row == 0) {
                    PROF_T(e_tma, tma_store_3d(&tmO, stg, ch0 + c16 * 16, mt * 128, sp); bulk_commit());
                }
                sbuf ^= 1;
            };

            if constexpr (SPLIT) {
                // The running fp32 sum starts from the residual: its loads are issued here, before the first chunk is
                // waited for, so their DRAM latency hides behind the tile's MMAs (and costs no extra registers).
                float acc[C::NC];
                const bool with_bias = p.bias && p.ksplit == 1;   // split-K slabs are raw partial sums
#pragma unroll
                for (int j = 0; j < C::NC / 4; ++j) {
                    float4 v = with_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!RES && rp && valid) {
                        const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + 4 * j));
                        v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
                    }
                    acc[4 * j] = v.x; acc[4 * j + 1] = v.y; acc[4 * j + 2] = v.z; acc[4 * j + 3] = v.w;
                }
                rp = nullptr;   // already folded in
                for (int ch = 0; ch < nchunks; ++ch, ++cc) {
                    const int buf = cc & 1;
                    PROF_T(cfull, mbar_wait(&cfull[buf], (cc >> 1) & 1));
                    tc_fence_after();
#ifdef SCOUTER_PROF
                    const long long _tm = clock64();
#endif
#pragma unroll
                    for (int c = 0; c < C::NC / 32; ++c) {
                        uint32_t r[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + col0 + c * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
                    }
                    tc_fence_before();
                    mbar_arrive(&cempty[buf]);
#ifdef SCOUTER_PROF
                    prof_merge += clock64() - _tm;
#endif
                }
#ifdef SCOUTER_PROF
                const long long _ts = clock64();
#endif
                if constexpr (RES) {
                    // residual slabs: add in place (own row, swizzled 16-byte chunks), ReLU, store from the same slab
                    PROF_T(e_wait, mbar_wait(rfull, rphase));
                    rphase ^= 1;
#pragma unroll
                    for (int c = 0; c < C::NC / 16; ++c) {
                        uint8_t* stg = out_stage + (col0 / 16 + c) * C::OUT_STAGE + row * 64;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float4* cell = reinterpret_cast<float4*>(stg + ((j ^ ((row >> 1) & 3)) << 4));
                            float4 v = *cell;
                            v.x += acc[c * 16 + 4 * j]; v.y += acc[c * 16 + 4 * j + 1]; v.z += acc[c * 16 + 4 * j + 2]; v.w += acc[c * 16 + 4 * j + 3];
                            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            if (p.round_out) { v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w); }
                            *cell = v;
                        }
                        // per slab: its store starts while the next slab is still being added (batching the fences was measured
                        // SLOWER on the HBM-bound residual convs: 355 -> 382 us on layer 1)
                        fence_proxy_async();
                        named_bar_sync(1 + grp, 128);
                        if (row == 0) {
                            tma_store_3d(&tmO, out_stage + (col0 / 16 + c) * C::OUT_STAGE, ch0 + c * 16, mt * 128, 0);
                            bulk_commit();
                        }
                    }
                    if (row == 0) {
                        bulk_wait_read<0>();      // the slabs may now be refilled with the next tile's residual
                        mbar_arrive(rempty);
                    }
#ifdef SCOUTER_PROF
                    prof_store += clock64() - _ts;
#endif
                } else if (p.tma_store) {
                    // all NC columns of the row go to the group's slabs, then one fence + barrier and the stores (bias and the
                    // register-path residual are already in acc)
                    if (row == 0) bulk_wait_read<0>();            // the previous tile's stores have read the slabs
                    named_bar_sync(1 + grp, 128);
#pragma unroll
                    for (int c = 0; c < C::NC / 16; ++c) {
                        uint8_t* stg = out_stage + (col0 / 16 + c) * C::OUT_STAGE + row * 64;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float4 v = make_float4(acc[c * 16 + 4 * j], acc[c * 16 + 4 * j + 1], acc[c * 16 + 4 * j + 2], acc[c * 16 + 4 * j + 3]);
                            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            if (p.round_out) { v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w); }
                            *reinterpret_cast<float4*>(stg + ((j ^ ((row >> 1) & 3)) << 4)) = v;   // SWIZZLE_64B
                        }
                    }
                    fence_proxy_async();
                    named_bar_sync(1 + grp, 128);
                    if (row == 0) {
#pragma unroll
                        for (int c = 0; c < C::NC / 16; ++c)
                            tma_store_3d(&tmO, out_stage + (col0 / 16 + c) * C::OUT_STAGE, ch0 + c * 16, mt * 128, sp);
                        bulk_commit();
                    }
#ifdef SCOUTER_PROF
                    prof_store += clock64() - _ts;
#endif
                } else {
#pragma unroll
                    for (int c = 0; c < C::NC / 32; ++c) {
                        uint32_t r[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(acc[c * 32 + j]);
                        finish(r, c);
                    }
                }
            } else {
                const int buf = cc & 1;
                mbar_wait(&cfull[buf], (cc >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + c * 32, r);
                    tmem_ld_wait();
                    if (p.tma_store) {
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        emit_tma16(&v[0], 2 * c);
                        emit_tma16(&v[16], 2 * c + 1);
                    } else {
                        finish(r, c);
                    }
                }
                tc_fence_before();
                mbar_arrive(&cempty[buf]);
                ++cc;
            }
        }
        if ((RES || p.tma_store) && row == 0) bulk_wait<0>();   // all bulk stores of this group have completed
        PROF_END(epi);
        if (threadIdx.x == 128) {
            PROF_STORE(g_prof_flat, 10, epi); PROF_STORE(g_prof_flat, 11, cfull); PROF_STORE(g_prof_flat, 12, store);
            PROF_STORE(g_prof_flat, 13, merge); PROF_STORE(g_prof_flat, 16, e_wait); PROF_STORE(g_prof_flat, 17, e_bar1);
            PROF_STORE(g_prof_flat, 18, e_math); PROF_STORE(g_prof_flat, 19, e_fence); PROF_STORE(g_prof_flat, 20, e_bar2); PROF_STORE(g_prof_flat, 21, e_tma);
        }
    } else if (SPLIT && warp >= 8 && warp < 12) {
        // ===== operand splitters: fp16(x), bf16(x), bf16(x - fp16(x)) tiles of the activation (and, unless pre-split, weight) tile =====
        const int tid = threadIdx.x - 256;  // 0..127
        PROF_DECL(pfull); PROF_DECL(spl); PROF_BEGIN(spl);
        if constexpr (TS) {
            // thread = row of the tile: fp16 | bf16 | bf16 remainder of its 32 channels -> TMEM operand buffer gk & 3
            const uint32_t sw = (uint32_t)(tid & 7);
            const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::OP_COL0;
            int sa_i = 0;
            uint32_t aphase = 0, gk = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                for (int kb = 0; kb < p.kblocks; ++kb, ++gk) {
                    PROF_T(pfull, mbar_wait(&fullA[sa_i], aphase));
                    const uint8_t* src = smem + sa_i * C::A_BYTES + tid * 128;
                    uint32_t f[32], xh[16], xb[16], rb[16];
#pragma unroll
                    for (uint32_t c = 0; c < 8; ++c) {
                        const uint4 v = *reinterpret_cast<const uint4*>(src + ((c ^ sw) << 4));   // SWIZZLE_128B
                        f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) split2_act(__uint_as_float(f[2 * i]), __uint_as_float(f[2 * i + 1]), xh[i], xb[i], rb[i]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&emptyA[sa_i]);          // the tile lives in registers now
                    if (gk >= (uint32_t)C::NO) {
                        const uint32_t j = gk - C::NO;
                        mbar_wait(&done[j & 3], (j >> 2) & 1u);         // the MMAs that read this operand buffer have retired
                    }
                    tc_fence_after();
                    const uint32_t t0 = t_lane + (uint32_t)C::OP_COLS * (gk & 3);
                    tmem_st_32x16(t0, xh);
                    tmem_st_32x16(t0 + 16, xb);
                    tmem_st_32x16(t0 + 32, rb);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&opfull[gk & 3]);
                    if (++sa_i == C::NA) { sa_i = 0; aphase ^= 1; }
                }
            }
        } else {
            const SplitLane sl = split_lane(tid);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    PROF_T(pfull, mbar_wait(&full[stage], phase));
                    uint8_t* st = smem + stage * C::STAGE;
                    split_tile_hbr<128>(st, st + C::OFF_A16, C::A_TILE, sl);
                    if (!p.rem_rows) split_tile_hbr<BN, true>(st + C::A_BYTES, st + C::OFF_W16, C::B_TILE, sl);
                    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&split_done[stage]);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        PROF_END(spl);
        if (tid == 0) { PROF_STORE(g_prof_flat, 14, spl); PROF_STORE(g_prof_flat, 15, pfull); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

void choose_tile(int B, int H, int W, int& Wb, int& Hb, int& Nb) {
    double best = -1.0;
    Wb = Hb = Nb = 1;
    for (int wb = 1; wb <= W && wb <= 128; ++wb)
        for (int hb = 1; hb <= H && wb * hb <= 128; ++hb) {
            int nbmax = (wb == W && hb == H) ? std::min(B, 128 / (wb * hb)) : 1;
            for (int nb = 1; nb <= nbmax; ++nb) {
                long long tiles = (long long)cdiv(W, wb) * cdiv(H, hb) * cdiv(B, nb);
                double util = (double)B * H * W / (tiles * 128.0) + 1e-6 * wb;  // tie -> longer contiguous runs
                if (util > best) { best = util; Wb = wb; Hb = hb; Nb = nb; }
            }
        }
}

int pick_bn(int cout_g) {
    static int cap = [] { const char* e = getenv("SCOUTER_UMMA_BN"); int v = e ? atoi(e) : 128; return v > 128 ? 128 : v; }();
    for (int bn : {128, 64, 32})
        if (bn <= cap && cout_g % bn == 0) return bn;
    return 0;
}

template <int BN, bool SPLIT, bool RES = false, bool TS = false>
int launch_bn(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tB2, const CUtensorMap& tO, const CUtensorMap& tR,
              const UmmaArgs& u, int grid, cudaStream_t s) {
    using C = Cfg<BN, SPLIT, RES, TS>;
    static_assert(TS ? (C::NA >= 3 && C::NW >= 2) : C::STAGES >= 2, "pipeline too shallow");
    static_assert(C::SMEM <= 227 * 1024, "shared memory budget");
    SC_CUDA(cudaFuncSetAttribute(conv_umma_kernel<BN, SPLIT, RES, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    conv_umma_kernel<BN, SPLIT, RES, TS><<<grid, C::THREADS, C::SMEM, s>>>(tA, tB, tB2, tO, tR, u);
    SC_LAUNCH_CHECK();
    return 0;
}

}  // namespace

bool umma_conv_supported(const ConvArgs& a) {
    if (a.stride != 1 || a.kh != a.kw) return false;
    if (!((a.kh == 1 && a.pad == 0) || (a.kh == 3 && a.pad == 1))) return false;
    if (a.Cin % a.groups || a.Cout % a.groups) return false;
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    if (cin_g % 32 || pick_bn(cout_g) == 0) return false;
    if ((long long)a.B * a.H * a.W >= (1ll << 31)) return false;
    return true;
}

int launch_conv_umma(const ConvArgs& a, UmmaConvPlan& plan, cudaStream_t s) {
    SC_CHECK_ARG(umma_conv_supported(a), SCOUTER_E_UNSUPPORTED, "conv_umma: unsupported geometry");
    EncodeTiledFn enc = encode_fn();
    SC_CHECK_ARG(enc, SCOUTER_E_UNSUPPORTED, "conv_umma: cuTensorMapEncodeTiled is not available from the driver");
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    const int BN = pick_bn(cout_g);
    UmmaArgs u;
    u.bias = a.bias; u.res = a.res; u.out = a.out;
    u.mode = a.kh == 3 ? 1 : 0;
    u.M = a.B * a.H * a.W;
    u.B = a.B; u.H = a.H; u.W = a.W;
    u.groups = a.groups; u.cin_g = cin_g; u.cout_g = cout_g; u.Cout = a.Cout;
    u.kw = a.kw; u.pad = a.pad;
    u.cblocks = cin_g / 32;
    u.kblocks = a.kh * a.kw * u.cblocks;
    u.relu = a.relu; u.round_out = a.round_out;
    static int chunk_kb = [] { const char* e = getenv("SCOUTER_UMMA_CHUNK"); int v = e ? atoi(e) : 3; return v < 1 ? 1 : v; }();
    u.chunk = a.split ? chunk_kb : u.kblocks;
    const bool presplit = a.split && a.w_rem != nullptr;
    u.rem_rows = presplit ? a.Cout : 0;
    u.n_tiles = cout_g / BN;

    const bool reuse = plan.valid && plan.in == a.in && plan.w == a.w && plan.B == a.B && plan.H == a.H && plan.W == a.W &&
                       plan.Cin == a.Cin && plan.Cout == a.Cout && plan.kh == a.kh && plan.groups == a.groups && plan.BN == BN &&
                       !plan.halo && plan.presplit == presplit;
    if (!reuse) {
        CUresult r;
        if (u.mode) {
            choose_tile(a.B, a.H, a.W, plan.Wb, plan.Hb, plan.Nb);
            cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
            cuuint64_t strides[3] = {(cuuint64_t)a.Cin * 4, (cuuint64_t)a.W * a.Cin * 4, (cuuint64_t)a.H * a.W * a.Cin * 4};
            cuuint32_t box[4] = {32, (cuuint32_t)plan.Wb, (cuuint32_t)plan.Hb, (cuuint32_t)plan.Nb};
            cuuint32_t es[4] = {1, 1, 1, 1};
            r = enc(&plan.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            plan.Wb = plan.Hb = plan.Nb = 0;
            cuuint64_t dims[2] = {(cuuint64_t)a.Cin, (cuuint64_t)u.M};
            cuuint64_t strides[1] = {(cuuint64_t)a.Cin * 4};
            cuuint32_t box[2] = {32, 128};
            cuuint32_t es[2] = {1, 1};
            r = enc(&plan.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_umma: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
        const cuuint64_t Kt = (cuuint64_t)a.kh * a.kw * cin_g;
        cuuint64_t dimsB[2] = {Kt, (cuuint64_t)a.Cout};
        cuuint64_t stridesB[1] = {Kt * 4};
        cuuint32_t boxB[2] = {32, (cuuint32_t)BN};
        cuuint32_t esB[2] = {1, 1};
        r = enc(&plan.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.w, dimsB, stridesB, boxB, esB, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_umma: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
        if (presplit) {   // 16-bit [fp16 W_h ; bf16 W_r]: (2*Cout) rows of Kt elements, moved as raw 16-bit words
            cuuint64_t dimsB2[2] = {Kt, (cuuint64_t)2 * a.Cout};
            cuuint64_t stridesB2[1] = {Kt * 2};
            r = enc(&plan.tmB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)a.w_rem, dimsB2, stridesB2, boxB, esB,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_umma: cuTensorMapEncodeTiled(bf16 W) failed with %d", (int)r);
        } else {
            plan.tmB2 = plan.tmB;
        }
        plan.valid = true; plan.halo = false; plan.presplit = presplit;
        plan.in = a.in; plan.w = a.w; plan.B = a.B; plan.H = a.H; plan.W = a.W; plan.Cin = a.Cin; plan.Cout = a.Cout;
        plan.kh = a.kh; plan.groups = a.groups; plan.BN = BN;
    }
    if (u.mode) {
        u.Wb = plan.Wb; u.Hb = plan.Hb; u.Nb = plan.Nb;
        u.tw = cdiv(a.W, u.Wb); u.th = cdiv(a.H, u.Hb);
        u.m_tiles = u.tw * u.th * cdiv(a.B, u.Nb);
        u.a_bytes = u.Wb * u.Hb * u.Nb * 128;
    } else {
        u.Wb = u.Hb = u.Nb = u.tw = u.th = 1;
        u.m_tiles = cdiv(u.M, 128);
        u.a_bytes = 128 * 128;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        SC_CUDA(cudaGetDevice(&dev));
        SC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    static bool no_tma_store = getenv("SCOUTER_NO_TMA_STORE") != nullptr;
    u.tma_store = (!u.mode && !no_tma_store && a.Cout % 16 == 0) ? 1 : 0;
    if (u.tma_store && !(reuse && plan.out == a.out && plan.ksplit == std::max(1, a.ksplit))) {
        const int ks = std::max(1, a.ksplit);
        cuuint64_t dimsO[3] = {(cuuint64_t)a.Cout, (cuuint64_t)u.M, (cuuint64_t)ks};
        cuuint64_t stridesO[2] = {(cuuint64_t)a.Cout * 4, (cuuint64_t)u.M * a.Cout * 4};
        cuuint32_t boxO[3] = {16, 128, 1};
        cuuint32_t esO[3] = {1, 1, 1};
        CUresult r = enc(&plan.tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a.out, dimsO, stridesO, boxO, esO,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_umma: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
        plan.out = a.out;
        plan.ksplit = ks;
    }
    u.ksplit = 1;
    u.slab = 0;
    if (a.ksplit > 1) {
        SC_CHECK_ARG(u.kblocks % a.ksplit == 0 && !a.res, SCOUTER_E_INVALID, "conv_umma: ksplit=%d does not divide %d k-blocks", a.ksplit, u.kblocks);
        u.ksplit = a.ksplit;
        u.kblocks /= a.ksplit;
        u.slab = (long long)u.M * a.Cout;
        u.bias = nullptr; u.relu = 0; u.round_out = 0;
        if (!a.split) u.chunk = u.kblocks;
    }
    const long long total = (long long)u.m_tiles * u.n_tiles * u.groups * u.ksplit;
    const int grid = (int)std::min<long long>(total, sms);
    // residual convs (the block-final 1x1 conv3 + shortcut) take the in-place residual epilogue
    static bool no_res_tma = getenv("SCOUTER_NO_RES_TMA") != nullptr;
    const bool res_tma = a.split && a.res && u.tma_store && u.ksplit == 1 && !no_res_tma;
    // activation operands in tensor memory (TS-mode MMAs): needs the host-pre-split weights
    static bool no_ts = getenv("SCOUTER_UMMA_NO_TS") != nullptr;
    const bool ts = presplit && !no_ts;
    if (res_tma) {
        if (!(reuse && plan.res == a.res)) {
            cuuint64_t dimsR[2] = {(cuuint64_t)a.Cout, (cuuint64_t)u.M};
            cuuint64_t stridesR[1] = {(cuuint64_t)a.Cout * 4};
            cuuint32_t boxR[2] = {16, 128};
            cuuint32_t esR[2] = {1, 1};
            CUresult r = enc(&plan.tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.res, dimsR, stridesR, boxR, esR,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_umma: cuTensorMapEncodeTiled(residual) failed with %d", (int)r);
            plan.res = a.res;
        }
    }
    if (res_tma && ts) {
        switch (BN) {
            case 32: return launch_bn<32, true, true, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmR, u, grid, s);
            case 64: return launch_bn<64, true, true, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmR, u, grid, s);
            case 128: return launch_bn<128, true, true, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmR, u, grid, s);
        }
    } else if (a.split && ts) {
        switch (BN) {
            case 32: return launch_bn<32, true, false, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmO, u, grid, s);
            case 64: return launch_bn<64, true, false, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmO, u, grid, s);
            case 128: return launch_bn<128, true, false, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmO, u, grid, s);
        }
    } else if (res_tma) {
        switch (BN) {
            case 32: return launch_bn<32, true, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmR, u, grid, s);
            case 64: return launch_bn<64, true, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmR, u, grid, s);
            case 128: return launch_bn<128, true, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmR, u, grid, s);
        }
    } else if (a.split) {
        switch (BN) {
            case 32: return launch_bn<32, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmO, u, grid, s);
            case 64: return launch_bn<64, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmO, u, grid, s);
            case 128: return launch_bn<128, true>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmO, u, grid, s);
        }
    } else {
        switch (BN) {
            case 32: return launch_bn<32, false>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmO, u, grid, s);
            case 64: return launch_bn<64, false>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmO, u, grid, s);
            case 128: return launch_bn<128, false>(plan.tmA, plan.tmB, plan.tmB2, plan.tmO, plan.tmO, u, grid, s);
        }
    }
    return SCOUTER_E_UNSUPPORTED;
}

}  // namespace scouter

#ifdef SCOUTER_PROF
extern "C" int scouter_prof_read_flat(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, scouter::g_prof_flat, sizeof(unsigned long long) * n);
}
#endif
