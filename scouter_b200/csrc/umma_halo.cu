// 3x3 (stride 1, pad 1) convolution on tcgen05 from ONE halo tile per channel block -- error-compensated fp16 product.
//
// The tap-reload kernel (umma_conv.cu) re-fetches the activation tile from L2 nine times, once per tap.  Here the
// producer loads the (Hb+2) x (Wb+2) pixel halo patch of 32 channels ONCE (a single 4-D TMA box, out-of-bound
// zero fill = the padding) and the nine taps are nine UMMA descriptors into that patch: a K-major SWIZZLE_128B
// operand may start at any 128-byte row of a TMA-written region (the swizzle is a function of the absolute
// shared-memory address; measured in profiles/r01_umma_desc_row_offset_probe.txt), so tap (r,s) is simply
//     start = patch + (r*PW + s) * 128 bytes,           PW = Wb + 2.
// The M dimension then walks 128 CONSECUTIVE patch rows, i.e. "virtual" output rows m' = hb*PW + wb' that include
// the two halo columns of every line (computed and discarded: Wb/PW of the MMA rows are useful).
//
// Precision (round 2): a*w ~= a_h*w_h + a_b*w_r + a_r*w_h (ptx.cuh split2_act / split2_wgt: h = fp16, r = x - h exact).  Per
// (channel block, tap) the weights arrive as two host-made 16-bit TMA tiles -- fp16 W_h, bf16 W_r -- and four splitter warps
// turn each fp32 patch into fp16 A_h, bf16 A_b and fp16 A_r patches once per patch (not once per tap).  The issuer accumulates
//     A_b*W_r  (bf16 x bf16)  +  A_r*W_h  +  A_h*W_h  (fp16 x fp16)        = 2 + 2 + 2 kind::f16 MMAs of K = 16
// in TMEM in chunks of `chunk` k-steps; epilogue threads merge the chunks in fp32 registers (the TMEM accumulator truncates
// on every MMA -- profiles/r01_tmem_accumulator_truncation.txt).  Against round 1's tf32 main product (4 MMAs of K = 8 +
// 2 + 2 bf16 corrections) this is 6 instead of 8 MMA times per 32 channels, half the weight bytes from L2 and through shared
// memory, and a rounded (not truncated) 11-bit main operand.  (fp16 x bf16 in ONE kind::f16 MMA would save the bf16 copy of
// A as well, but is an illegal instruction on sm_100a -- tried.)
//
// FOLD (narrow outputs, cout/groups <= 64): every MMA re-reads its 128 x 32-byte A slice from shared memory whatever N is,
// so nine taps x three products x two K halves of N = 32 cost 216 KB of operand reads per 128-row tile -- these layers
// (deep-stem 3x3s, layer-1 split-attention convs) ran at 1.5-1.9 TB/s of their HBM traffic.  The folded form puts the
// three taps s = 0,1,2 of one kernel row r into N:  D[m, s*Cout + co] = sum_{r,c} patch[m + r*PW, c] * W[co, r, s, c]  (three
// A start rows instead of nine, N = 3*Cout), and the epilogue finishes  out[m] = sum_s D[m + s, block s]  with warp shuffles
// -- patch lines are PW = 16 rows, so a line never straddles a warp's 32 TMEM lanes and the lanes that would read across
// the line end are the two discarded halo columns.  A reads and MMA issues drop 3x.
//
// Warp roles (512 threads, 1 CTA/SM, persistent): 0 TMA producer | 1 MMA issuer | 2 TMEM alloc | 4-7 and 12-15
// epilogue (two column halves) | 8-11 splitters.
//
// Replaces: the grouped 3x3 conv + bn0 + ReLU of timm/models/layers/split_attn.py:43-45,55-60 and the deep-stem
// 3x3 convs of timm/models/resnet.py:404-408 (eval mode, BN folded).
#include "ptx.cuh"
#include "umma.cuh"

namespace scouter {
using namespace ptx;

namespace {

#ifdef SCOUTER_PROF
__device__ unsigned long long g_prof_halo[256 * 32];
#endif

struct HaloArgs {
    const float* bias;
    const float* res;
    float* out;
    int B, H, W;
    int Wb, Hb, PW, tw, th;
    int m_tiles, n_tiles, groups;
    int cin_g, cout_g, Cout, cblocks;
    int relu;
    int patch_bytes;   // TMA bytes of one raw patch = PH*PW*128
    int patch_alloc;   // bytes reserved for the raw fp32 patch (each 16-bit patch takes half of it), multiple of 2048
    int pst, bst;      // ring depths
    int chunk;         // k-steps (cb,tap pairs) per accumulation chunk
    int tma_store;     // epilogue leaves through swizzled staging + 4-D TMA stores (box {16 ch, Wb, Hb, 1})
    float* gap_part;   // optional (B, gap_slots, Cout): column sums of the stored tile per epilogue warp (slot = tile-in-image * 4 + q)
    int gap_slots;
    int wres;          // FOLD: the weights of one (group, n-tile) stay RESIDENT in shared memory (bst = cblocks*3 stages, loaded
                       // when the key changes) -- streamed per tile they are 108 KB from L2 against a 20 KB patch (L2-bound)
};

template <int BN, bool FOLD = false>
struct HCfg {
    static_assert(!FOLD || BN <= 64, "the folded form needs N = 3*BN <= 256");
    static constexpr int NB = FOLD ? 3 * BN : BN;   // N of one MMA = columns of one TMEM accumulator buffer
    static constexpr int STEPS = FOLD ? 3 : 9;      // k-steps (weight stages) per channel block: kernel rows / taps
    static constexpr int B_TILE = NB * 64;          // one 16-bit weight tile: NB rows x 64 bytes
    static constexpr int B_STAGE = 2 * B_TILE;      // [fp16 W_h | bf16 W_r]
    static constexpr int BUF_COLS = NB <= 32 ? 32 : (NB <= 64 ? 64 : (NB <= 128 ? 128 : 256));   // column stride of the accumulator buffers
    // Accumulator buffers (one per accumulation chunk in flight).  Four where tensor memory allows: the store phase of a tile
    // takes ~3.4k clk during which the issuer can only run NBUF chunks ahead -- with two buffers of 2 k-steps it stalled on
    // `cempty` for 30 % of its time (role profile, layer2.0 conv2).
    static constexpr int NBUF = 4 * BUF_COLS <= 512 ? 4 : 2;
    static constexpr int TMEM_COLS = NBUF * BUF_COLS < 32 ? 32 : NBUF * BUF_COLS;
    static constexpr int EPI_GROUPS = (BN == 128 || FOLD) ? 2 : 1;   // FOLD: 3*BN accumulator columns + shuffles per output row
    static constexpr int NC = BN / EPI_GROUPS;
    static constexpr int MAX_ST = 8;
    static constexpr int OUT_STAGE = 128 * 64;   // staging of 16 columns x <=128 compact tile rows, SWIZZLE_64B
    static constexpr int OUT_BUFS = FOLD ? 1 : 2;   // staging buffers per epilogue group (FOLD: shared memory holds the weights)
    static constexpr int OUT_BYTES = EPI_GROUPS * OUT_BUFS * OUT_STAGE;
};

template <int BN, bool FOLD>
__global__ void __launch_bounds__(512, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB2,
                    const __grid_constant__ CUtensorMap tmO, const HaloArgs p) {
    using C = HCfg<BN, FOLD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int pslot = p.patch_alloc / 2 * 5;                     // one patch stage: [raw fp32 | fp16 A_h | bf16 A_b | bf16 A_r]
    uint8_t* patch0 = smem;
    uint8_t* bt0 = smem + p.pst * pslot;                         // bst x [W_h | W_r]
    uint8_t* out_stage = bt0 + p.bst * C::B_STAGE;               // [EPI_GROUPS][2][OUT_STAGE], 1024-aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + C::OUT_BYTES);
    uint64_t* pfull = bars;                  // [MAX_ST] patch landed (TMA)
    uint64_t* pready = pfull + C::MAX_ST;    // [MAX_ST] remainder written (128 splitter threads)
    uint64_t* pempty = pready + C::MAX_ST;   // [MAX_ST] all MMAs that read the patch retired
    uint64_t* bfull = pempty + C::MAX_ST;    // [MAX_ST]
    uint64_t* bempty = bfull + C::MAX_ST;    // [MAX_ST]
    uint64_t* cfull = bempty + C::MAX_ST;    // [<= 4] accumulator chunk complete
    uint64_t* cempty = cfull + 4;            // [<= 4]
    uint64_t* wfull = cempty + 4;            // resident weights landed
    uint64_t* wempty = wfull + 1;            // every MMA that read the resident weights retired (before a reload)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wempty + 1);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB2);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < C::MAX_ST; ++i) {
            mbar_init(&pfull[i], 1);
            mbar_init(&pready[i], 4);   // one arrival per splitter warp
            mbar_init(&pempty[i], 1);
            mbar_init(&bfull[i], 1);
            mbar_init(&bempty[i], 1);
        }
        for (int i = 0; i < C::NBUF; ++i) {
            mbar_init(&cfull[i], 1);
            mbar_init(&cempty[i], 128 * C::EPI_GROUPS);
        }
        mbar_init(wfull, 1);
        mbar_init(wempty, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int total = p.m_tiles * p.n_tiles * p.groups;
    const int ksteps = p.cblocks * C::STEPS;
    const int nchunks = (ksteps + p.chunk - 1) / p.chunk;

    auto tile_coords = [&](int t, int& nt, int& g, int& w0, int& h0, int& b) {
        nt = t % p.n_tiles;
        const int mt = (t / p.n_tiles) % p.m_tiles;
        g = t / (p.n_tiles * p.m_tiles);
        w0 = (mt % p.tw) * p.Wb;
        h0 = ((mt / p.tw) % p.th) * p.Hb;
        b = mt / (p.tw * p.th);
    };

    if (warp == 0) {
        if (elect_one()) {
            // ===== TMA producer: patch job j+1 is issued before the nine weight tiles of job j =====
            int ps = 0, bs = 0;
            uint32_t pphase = 0, bphase = 0;
            PROF_DECL(pempty); PROF_DECL(bempty); PROF_DECL(prod); PROF_BEGIN(prod);
            auto issue_patch = [&](int t, int cb) {
                int nt, g, w0, h0, b;
                tile_coords(t, nt, g, w0, h0, b);
                PROF_T(pempty, mbar_wait(&pempty[ps], pphase ^ 1));
                mbar_arrive_expect_tx(&pfull[ps], (uint32_t)p.patch_bytes);
                tma_load_4d(patch0 + ps * pslot, &tmA, &pfull[ps], g * p.cin_g + cb * 32, w0 - 1, h0 - 1, b);
                if (++ps == p.pst) { ps = 0; pphase ^= 1; }
            };
            // patch cursor runs (pst - 1) jobs ahead of the weight-tile cursor
            int pt = blockIdx.x, pcb = 0;
            auto issue_next_patch = [&]() {
                if (pt >= total) return;
                issue_patch(pt, pcb);
                if (++pcb == p.cblocks) { pcb = 0; pt += gridDim.x; }
            };
            for (int i = 0; i < p.pst - 1; ++i) issue_next_patch();
            int wkey = -1;
            uint32_t wephase = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                int nt, g, w0, h0, b;
                tile_coords(t, nt, g, w0, h0, b);
                if (FOLD && p.wres) {
                    const int key = g * p.n_tiles + nt;
                    if (key != wkey) {     // (re)load the resident weights: all channel blocks x kernel rows of this (group, n-tile)
                        if (wkey >= 0) { mbar_wait(wempty, wephase); wephase ^= 1; }
                        mbar_arrive_expect_tx(wfull, (uint32_t)(p.cblocks * 3 * C::B_STAGE));
                        const int nrow = g * p.cout_g + nt * BN;
                        for (int st = 0; st < p.cblocks * 3; ++st)
                            for (int s3 = 0; s3 < 3; ++s3) {
                                const int kcol = ((st % 3) * 3 + s3) * p.cin_g + (st / 3) * 32;
                                uint8_t* d = bt0 + st * C::B_STAGE + s3 * BN * 64;
#pragma unroll
                                for (int j = 0; j < 2; ++j) tma_load_2d(d + j * C::B_TILE, &tmB2, wfull, kcol, j * p.Cout + nrow);
                            }
                        wkey = key;
                    }
                    for (int cb = 0; cb < p.cblocks; ++cb) issue_next_patch();
                    continue;
                }
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    for (int st = 0; st < C::STEPS; ++st) {
                        if (st == C::STEPS / 3) issue_next_patch();  // the issuer is inside this job: the oldest patch stage is (about to be) free
                        PROF_T(bempty, mbar_wait(&bempty[bs], bphase ^ 1));
                        uint8_t* sb = bt0 + bs * C::B_STAGE;
                        mbar_arrive_expect_tx(&bfull[bs], (uint32_t)C::B_STAGE);
                        const int nrow = g * p.cout_g + nt * BN;
#pragma unroll
                        for (int s3 = 0; s3 < (FOLD ? 3 : 1); ++s3) {     // FOLD: taps (st, 0..2) stacked along N
                            const int tap = FOLD ? st * 3 + s3 : st;
                            const int kcol = tap * p.cin_g + cb * 32;
                            uint8_t* d = sb + s3 * BN * 64;
                            tma_load_2d(d, &tmB2, &bfull[bs], kcol, nrow);                      // fp16 W_h
                            tma_load_2d(d + C::B_TILE, &tmB2, &bfull[bs], kcol, p.Cout + nrow);  // bf16 W_r
                        }
                        if (++bs == p.bst) { bs = 0; bphase ^= 1; }
                    }
                }
            }
            PROF_END(prod);
            PROF_STORE(g_prof_halo, 0, prod); PROF_STORE(g_prof_halo, 1, pempty); PROF_STORE(g_prof_halo, 2, bempty);
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            // One thread, ~4.5 clk per dependent instruction: everything that can be hoisted is.  Descriptors are a
            // constant high word plus a low word advanced by 32-bit adds; barrier addresses are 32-bit shared-window
            // addresses; the nine taps are unrolled so that (r, s) and the k sub-steps are immediates.
            constexpr uint32_t id_h = idesc_f16(128, C::NB), id_b = idesc_bf16(128, C::NB);
            constexpr uint32_t BSTEP = C::B_STAGE >> 4, WR_OFF = C::B_TILE >> 4;
            const uint32_t pa_lo0 = desc_lo(smem_u32(patch0)), b_lo0 = desc_lo(smem_u32(bt0));
            const uint32_t pstep = (uint32_t)pslot >> 4;
            const uint32_t ah_off = (uint32_t)p.patch_alloc >> 4, half = (uint32_t)(p.patch_alloc / 2) >> 4;
            const uint32_t pw4 = (uint32_t)p.PW * 4u;   // one patch line of 64-byte rows in 16-byte units
            const uint32_t pready_a = smem_u32(pready), pempty_a = smem_u32(pempty), bfull_a = smem_u32(bfull),
                           bempty_a = smem_u32(bempty), cfull_a = smem_u32(cfull), cempty_a = smem_u32(cempty);
            const int last_cb = p.cblocks - 1, chunk = p.chunk, pst = p.pst, bst = p.bst;
            int ps = 0, bs = 0, in_chunk = 0;
            uint32_t pphase = 0, bphase = 0, cc = 0;
            uint32_t pa_lo = pa_lo0, b_lo = b_lo0;
            const bool wres = FOLD && p.wres;
            int wkey = -1;
            uint32_t wfphase = 0;
            auto tile_key = [&](int t) { return (t / (p.n_tiles * p.m_tiles)) * p.n_tiles + t % p.n_tiles; };
            PROF_DECL(pready); PROF_DECL(cempty); PROF_DECL(bfull); PROF_DECL(iss); PROF_DECL(ntaps); PROF_BEGIN(iss);
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                if (wres && tile_key(t) != wkey) {
                    mbar_wait(wfull, wfphase);
                    wfphase ^= 1;
                    wkey = tile_key(t);
                }
                for (int cb = 0; cb <= last_cb; ++cb) {
                    if (wres) b_lo = b_lo0 + (uint32_t)(cb * 3) * BSTEP;
                    PROF_T(pready, mbar_wait_a(pready_a + 8 * ps, pphase));
#pragma unroll
                    for (int tap = 0; tap < C::STEPS; ++tap) {
                        const uint32_t buf = cc % C::NBUF;
                        if (in_chunk == 0) PROF_T(cempty, mbar_wait_a(cempty_a + 8 * buf, ((cc / C::NBUF) & 1) ^ 1));
                        if (!wres) PROF_T(bfull, mbar_wait_a(bfull_a + 8 * bs, bphase));
                        tc_fence_after();
                        const uint32_t r = FOLD ? tap : tap / 3, sx = FOLD ? 0 : tap % 3;   // immediates after unrolling
                        const uint32_t ah = pa_lo + ah_off + r * pw4 + sx * 4u;   // fp16 patch, tap row (r*PW + s) * 64 B
                        const uint32_t ab = ah + half, ar = ab + half;            // bf16 patch, fp16 remainder patch
                        const uint32_t d_tmem = tmem_base + buf * C::BUF_COLS;
                        const uint32_t acc = in_chunk != 0;
#pragma unroll
                        for (uint32_t k = 0; k < 2; ++k)   // A_b * W_r   (bf16)
                            umma_bf16(d_tmem, desc_make(DESC_HI_SW64, ab + 2 * k), desc_make(DESC_HI_SW64, b_lo + WR_OFF + 2 * k), id_b, acc | k);
#pragma unroll
                        for (uint32_t k = 0; k < 2; ++k)   // A_r * W_h   (fp16)
                            umma_bf16(d_tmem, desc_make(DESC_HI_SW64, ar + 2 * k), desc_make(DESC_HI_SW64, b_lo + 2 * k), id_h, 1);
#pragma unroll
                        for (uint32_t k = 0; k < 2; ++k)   // A_h * W_h   (fp16)
                            umma_bf16(d_tmem, desc_make(DESC_HI_SW64, ah + 2 * k), desc_make(DESC_HI_SW64, b_lo + 2 * k), id_h, 1);
                        if (wres) {
                            b_lo += BSTEP;
                        } else {
                            umma_commit_a(bempty_a + 8 * bs);
                            if (++bs == bst) { bs = 0; bphase ^= 1; b_lo = b_lo0; } else { b_lo += BSTEP; }
                        }
                        if (++in_chunk == chunk || (tap == C::STEPS - 1 && cb == last_cb)) {
                            umma_commit_a(cfull_a + 8 * buf);
                            ++cc;
                            in_chunk = 0;
                        }
                    }
                    umma_commit_a(pempty_a + 8 * ps);
                    if (++ps == pst) { ps = 0; pphase ^= 1; pa_lo = pa_lo0; } else { pa_lo += pstep; }
#ifdef SCOUTER_PROF
                    prof_ntaps += C::STEPS;
#endif
                }
                if (wres && t + (int)gridDim.x < total && tile_key(t + gridDim.x) != wkey) umma_commit_a(smem_u32(wempty));
            }
            PROF_END(iss);
            PROF_STORE(g_prof_halo, 4, iss); PROF_STORE(g_prof_halo, 5, pready); PROF_STORE(g_prof_halo, 6, cempty);
            PROF_STORE(g_prof_halo, 7, bfull); PROF_STORE(g_prof_halo, 8, ntaps);
        }
    } else if ((warp >= 4 && warp < 8) || (C::EPI_GROUPS == 2 && warp >= 12)) {
        // ===== epilogue: thread = one virtual output row (hb*PW + wb') x NC columns =====
        const int q = warp & 3;
        const int grp = warp >= 12 ? 1 : 0;
        const int row = q * 32 + lane;
        const int col0 = grp * C::NC;
        const int hb = row / p.PW, wb = row - hb * p.PW;
        uint32_t cc = 0;
        PROF_DECL(cfull); PROF_DECL(store); PROF_DECL(epi); PROF_BEGIN(epi);
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            int nt, g, w0, h0, b;
            tile_coords(t, nt, g, w0, h0, b);
            const int h = h0 + hb, w = w0 + wb;
            const bool valid = hb < p.Hb && wb < p.Wb && h < p.H && w < p.W;
            const long long orow = ((long long)b * p.H + h) * p.W + w;
            const int ch0 = g * p.cout_g + nt * BN + col0;
            // The running fp32 sums start from bias (+ residual): those loads are issued here, before the first chunk
            // is waited for, so their latency hides behind the tile's MMAs instead of sitting in the store phase.
            float acc[C::NC];
            const float* rp = (p.res && valid) ? p.res + orow * p.Cout + ch0 : nullptr;
#pragma unroll
            for (int j = 0; j < C::NC / 4; ++j) {
                float4 v = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (rp) {
                    const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + 4 * j));
                    v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
                }
                acc[4 * j] = v.x; acc[4 * j + 1] = v.y; acc[4 * j + 2] = v.z; acc[4 * j + 3] = v.w;
            }
            for (int ch = 0; ch < nchunks; ++ch, ++cc) {
                const int buf = cc % C::NBUF;
                PROF_T(cfull, mbar_wait(&cfull[buf], (cc / C::NBUF) & 1));
                tc_fence_after();
                if constexpr (FOLD) {
                    // out[m] = sum_s D[m + s, block s]: lane m takes block s from lane m + s of its own warp (a patch line is 16
                    // lanes; the lanes that would wrap are the discarded halo columns wb >= Wb)
#pragma unroll
                    for (int s3 = 0; s3 < 3; ++s3) {
                        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + buf * C::BUF_COLS + s3 * BN + col0;
                        uint32_t r[C::NC];
                        if constexpr (C::NC == 32) {
                            tmem_ld_32x32(ta, r);
                        } else {
                            static_assert(C::NC == 16, "folded epilogue: 16 or 32 columns per thread");
                            uint32_t r0[8], r1[8];
                            tmem_ld_32x8(ta, r0);
                            tmem_ld_32x8(ta + 8, r1);
#pragma unroll
                            for (int j = 0; j < 8; ++j) { r[j] = r0[j]; r[8 + j] = r1[j]; }
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < C::NC; ++j) {
                            float v = __uint_as_float(r[j]);
                            if (s3) v = __shfl_down_sync(0xffffffffu, v, s3);
                            acc[j] += v;
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C::NC / 32; ++c) {
                        uint32_t r[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * C::BUF_COLS + col0 + c * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&cempty[buf]);
            }
#ifdef SCOUTER_PROF
            const long long _ts = clock64();
#endif
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < C::NC; ++j) acc[j] = fmaxf(acc[j], 0.f);
            }
            if (p.gap_part) {
                // Split-attention GAP partial sums (split_attn.py:64-66) while the tile is in registers: per 32 columns a
                // butterfly over the warp's 32 rows (31 shuffles: each step halves the columns a lane holds) leaves lane l with
                // the sum of column l.  Fixed order; the finish kernel adds the slots in order.  Rows outside the image add 0.
                if constexpr (C::NC % 32 == 0) {
                    const int mt_img = ((t / p.n_tiles) % p.m_tiles) % (p.tw * p.th);
                    float* gp = p.gap_part + ((size_t)b * p.gap_slots + mt_img * 4 + q) * p.Cout + ch0 + lane;
#pragma unroll
                    for (int c = 0; c < C::NC / 32; ++c) {
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = valid ? acc[c * 32 + j] : 0.f;
#pragma unroll
                        for (int off = 16, nn = 32; off >= 1; off >>= 1, nn >>= 1) {
                            const bool upper = (lane & off) != 0;
#pragma unroll
                            for (int j = 0; j < nn / 2; ++j) {
                                const float send = upper ? v[j] : v[j + nn / 2];
                                const float keep = upper ? v[j + nn / 2] : v[j];
                                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                        }
                        gp[c * 32] = v[0];
                    }
                }
            }
            if (p.tma_store) {
                // 16 columns of every tile pixel go through a swizzled staging buffer (rows = compact pixel index
                // hb*Wb + wb, the order of the TMA box) and leave as one 4-D bulk tensor store; pixels outside the
                // image are clipped by TMA.
                const bool inbox = hb < p.Hb && wb < p.Wb;
                const int cr = hb * p.Wb + wb;
#pragma unroll
                for (int c16 = 0; c16 < C::NC / 16; ++c16) {
                    uint8_t* stg = out_stage + (grp * C::OUT_BUFS + (c16 & (C::OUT_BUFS - 1))) * C::OUT_STAGE;
                    if (row == 0) { if constexpr (C::OUT_BUFS == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
                    named_bar_sync(1 + grp, 128);
                    if (inbox) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)   // SWIZZLE_64B: 16-byte chunk index ^= (row / 2) % 4
                            *reinterpret_cast<float4*>(stg + cr * 64 + ((j ^ ((cr >> 1) & 3)) << 4)) =
                                make_float4(acc[c16 * 16 + 4 * j], acc[c16 * 16 + 4 * j + 1], acc[c16 * 16 + 4 * j + 2], acc[c16 * 16 + 4 * j + 3]);
                    }
                    fence_proxy_async();
                    named_bar_sync(1 + grp, 128);
                    if (row == 0) {
                        tma_store_4d(&tmO, stg, ch0 + c16 * 16, w0, h0, b);
                        bulk_commit();
                    }
                }
            } else if (valid) {
                float* op = p.out + orow * p.Cout + ch0;
#pragma unroll
                for (int j = 0; j < C::NC / 4; ++j)
                    *reinterpret_cast<float4*>(op + 4 * j) = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            }
#ifdef SCOUTER_PROF
            prof_store += clock64() - _ts;
#endif
        }
        if (p.tma_store && row == 0) bulk_wait<0>();   // all bulk stores of this group have completed
        PROF_END(epi);
        if (threadIdx.x == 128) { PROF_STORE(g_prof_halo, 10, epi); PROF_STORE(g_prof_halo, 11, cfull); PROF_STORE(g_prof_halo, 12, store); }
    } else if (warp >= 8 && warp < 12) {
        // ===== splitters: fp16 / bf16 patches of A and the bf16 patch of A - fp16(A), once per patch =====
        const int tid = threadIdx.x - 256;
        const int prows = p.patch_bytes / 128;
        const SplitLane sl = split_lane(tid);
        int ps = 0;
        uint32_t pphase = 0;
        PROF_DECL(pfull); PROF_DECL(spl); PROF_BEGIN(spl);
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            for (int cb = 0; cb < p.cblocks; ++cb) {
                PROF_T(pfull, mbar_wait(&pfull[ps], pphase));
                uint8_t* raw = patch0 + ps * pslot;
                split_rows_hbr(raw, raw + p.patch_alloc, (uint32_t)p.patch_alloc / 2, sl, prows);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pready[ps]);
                if (++ps == p.pst) { ps = 0; pphase ^= 1; }
            }
        }
        PROF_END(spl);
        if (tid == 0) { PROF_STORE(g_prof_halo, 14, spl); PROF_STORE(g_prof_halo, 15, pfull); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn halo_encode_fn() {
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

// (Wb, Hb) with Hb*(Wb+2) <= 128 virtual rows: maximise the useful rows per 128-row MMA, and among (near-)equal
// choices take the tile whose halo patch is smallest relative to its output (square-ish tiles: fewer bytes per
// output and a small patch buffer).
double choose_halo_tile(int H, int W, int& Wb, int& Hb) {
    double best = -1.0, best_ratio = 1e30;
    Wb = Hb = 1;
    for (int wb = 1; wb <= W && wb + 2 <= 128; ++wb)
        for (int hb = 1; hb <= H && hb * (wb + 2) <= 128; ++hb) {
            const long long tiles = (long long)cdiv(W, wb) * cdiv(H, hb);
            const double util = (double)H * W / (tiles * 128.0);
            const double ratio = (double)(hb + 2) * (wb + 2) / ((double)hb * wb);
            if (util > best + 1e-3 || (util > best - 1e-3 && ratio < best_ratio)) {
                best = std::max(best, util);
                best_ratio = ratio;
                Wb = wb;
                Hb = hb;
            }
        }
    return best;
}

int halo_bn(int cout_g) {
    for (int bn : {128, 64, 32})
        if (cout_g % bn == 0) return bn;
    return 0;
}

template <int BN, bool FOLD = false>
int launch_halo_bn(const CUtensorMap& tA, const CUtensorMap& tB2, const CUtensorMap& tO, const HaloArgs& u,
                   int grid, int smem, cudaStream_t s) {
    SC_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<BN, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    conv3x3_halo_kernel<BN, FOLD><<<grid, 512, smem, s>>>(tA, tB2, tO, u);
    SC_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int halo_gap_slots(const ConvArgs& a) {
    if (!halo_conv_supported(a)) return 0;
    const int BN = halo_bn(a.Cout / a.groups);
    static bool no_fold = getenv("SCOUTER_HALO_NO_FOLD") != nullptr;
    const bool fold = BN <= 64 && !no_fold && a.W >= 14 && a.H >= 8;
    if (BN == 32) return 0;                       // the shuffle reduction works on 32-column groups per thread
    int wb, hb;
    choose_halo_tile(a.H, a.W, wb, hb);
    if (fold) { wb = 14; hb = 8; }
    return cdiv(a.W, wb) * cdiv(a.H, hb) * 4;
}

bool halo_conv_supported(const ConvArgs& a) {
    static bool off = getenv("SCOUTER_NO_HALO") != nullptr;
    if (off || !a.split || !a.w_rem) return false;
    if (a.stride != 1 || a.kh != 3 || a.kw != 3 || a.pad != 1) return false;
    if (a.Cin % a.groups || a.Cout % a.groups) return false;
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    if (cin_g % 32 || halo_bn(cout_g) == 0) return false;
    if ((long long)a.B * a.H * a.W >= (1ll << 31)) return false;
    int wb, hb;
    // small maps (7x7, 9x9) waste most of the 128 virtual rows: the tap-reload kernel packs several images per box
    return choose_halo_tile(a.H, a.W, wb, hb) >= 0.55;
}

int launch_conv_halo(const ConvArgs& a, UmmaConvPlan& plan, cudaStream_t s) {
    SC_CHECK_ARG(halo_conv_supported(a), SCOUTER_E_UNSUPPORTED, "conv_halo: unsupported geometry");
    EncodeTiledFn enc = halo_encode_fn();
    SC_CHECK_ARG(enc, SCOUTER_E_UNSUPPORTED, "conv_halo: cuTensorMapEncodeTiled is not available from the driver");
    const int cin_g = a.Cin / a.groups, cout_g = a.Cout / a.groups;
    const int BN = halo_bn(cout_g);
    HaloArgs u;
    u.bias = a.bias; u.res = a.res; u.out = a.out;
    u.B = a.B; u.H = a.H; u.W = a.W;
    u.gap_part = a.gap_part; u.gap_slots = a.gap_slots;
    SC_CHECK_ARG(!a.gap_part || a.gap_slots == halo_gap_slots(a), SCOUTER_E_INVALID, "conv_halo: gap_slots=%d, this geometry has %d",
                 a.gap_slots, halo_gap_slots(a));
    choose_halo_tile(a.H, a.W, u.Wb, u.Hb);
    // narrow outputs: taps folded into N (see the header); fixed 14 x 8 tiles = 8 patch lines of 16 rows
    static bool no_fold = getenv("SCOUTER_HALO_NO_FOLD") != nullptr;
    const bool fold = BN <= 64 && !no_fold && a.W >= 14 && a.H >= 8;
    if (fold) { u.Wb = 14; u.Hb = 8; }
    u.PW = u.Wb + 2;
    const int PH = u.Hb + 2;
    u.tw = cdiv(a.W, u.Wb); u.th = cdiv(a.H, u.Hb);
    u.m_tiles = u.tw * u.th * a.B;
    u.n_tiles = cout_g / BN;
    u.groups = a.groups; u.cin_g = cin_g; u.cout_g = cout_g; u.Cout = a.Cout;
    u.cblocks = cin_g / 32;
    u.relu = a.relu;
    u.patch_bytes = PH * u.PW * 128;
    // the last tap starts at row 2*PW+2 and the MMA reads 128 rows from there
    // multiple of 2048: a stage is 2.5 patch buffers and the next stage's SWIZZLE_128B patch must start 1024-aligned
    u.patch_alloc = (int)align_up((size_t)std::max(PH * u.PW, 2 * u.PW + 2 + 128) * 128, 2048);
    if (fold) u.patch_alloc = (int)align_up((size_t)PH * u.PW * 128, 2048);   // the last kernel row starts at 2*PW and reads 128 rows: exactly the patch
    static int chunk_kb = [] { const char* e = getenv("SCOUTER_UMMA_CHUNK"); int v = e ? atoi(e) : 3; return v < 1 ? 1 : v; }();
    u.chunk = fold ? 3 : chunk_kb;   // folded: one chunk = the three kernel rows of a channel block (18 accumulations)
    const int b_stage = (fold ? 3 : 1) * BN * 128;   // HCfg<BN, FOLD>::B_STAGE
    static bool no_tma_store = getenv("SCOUTER_NO_TMA_STORE") != nullptr;
    u.tma_store = no_tma_store ? 0 : 1;
    const int scratch = (BN == 128 ? 2 : 1) * 2 * 128 * 64;   // HCfg<BN>::OUT_BYTES
    const int budget = 226 * 1024 - 1536 - scratch;
    // narrow tiles (small BN) do little MMA work per patch and are latency/bandwidth bound: deeper patch prefetch
    const int want_pst = BN == 128 ? 2 : (BN == 64 ? 3 : 4);
    const int min_bst = fold ? 2 : 4;
    const int pslot = u.patch_alloc / 2 * 5;    // [raw | fp16 | bf16 | bf16 remainder]
    u.pst = 2;
    for (int pst = want_pst; pst >= 2; --pst)
        if ((budget - pst * pslot) / b_stage >= min_bst) { u.pst = pst; break; }
    u.bst = std::min(8, (budget - u.pst * pslot) / b_stage);
    u.wres = 0;
    if (fold) {
        static bool no_wres = getenv("SCOUTER_HALO_NO_WRES") != nullptr;
        const int w_bytes = u.cblocks * 3 * b_stage;
        const int fit = (budget - w_bytes) / pslot;
        if (!no_wres && fit >= 2) {
            u.wres = 1;
            u.pst = std::min(want_pst, fit);
            u.bst = u.cblocks * 3;
        }
    }
    SC_CHECK_ARG(u.bst >= 2, SCOUTER_E_UNSUPPORTED, "conv_halo: patch of %d bytes leaves no room for the weight ring", u.patch_alloc);
    const int smem = u.pst * pslot + u.bst * b_stage + 1024 + 512 + scratch;

    const bool reuse = plan.valid && plan.halo && plan.in == a.in && plan.w == a.w && plan.B == a.B && plan.H == a.H &&
                       plan.W == a.W && plan.Cin == a.Cin && plan.Cout == a.Cout && plan.groups == a.groups && plan.BN == BN && plan.out == a.out;
    if (!reuse) {
        cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t strides[3] = {(cuuint64_t)a.Cin * 4, (cuuint64_t)a.W * a.Cin * 4, (cuuint64_t)a.H * a.W * a.Cin * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)u.PW, (cuuint32_t)PH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&plan.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a.in, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_halo: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
        const cuuint64_t Kt = (cuuint64_t)9 * cin_g;
        cuuint32_t boxB[2] = {32, (cuuint32_t)BN};
        cuuint32_t esB[2] = {1, 1};
        // 16-bit [fp16 W_h ; bf16 W_r]: (2*Cout) rows of 9*cin_g elements (moved as raw 16-bit words)
        cuuint64_t dimsB2[2] = {Kt, (cuuint64_t)2 * a.Cout};
        cuuint64_t stridesB2[1] = {Kt * 2};
        r = enc(&plan.tmB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)a.w_rem, dimsB2, stridesB2, boxB, esB,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_halo: cuTensorMapEncodeTiled(bf16 W) failed with %d", (int)r);
        // output (Cout, W, H, B) fp32, box {16, Wb, Hb, 1}: one epilogue staging buffer = 16 channels of every tile pixel
        cuuint64_t dimsO[4] = {(cuuint64_t)a.Cout, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
        cuuint64_t stridesO[3] = {(cuuint64_t)a.Cout * 4, (cuuint64_t)a.W * a.Cout * 4, (cuuint64_t)a.H * a.W * a.Cout * 4};
        cuuint32_t boxO[4] = {16, (cuuint32_t)u.Wb, (cuuint32_t)u.Hb, 1};
        r = enc(&plan.tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a.out, dimsO, stridesO, boxO, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_halo: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
        plan.out = a.out;
        plan.valid = true; plan.halo = true;
        plan.in = a.in; plan.w = a.w; plan.B = a.B; plan.H = a.H; plan.W = a.W; plan.Cin = a.Cin; plan.Cout = a.Cout;
        plan.kh = 3; plan.groups = a.groups; plan.BN = BN;
    }
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        SC_CUDA(cudaGetDevice(&dev));
        SC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long total = (long long)u.m_tiles * u.n_tiles * u.groups;
    const int grid = (int)std::min<long long>(total, sms);
    switch (BN) {
        case 32: return fold ? launch_halo_bn<32, true>(plan.tmA, plan.tmB2, plan.tmO, u, grid, smem, s)
                             : launch_halo_bn<32>(plan.tmA, plan.tmB2, plan.tmO, u, grid, smem, s);
        case 64: return fold ? launch_halo_bn<64, true>(plan.tmA, plan.tmB2, plan.tmO, u, grid, smem, s)
                             : launch_halo_bn<64>(plan.tmA, plan.tmB2, plan.tmO, u, grid, smem, s);
        case 128: return launch_halo_bn<128>(plan.tmA, plan.tmB2, plan.tmO, u, grid, smem, s);
    }
    return SCOUTER_E_UNSUPPORTED;
}

}  // namespace scouter

#ifdef SCOUTER_PROF
// per-CTA role counters of the last halo launch: 32 slots per CTA (see scripts/prof_roles.py for the slot names)
extern "C" int scouter_prof_read_halo(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, scouter::g_prof_halo, sizeof(unsigned long long) * n);
}
#endif
