// 3x3 (stride 1, pad 1) convolution on tcgen05 from ONE halo tile per channel block -- error-compensated 3xTF32.
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
// Per (channel block, tap) the weights arrive as three TMA tiles: W in fp32 (kind::tf32 reads trunc19(W) exactly) and
// the host-made bf16 copies of W and of W_r = W - trunc19(W).  Four splitter warps turn each fp32 patch into bf16
// patches of A and A_r = A - trunc19(A) once per patch (not once per tap).  The issuer accumulates
// A_t*W_t (4 tf32 MMAs) + A*W_r + A_r*W (2 + 2 bf16 MMAs) in TMEM in chunks of `chunk` k-steps; epilogue threads merge
// the chunks in fp32 registers (the TMEM accumulator truncates on every MMA --
// profiles/r01_tmem_accumulator_truncation.txt).
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
    int patch_alloc;   // bytes reserved per patch buffer (raw or remainder), multiple of 1024
    int pst, bst;      // ring depths
    int chunk;         // k-steps (cb,tap pairs) per accumulation chunk
};

template <int BN>
struct HCfg {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int B_STAGE = 2 * B_BYTES;  // [W fp32 | bf16 W | bf16 W_r]
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int EPI_GROUPS = BN == 128 ? 2 : 1;
    static constexpr int NC = BN / EPI_GROUPS;
    static constexpr int THREADS = 512;
    static constexpr int MAX_ST = 8;
};

template <int BN>
__global__ void __launch_bounds__(512, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmB2, const HaloArgs p) {
    using C = HCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* patch0 = smem;                                      // pst x [raw | rem]
    uint8_t* bt0 = smem + p.pst * 2 * p.patch_alloc;            // bst x [W_t | W_r]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bt0 + p.bst * C::B_STAGE);
    uint64_t* pfull = bars;                  // [MAX_ST] patch landed (TMA)
    uint64_t* pready = pfull + C::MAX_ST;    // [MAX_ST] remainder written (128 splitter threads)
    uint64_t* pempty = pready + C::MAX_ST;   // [MAX_ST] all MMAs that read the patch retired
    uint64_t* bfull = pempty + C::MAX_ST;    // [MAX_ST]
    uint64_t* bempty = bfull + C::MAX_ST;    // [MAX_ST]
    uint64_t* cfull = bempty + C::MAX_ST;    // [2]
    uint64_t* cempty = cfull + 2;            // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(cempty + 2);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < C::MAX_ST; ++i) {
            mbar_init(&pfull[i], 1);
            mbar_init(&pready[i], 128);
            mbar_init(&pempty[i], 1);
            mbar_init(&bfull[i], 1);
            mbar_init(&bempty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&cfull[i], 1);
            mbar_init(&cempty[i], 128 * C::EPI_GROUPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int total = p.m_tiles * p.n_tiles * p.groups;
    const int ksteps = p.cblocks * 9;
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
            auto issue_patch = [&](int t, int cb) {
                int nt, g, w0, h0, b;
                tile_coords(t, nt, g, w0, h0, b);
                mbar_wait(&pempty[ps], pphase ^ 1);
                mbar_arrive_expect_tx(&pfull[ps], (uint32_t)p.patch_bytes);
                tma_load_4d(patch0 + ps * 2 * p.patch_alloc, &tmA, &pfull[ps], g * p.cin_g + cb * 32, w0 - 1, h0 - 1, b);
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
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                int nt, g, w0, h0, b;
                tile_coords(t, nt, g, w0, h0, b);
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    for (int tap = 0; tap < 9; ++tap) {
                        if (tap == 3) issue_next_patch();  // the issuer is inside this job: the oldest patch stage is (about to be) free
                        mbar_wait(&bempty[bs], bphase ^ 1);
                        uint8_t* sb = bt0 + bs * C::B_STAGE;
                        mbar_arrive_expect_tx(&bfull[bs], (uint32_t)C::B_STAGE);
                        const int kcol = tap * p.cin_g + cb * 32;
                        const int nrow = g * p.cout_g + nt * BN;
                        tma_load_2d(sb, &tmB, &bfull[bs], kcol, nrow);                                   // W fp32
                        tma_load_2d(sb + C::B_BYTES, &tmB2, &bfull[bs], kcol, nrow);                      // bf16 W
                        tma_load_2d(sb + C::B_BYTES + C::B_BYTES / 2, &tmB2, &bfull[bs], kcol, p.Cout + nrow);  // bf16 W_r
                        if (++bs == p.bst) { bs = 0; bphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = idesc_tf32(128, BN);
            int ps = 0, bs = 0;
            uint32_t pphase = 0, bphase = 0, cc = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                int step = 0;
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    mbar_wait(&pready[ps], pphase);
                    tc_fence_after();
                    const uint32_t pa = smem_u32(patch0 + ps * 2 * p.patch_alloc);
                    for (int tap = 0; tap < 9; ++tap, ++step) {
                        const int buf = cc & 1;
                        const bool chunk_start = step % p.chunk == 0;
                        if (chunk_start) {
                            mbar_wait(&cempty[buf], ((cc >> 1) & 1) ^ 1);
                            tc_fence_after();
                        }
                        mbar_wait(&bfull[bs], bphase);
                        tc_fence_after();
                        const int r = tap / 3, s = tap - 3 * r;
                        const uint32_t arow = (uint32_t)(r * p.PW + s);
                        const uint64_t da = smem_desc_sw128(pa + arow * 128u);
                        const uint64_t dab = smem_desc_sw64(pa + p.patch_alloc + arow * 64u);
                        const uint64_t darb = smem_desc_sw64(pa + p.patch_alloc + p.patch_alloc / 2 + arow * 64u);
                        const uint32_t sb = smem_u32(bt0 + bs * C::B_STAGE);
                        const uint64_t db = smem_desc_sw128(sb);
                        const uint64_t dwb = smem_desc_sw64(sb + C::B_BYTES);
                        const uint64_t dwrb = smem_desc_sw64(sb + C::B_BYTES + C::B_BYTES / 2);
                        const uint32_t d_tmem = tmem_base + buf * BN;
                        const uint32_t first = chunk_start ? 0u : 1u;
                        constexpr uint32_t idesc_b = idesc_bf16(128, BN);
#pragma unroll
                        for (int k = 0; k < 2; ++k) umma_bf16(d_tmem, dab + 2 * k, dwrb + 2 * k, idesc_b, first | k);  // A * W_r
#pragma unroll
                        for (int k = 0; k < 2; ++k) umma_bf16(d_tmem, darb + 2 * k, dwb + 2 * k, idesc_b, 1);          // A_r * W
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, 1);               // A_t * W_t
                        umma_commit(&bempty[bs]);
                        if (++bs == p.bst) { bs = 0; bphase ^= 1; }
                        if ((step + 1) % p.chunk == 0 || step + 1 == ksteps) {
                            umma_commit(&cfull[buf]);
                            ++cc;
                        }
                    }
                    umma_commit(&pempty[ps]);
                    if (++ps == p.pst) { ps = 0; pphase ^= 1; }
                }
            }
        }
    } else if ((warp >= 4 && warp < 8) || (C::EPI_GROUPS == 2 && warp >= 12)) {
        // ===== epilogue: thread = one virtual output row (hb*PW + wb') x NC columns =====
        const int q = warp & 3;
        const int grp = warp >= 12 ? 1 : 0;
        const int row = q * 32 + lane;
        const int col0 = grp * C::NC;
        const int hb = row / p.PW, wb = row - hb * p.PW;
        uint32_t cc = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            int nt, g, w0, h0, b;
            tile_coords(t, nt, g, w0, h0, b);
            const int h = h0 + hb, w = w0 + wb;
            const bool valid = hb < p.Hb && wb < p.Wb && h < p.H && w < p.W;
            const long long orow = ((long long)b * p.H + h) * p.W + w;
            const int ch0 = g * p.cout_g + nt * BN + col0;
            float acc[C::NC];
#pragma unroll
            for (int j = 0; j < C::NC; ++j) acc[j] = 0.f;
            for (int ch = 0; ch < nchunks; ++ch, ++cc) {
                const int buf = cc & 1;
                mbar_wait(&cfull[buf], (cc >> 1) & 1);
                tc_fence_after();
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
            }
            if (valid) {
                float* op = p.out + orow * p.Cout + ch0;
                const float* rp = p.res ? p.res + orow * p.Cout + ch0 : nullptr;
#pragma unroll
                for (int j = 0; j < C::NC / 4; ++j) {
                    float4 v = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                    if (p.bias) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + 4 * j));
                        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                    }
                    if (rp) {
                        const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + 4 * j));
                        v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
                    }
                    if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    *reinterpret_cast<float4*>(op + 4 * j) = v;
                }
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ===== splitters: bf16 patches of A and of A - trunc19(A), once per patch =====
        const int tid = threadIdx.x - 256;
        const int prows = p.patch_bytes / 128;
        int ps = 0;
        uint32_t pphase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            for (int cb = 0; cb < p.cblocks; ++cb) {
                mbar_wait(&pfull[ps], pphase);
                uint8_t* raw = patch0 + ps * 2 * p.patch_alloc;
                split_tile_bf16(raw, raw + p.patch_alloc, raw + p.patch_alloc + p.patch_alloc / 2, prows, tid, 128);
                fence_proxy_async();
                mbar_arrive(&pready[ps]);
                if (++ps == p.pst) { ps = 0; pphase ^= 1; }
            }
        }
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

template <int BN>
int launch_halo_bn(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tB2, const HaloArgs& u, int grid, int smem,
                   cudaStream_t s) {
    SC_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    conv3x3_halo_kernel<BN><<<grid, 512, smem, s>>>(tA, tB, tB2, u);
    SC_LAUNCH_CHECK();
    return 0;
}

}  // namespace

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
    choose_halo_tile(a.H, a.W, u.Wb, u.Hb);
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
    u.patch_alloc = (int)align_up((size_t)std::max(PH * u.PW, 2 * u.PW + 2 + 128) * 128, 1024);
    static int chunk_kb = [] { const char* e = getenv("SCOUTER_UMMA_CHUNK"); int v = e ? atoi(e) : 4; return v < 1 ? 1 : v; }();
    u.chunk = chunk_kb;
    const int b_stage = 2 * BN * 128;
    const int scratch = 0;
    const int budget = 224 * 1024 - 1536;
    // narrow tiles (small BN) do little MMA work per patch and are latency/bandwidth bound: deeper patch prefetch
    const int want_pst = BN == 128 ? 2 : (BN == 64 ? 3 : 4);
    u.pst = 2;
    for (int pst = want_pst; pst >= 2; --pst)
        if ((budget - pst * 2 * u.patch_alloc) / b_stage >= 4) { u.pst = pst; break; }
    u.bst = std::min(8, (budget - u.pst * 2 * u.patch_alloc) / b_stage);
    SC_CHECK_ARG(u.bst >= 2, SCOUTER_E_UNSUPPORTED, "conv_halo: patch of %d bytes leaves no room for the weight ring", u.patch_alloc);
    const int smem = u.pst * 2 * u.patch_alloc + u.bst * b_stage + 1024 + 512 + scratch;

    const bool reuse = plan.valid && plan.halo && plan.in == a.in && plan.w == a.w && plan.B == a.B && plan.H == a.H &&
                       plan.W == a.W && plan.Cin == a.Cin && plan.Cout == a.Cout && plan.groups == a.groups && plan.BN == BN;
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
        cuuint64_t dimsB[2] = {Kt, (cuuint64_t)a.Cout};
        cuuint64_t stridesB[1] = {Kt * 4};
        cuuint32_t boxB[2] = {32, (cuuint32_t)BN};
        cuuint32_t esB[2] = {1, 1};
        r = enc(&plan.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.w, dimsB, stridesB, boxB, esB, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_halo: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
        // bf16 [W ; W - trunc19(W)]: (2*Cout) rows of 9*cin_g bf16
        cuuint64_t dimsB2[2] = {Kt, (cuuint64_t)2 * a.Cout};
        cuuint64_t stridesB2[1] = {Kt * 2};
        r = enc(&plan.tmB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)a.w_rem, dimsB2, stridesB2, boxB, esB,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SC_CHECK_ARG(r == CUDA_SUCCESS, SCOUTER_E_UNSUPPORTED, "conv_halo: cuTensorMapEncodeTiled(bf16 W) failed with %d", (int)r);
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
        case 32: return launch_halo_bn<32>(plan.tmA, plan.tmB, plan.tmB2, u, grid, smem, s);
        case 64: return launch_halo_bn<64>(plan.tmA, plan.tmB, plan.tmB2, u, grid, smem, s);
        case 128: return launch_halo_bn<128>(plan.tmA, plan.tmB, plan.tmB2, u, grid, smem, s);
    }
    return SCOUTER_E_UNSUPPORTED;
}

}  // namespace scouter
