// xSlot head, exact-fp32 CUDA-core path (SCOUTER_MATH_FP32) + the small kernels around it.
//
// Replaces, for eval-mode forward:
//   sloter/utils/position_encode.py:26-46     -> pe_sine_kernel          (a7)
//   sloter/utils/slot_attention.py:44-96      -> xslot_loop_kernel       (a9)   one CTA per image
//   sloter/utils/slot_attention.py:68-80      -> vis_maps_kernel         (a10)
//   sloter/slot_model.py:117-125 + :93-96     -> head_finalize_kernel    (a11)
// The math follows SURVEY.md A.1 literally: no eps in the sum-normalisation, IEEE division,
// updates divided by d (not n), GRU gate order [r|z|n], third GRU step skipped (its result is dead).
#include "xslot.cuh"

namespace scouter {

int validate_xslot_desc(const scouter_xslot_desc_t* d) {
    SC_CHECK_ARG(d != nullptr, SCOUTER_E_INVALID, "xslot: desc is NULL");
    SC_CHECK_ARG(d->d == XD, SCOUTER_E_UNSUPPORTED, "xslot: hidden dim %d (only %d is implemented)", d->d, XD);
    SC_CHECK_ARG(d->num_classes >= 1 && d->slots_per_class >= 1, SCOUTER_E_INVALID,
                 "xslot: num_classes=%d slots_per_class=%d", d->num_classes, d->slots_per_class);
    SC_CHECK_ARG(d->to_k_layers >= 1 && d->to_k_layers <= SCOUTER_MAX_TO_K_LAYERS, SCOUTER_E_UNSUPPORTED,
                 "xslot: to_k_layers=%d (1..%d)", d->to_k_layers, SCOUTER_MAX_TO_K_LAYERS);
    SC_CHECK_ARG(d->iters >= 1, SCOUTER_E_INVALID, "xslot: iters=%d", d->iters);
    SC_CHECK_ARG(d->loss_status == 1 || d->loss_status == -1, SCOUTER_E_INVALID, "xslot: loss_status=%d", d->loss_status);
    return 0;
}

namespace {

// ------------------------------------------------------------------------------------------------
// a7: sine position table, token-major (n, d).  fp32 op order mirrors the reference so the table
// agrees to ~1e-7: embed = (idx+1) / (size + 1e-6) * 2pi ; pos = embed / T^(2*floor(c/2)/(d/2)).
// ------------------------------------------------------------------------------------------------
__global__ void pe_sine_kernel(float* __restrict__ pe, int d, int h, int w) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= h * w * d) return;
    int c = idx % d, j = idx / d;
    int y = j / w, x = j % w;
    int half = d / 2;
    bool is_x = c >= half;
    int cc = is_x ? c - half : c;
    float coord = is_x ? (float)(x + 1) : (float)(y + 1);
    float last = is_x ? (float)w : (float)h;
    float embed = coord / (last + 1e-6f) * 6.283185307179586f;
    float expo = (float)(2 * (cc / 2)) / (float)half;
    float dim_t = powf(10000.0f, expo);
    float pos = embed / dim_t;
    pe[idx] = (cc & 1) ? cosf(pos) : sinf(pos);
}

// ------------------------------------------------------------------------------------------------
// Parameter packing (runs once per parameter version).
// ------------------------------------------------------------------------------------------------
__global__ void xslot_pack_kernel(scouter_xslot_desc_t d, float* __restrict__ out) {
    const int S = d.num_classes * d.slots_per_class;
    XSlotPacked pk{S, d.to_k_layers};
    size_t total = pk.total();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float v;
        if (i < pk.tok_wt(0)) {
            v = d.initial_slots[i];
        } else if (i < pk.gru_wih_t()) {
            size_t r = i - pk.tok_wt(0);
            int l = (int)(r / (XD * XD + XD));
            size_t q = r - (size_t)l * (XD * XD + XD);
            if (q < XD * XD) {
                int e = (int)(q / XD), o = (int)(q % XD);
                v = d.to_k_w[l][o * XD + e];
            } else {
                v = d.to_k_b[l][q - XD * XD];
            }
        } else if (i < pk.gru_whh_t()) {
            size_t q = i - pk.gru_wih_t();
            int e = (int)(q / XG), g = (int)(q % XG);
            v = d.gru_w_ih[g * XD + e];
        } else if (i < pk.gru_bih()) {
            size_t q = i - pk.gru_whh_t();
            int e = (int)(q / XG), g = (int)(q % XG);
            v = d.gru_w_hh[g * XD + e];
        } else if (i < pk.gru_bhh()) {
            v = d.gru_b_ih[i - pk.gru_bih()];
        } else if (i < pk.tok_w_raw(0)) {
            v = d.gru_b_hh[i - pk.gru_bhh()];
        } else if (i < pk.tok_w_pair(0)) {
            size_t r = i - pk.tok_w_raw(0);
            v = d.to_k_w[r / (XD * XD)][r % (XD * XD)];
        } else {
            // two bf16 per float slot: element index eb = 2*(i - pair(0)) + {0,1}; row = eb / 64 (0..127 per layer), col = eb % 64
            size_t r = i - pk.tok_w_pair(0);
            int l = (int)(r / (XD * XD));
            size_t eb = 2 * (r - (size_t)l * XD * XD);
            int row = (int)(eb / XD), col = (int)(eb % XD);
            float w0 = d.to_k_w[l][(row & 63) * XD + col], w1 = d.to_k_w[l][(row & 63) * XD + col + 1];
            if (row >= XD) {
                w0 -= __uint_as_float(__float_as_uint(w0) & 0xFFFFE000u);
                w1 -= __uint_as_float(__float_as_uint(w1) & 0xFFFFE000u);
            }
            uint32_t pr;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pr) : "f"(w1), "f"(w0));
            v = __uint_as_float(pr);
        }
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// a9: the slot-attention loop.  One CTA (256 threads) per image; the image's tokens, keys and slots
// live in shared memory for the whole kernel.  Slots are processed in chunks of SC so S = 400
// (CUB-200 x 2 slots) fits: pass 1 gets every row sum r_i and the total t, pass 2 re-derives the
// chunk's dots (only when there is more than one chunk), forms the attention, the update and the GRU.
// ------------------------------------------------------------------------------------------------
#ifndef SCOUTER_XSLOT_NT
#define SCOUTER_XSLOT_NT 512
#endif
constexpr int NT = SCOUTER_XSLOT_NT, NW = NT / 32, NRG = NT / XD;
constexpr int LDX = XD + 4;

// Kb holds the MLP ping-pong buffer first and {dots[SC][n] (rounded up to 4 floats), upd[SC][64]} afterwards.
__host__ __device__ inline int kb_floats(int n, int SC) {
    int a = n * LDX, b = ((SC * n + 3) & ~3) + SC * XD;
    return ((a > b ? a : b) + 3) & ~3;
}

struct LoopArgs {
    const float* packed;
    const float* x;
    long long x_sb, x_sn, x_sd;
    const float* xpe;
    long long p_sb, p_sn, p_sd;
    const float* pe;
    float* logits;
    float* attn;
    float* attn_sum;
    int n, S, C, spc, L, iters, loss_status, SC;
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

__global__ void __launch_bounds__(NT) xslot_loop_kernel(LoopArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int n = a.n, S = a.S, SC = a.SC;
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int b = blockIdx.x;
    const XSlotPacked pk{S, a.L};

    float* Xs = smem;
    float* Ka = Xs + n * LDX;
    float* Kb = Ka + n * LDX;
    float* slots = Kb + kb_floats(n, SC);
    float* rsum = slots + S * XD;
    float* usum = rsum + S;
    float* misc = usum + S;  // [NW + 2]

    // ---- load tokens ---------------------------------------------------------------------------
    float* kin = (a.L & 1) ? Kb : Ka;  // after L ping-pong layers the keys end up in Ka
    float* kout = (a.L & 1) ? Ka : Kb;
    {
        const float* xb = a.x + (long long)b * a.x_sb;
        const float* pb = a.xpe ? a.xpe + (long long)b * a.p_sb : nullptr;
        const bool jfast = a.x_sd != 1;  // permuted (B,d,n) views: walk j fastest to stay coalesced
        for (int idx = tid; idx < n * XD; idx += NT) {
            int j, e;
            if (jfast) { e = idx / n; j = idx - e * n; } else { j = idx / XD; e = idx - j * XD; }
            float xv = xb[j * a.x_sn + e * a.x_sd];
            float pv = pb ? pb[j * a.p_sn + e * a.p_sd] : xv + __ldg(a.pe + j * XD + e);
            Xs[j * LDX + e] = xv;
            kin[j * LDX + e] = pv;
        }
        for (int idx = tid; idx < S * XD; idx += NT) slots[idx] = __ldg(a.packed + pk.slots() + idx);
    }
    __syncthreads();

    // ---- to_k MLP (slot_attention.py:30-37,47): thread = output feature o, 4 token rows at a time --
    {
        const int o = tid % XD, rg = tid / XD;
        for (int l = 0; l < a.L; ++l) {
            float wreg[XD];
#pragma unroll
            for (int e = 0; e < XD; ++e) wreg[e] = __ldg(a.packed + pk.tok_wt(l) + e * XD + o);
            const float bias = __ldg(a.packed + pk.tok_b(l) + o);
            const bool relu = l + 1 < a.L;
            for (int j0 = rg * 4; j0 < n; j0 += NRG * 4) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                const float* r0 = kin + min(j0 + 0, n - 1) * LDX;
                const float* r1 = kin + min(j0 + 1, n - 1) * LDX;
                const float* r2 = kin + min(j0 + 2, n - 1) * LDX;
                const float* r3 = kin + min(j0 + 3, n - 1) * LDX;
#pragma unroll
                for (int e4 = 0; e4 < XD / 4; ++e4) {
                    float4 v0 = *reinterpret_cast<const float4*>(r0 + e4 * 4);
                    float4 v1 = *reinterpret_cast<const float4*>(r1 + e4 * 4);
                    float4 v2 = *reinterpret_cast<const float4*>(r2 + e4 * 4);
                    float4 v3 = *reinterpret_cast<const float4*>(r3 + e4 * 4);
                    acc[0] = fmaf(v0.x, wreg[e4 * 4 + 0], acc[0]); acc[1] = fmaf(v1.x, wreg[e4 * 4 + 0], acc[1]);
                    acc[2] = fmaf(v2.x, wreg[e4 * 4 + 0], acc[2]); acc[3] = fmaf(v3.x, wreg[e4 * 4 + 0], acc[3]);
                    acc[0] = fmaf(v0.y, wreg[e4 * 4 + 1], acc[0]); acc[1] = fmaf(v1.y, wreg[e4 * 4 + 1], acc[1]);
                    acc[2] = fmaf(v2.y, wreg[e4 * 4 + 1], acc[2]); acc[3] = fmaf(v3.y, wreg[e4 * 4 + 1], acc[3]);
                    acc[0] = fmaf(v0.z, wreg[e4 * 4 + 2], acc[0]); acc[1] = fmaf(v1.z, wreg[e4 * 4 + 2], acc[1]);
                    acc[2] = fmaf(v2.z, wreg[e4 * 4 + 2], acc[2]); acc[3] = fmaf(v3.z, wreg[e4 * 4 + 2], acc[3]);
                    acc[0] = fmaf(v0.w, wreg[e4 * 4 + 3], acc[0]); acc[1] = fmaf(v1.w, wreg[e4 * 4 + 3], acc[1]);
                    acc[2] = fmaf(v2.w, wreg[e4 * 4 + 3], acc[2]); acc[3] = fmaf(v3.w, wreg[e4 * 4 + 3], acc[3]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (j0 + r < n) {
                        float v = acc[r] + bias;
                        kout[(j0 + r) * LDX + o] = relu ? fmaxf(v, 0.f) : v;
                    }
                }
            }
            __syncthreads();
            float* t = kin; kin = kout; kout = t;
        }
    }
    const float* K = Ka;        // == kin after the loop
    float* dots = Kb;           // [SC][n]
    float* upd = Kb + ((SC * n + 3) & ~3);  // [SC][64], 16-byte aligned for the float4 reads in the GRU

    const int nchunks = (S + SC - 1) / SC;
    float asum_local = 0.f;

    auto compute_dots = [&](int c0, int sc) {
        for (int idx = tid; idx < sc * n; idx += NT) {
            int i = idx / n, j = idx - i * n;
            const float4* sp = reinterpret_cast<const float4*>(slots + (c0 + i) * XD);
            const float4* kp = reinterpret_cast<const float4*>(K + j * LDX);
            float acc = 0.f;
#pragma unroll
            for (int e4 = 0; e4 < XD / 4; ++e4) {
                float4 s4 = sp[e4], k4 = kp[e4];
                acc = fmaf(s4.x, k4.x, acc); acc = fmaf(s4.y, k4.y, acc);
                acc = fmaf(s4.z, k4.z, acc); acc = fmaf(s4.w, k4.w, acc);
            }
            dots[idx] = acc * 0.125f;  // scale = d^-1/2 (slot_attention.py:16,55)
        }
    };

    for (int it = 0; it < a.iters; ++it) {
        const bool last = it == a.iters - 1;
        // ---- pass 1: row sums r_i of every slot, then the image total t (:56) -----------------------
        for (int c0 = 0; c0 < S; c0 += SC) {
            const int sc = min(SC, S - c0);
            compute_dots(c0, sc);
            __syncthreads();
            for (int i = warp; i < sc; i += NW) {
                float s = 0.f;
                for (int j = lane; j < n; j += 32) s += dots[i * n + j];
                s = warp_sum(s);
                if (lane == 0) rsum[c0 + i] = s;
            }
            __syncthreads();
        }
        if (warp == 0) {
            float s = 0.f;
            for (int i = lane; i < S; i += 32) s += rsum[i];
            s = warp_sum(s);
            if (lane == 0) misc[0] = s;
        }
        __syncthreads();
        const float tot = misc[0];

        // ---- pass 2 per chunk: attention, update, GRU -------------------------------------------------
        for (int c0 = 0; c0 < S; c0 += SC) {
            const int sc = min(SC, S - c0);
            if (nchunks > 1) {
                compute_dots(c0, sc);
                __syncthreads();
            }
            for (int idx = tid; idx < sc * n; idx += NT) {
                int i = idx / n;
                float v = dots[idx] / rsum[c0 + i] * tot;  // (D / r) * t, IEEE, no eps (:56)
                float at = sigmoidf_(v);                   // (:57)
                dots[idx] = at;
                if (last) {
                    asum_local += at;
                    if (a.attn) a.attn[((long long)b * S + c0) * n + idx] = at;
                }
            }
            __syncthreads();
            // updates = attn @ X / d  (:58-59)
            {
                const int e = tid % XD, ig = tid / XD;
                for (int i0 = ig * 4; i0 < sc; i0 += NRG * 4) {
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    const float* d0 = dots + min(i0 + 0, sc - 1) * n;
                    const float* d1 = dots + min(i0 + 1, sc - 1) * n;
                    const float* d2 = dots + min(i0 + 2, sc - 1) * n;
                    const float* d3 = dots + min(i0 + 3, sc - 1) * n;
                    for (int j = 0; j < n; ++j) {
                        float xv = Xs[j * LDX + e];
                        acc[0] = fmaf(d0[j], xv, acc[0]);
                        acc[1] = fmaf(d1[j], xv, acc[1]);
                        acc[2] = fmaf(d2[j], xv, acc[2]);
                        acc[3] = fmaf(d3[j], xv, acc[3]);
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (i0 + r < sc) upd[(i0 + r) * XD + e] = acc[r] * (1.0f / XD);
                }
            }
            __syncthreads();
            if (last) {
                // logits need only sum_e updates (:96)
                for (int i = warp; i < sc; i += NW) {
                    float s = upd[i * XD + lane] + upd[i * XD + 32 + lane];
                    s = warp_sum(s);
                    if (lane == 0) usum[c0 + i] = s;
                }
            } else {
                // GRU cell (:60-66), gate order [r|z|n]; thread = hidden unit e, 4 slot rows at a time.
                const int e = tid % XD, ig = tid / XD;
                const float* wih = a.packed + pk.gru_wih_t();
                const float* whh = a.packed + pk.gru_whh_t();
                const float bir = __ldg(a.packed + pk.gru_bih() + e), biz = __ldg(a.packed + pk.gru_bih() + XD + e),
                            bin = __ldg(a.packed + pk.gru_bih() + 2 * XD + e);
                const float bhr = __ldg(a.packed + pk.gru_bhh() + e), bhz = __ldg(a.packed + pk.gru_bhh() + XD + e),
                            bhn = __ldg(a.packed + pk.gru_bhh() + 2 * XD + e);
                const int nblk = (((sc + 3) / 4) + NRG - 1) / NRG;
                for (int blk = 0; blk < nblk; ++blk) {
                    const int i0 = (ig + blk * NRG) * 4;
                    const bool active = i0 < sc;
                    float gi[3][4], gh[3][4];
#pragma unroll
                    for (int g = 0; g < 3; ++g)
#pragma unroll
                        for (int r = 0; r < 4; ++r) gi[g][r] = gh[g][r] = 0.f;
                    float hnew[4] = {0.f, 0.f, 0.f, 0.f};
                    if (active) {
                        const float* up[4];
                        const float* sp[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            int i = min(i0 + r, sc - 1);
                            up[r] = upd + i * XD;
                            sp[r] = slots + (c0 + i) * XD;
                        }
                        for (int k4 = 0; k4 < XD / 4; ++k4) {
                            float4 u4[4], h4[4];
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                u4[r] = *reinterpret_cast<const float4*>(up[r] + k4 * 4);
                                h4[r] = *reinterpret_cast<const float4*>(sp[r] + k4 * 4);
                            }
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const int k = k4 * 4 + kk;
                                float wi0 = __ldg(wih + k * XG + e), wi1 = __ldg(wih + k * XG + XD + e),
                                      wi2 = __ldg(wih + k * XG + 2 * XD + e);
                                float wh0 = __ldg(whh + k * XG + e), wh1 = __ldg(whh + k * XG + XD + e),
                                      wh2 = __ldg(whh + k * XG + 2 * XD + e);
#pragma unroll
                                for (int r = 0; r < 4; ++r) {
                                    float u = kk == 0 ? u4[r].x : kk == 1 ? u4[r].y : kk == 2 ? u4[r].z : u4[r].w;
                                    float h = kk == 0 ? h4[r].x : kk == 1 ? h4[r].y : kk == 2 ? h4[r].z : h4[r].w;
                                    gi[0][r] = fmaf(u, wi0, gi[0][r]);
                                    gi[1][r] = fmaf(u, wi1, gi[1][r]);
                                    gi[2][r] = fmaf(u, wi2, gi[2][r]);
                                    gh[0][r] = fmaf(h, wh0, gh[0][r]);
                                    gh[1][r] = fmaf(h, wh1, gh[1][r]);
                                    gh[2][r] = fmaf(h, wh2, gh[2][r]);
                                }
                            }
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            float rg_ = sigmoidf_((gi[0][r] + bir) + (gh[0][r] + bhr));
                            float zg = sigmoidf_((gi[1][r] + biz) + (gh[1][r] + bhz));
                            float ng = tanhf((gi[2][r] + bin) + rg_ * (gh[2][r] + bhn));
                            float hp = sp[r][e];
                            hnew[r] = (hp - ng) * zg + ng;  // ATen's form of (1-z)*n + z*h
                        }
                    }
                    __syncthreads();  // every read of the old slot rows of this block is done
                    if (active) {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (i0 + r < sc) slots[(c0 + i0 + r) * XD + e] = hnew[r];
                    }
                }
            }
            __syncthreads();
        }
    }

    // ---- outputs -------------------------------------------------------------------------------------
    for (int c = tid; c < a.C; c += NT) {
        float s = 0.f;
        for (int m = 0; m < a.spc; ++m) s += usum[c * a.spc + m];  // class = consecutive slots (:87-91)
        a.logits[(long long)b * a.C + c] = (float)a.loss_status * s;
    }
    if (a.attn_sum) {
        float s = warp_sum(asum_local);
        if (lane == 0) misc[1 + warp] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < NW; ++i) t += misc[1 + i];
            a.attn_sum[b] = t;
        }
    }
}

size_t loop_smem_bytes(int n, int S, int SC) {
    size_t fl = (size_t)2 * n * LDX + kb_floats(n, SC) + (size_t)S * XD + 2 * (size_t)S + NW + 8;
    return fl * sizeof(float);
}

// ------------------------------------------------------------------------------------------------
// a11 + batch part of a9.  One CTA.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_finalize_kernel(const float* __restrict__ logits,
                                                            const float* __restrict__ attn_sum,
                                                            const int64_t* __restrict__ target, int B, int C, int S,
                                                            int n, float power, float lambda_value,
                                                            float* __restrict__ log_probs, float* __restrict__ losses) {
    __shared__ float red[256];
    const int tid = threadIdx.x;
    float nll_local = 0.f;
    for (int b = tid; b < B; b += 256) {
        const float* row = logits + (long long)b * C;
        float mx = -INFINITY;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, row[c]);
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
        float lse = logf(se);
        for (int c = 0; c < C; ++c) log_probs[(long long)b * C + c] = row[c] - mx - lse;
        if (target) {
            long long t = target[b];
            if (t >= 0 && t < C) nll_local -= row[t] - mx - lse;
        }
    }
    if (!losses) return;
    // fixed-order tree reductions -> bit-reproducible scalars
    red[tid] = nll_local;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    float nll = red[0] / (float)B;
    __syncthreads();
    float as = 0.f;
    if (attn_sum)
        for (int b = tid; b < B; b += 256) as += attn_sum[b];
    red[tid] = as;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        // slot_attention.py:94: sum / B / S / n, then pow (:96)
        float m = red[0] / (float)B / (float)S / (float)n;
        float al = power == 1.f ? m : (power == 2.f ? m * m : powf(m, power));
        losses[2] = al;
        if (target) {
            losses[1] = nll;
            losses[0] = nll + lambda_value * al;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// a10: uint8 explanation maps of one image.  One CTA; values recomputed in the second sweep.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vis_maps_kernel(const float* __restrict__ attn, int C, int spc, int n,
                                                       uint8_t* __restrict__ maps) {
    __shared__ float rmin[8], rmax[8];
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    auto val = [&](int idx) {
        int c = idx / n, j = idx - c * n;
        float s = 0.f;
        for (int m = 0; m < spc; ++m) s = __fadd_rn(s, attn[(long long)(c * spc + m) * n + j]);
        return s;
    };
    float mn = INFINITY, mx = -INFINITY;
    for (int idx = tid; idx < C * n; idx += 256) {
        float v = val(idx);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) { rmin[warp] = mn; rmax[warp] = mx; }
    __syncthreads();
    mn = rmin[0]; mx = rmax[0];
    for (int i = 1; i < 8; ++i) { mn = fminf(mn, rmin[i]); mx = fmaxf(mx, rmax[i]); }
    for (int idx = tid; idx < C * n; idx += 256) {
        // byte work is held to bit-exactness: IEEE sub / div / mul in the reference's order, nothing contracted
        const float v = __fmul_rn(__fdiv_rn(__fsub_rn(val(idx), mn), __fsub_rn(mx, mn)), 255.0f);
        maps[idx] = (uint8_t)v;  // numpy astype(uint8) truncates
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host entry points
// ------------------------------------------------------------------------------------------------
static int pick_slot_chunk(int S) { return S <= 64 ? S : 64; }

int xslot_loop_launch(const scouter_xslot_desc_t* d, const void* packed, const scouter_xslot_io_t* io, cudaStream_t s) {
    const int S = d->num_classes * d->slots_per_class;
    LoopArgs a;
    a.packed = (const float*)packed;
    a.x = io->x; a.x_sb = io->x_sb; a.x_sn = io->x_sn; a.x_sd = io->x_sd;
    a.xpe = io->x_pe; a.p_sb = io->xpe_sb; a.p_sn = io->xpe_sn; a.p_sd = io->xpe_sd;
    a.pe = io->pe;
    a.logits = io->logits; a.attn = io->attn; a.attn_sum = io->attn_sum;
    a.n = io->n; a.S = S; a.C = d->num_classes; a.spc = d->slots_per_class; a.L = d->to_k_layers;
    a.iters = d->iters; a.loss_status = d->loss_status; a.SC = pick_slot_chunk(S);
    size_t smem = loop_smem_bytes(a.n, S, a.SC);
    SC_CHECK_ARG(smem <= 227 * 1024, SCOUTER_E_UNSUPPORTED,
                 "xslot: n=%d, S=%d needs %zu bytes of shared memory per image (limit 227 KB)", a.n, S, smem);
    SC_CUDA(cudaFuncSetAttribute(xslot_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xslot_loop_kernel<<<io->batch, NT, smem, s>>>(a);
    SC_LAUNCH_CHECK();
    return 0;
}

}  // namespace scouter

using namespace scouter;

extern "C" int scouter_pe_sine(float* pe, int d, int h, int w, scouter_stream_t stream) {
    SC_CHECK_ARG(pe && d > 0 && d % 4 == 0 && h > 0 && w > 0, SCOUTER_E_INVALID, "pe_sine: bad arguments d=%d h=%d w=%d", d, h, w);
    int total = h * w * d;
    pe_sine_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(pe, d, h, w);
    SC_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t scouter_xslot_packed_bytes(const scouter_xslot_desc_t* desc) {
    if (validate_xslot_desc(desc)) return 0;
    XSlotPacked pk{desc->num_classes * desc->slots_per_class, desc->to_k_layers};
    return pk.total() * sizeof(float);
}

extern "C" int scouter_xslot_pack(const scouter_xslot_desc_t* desc, void* packed, scouter_stream_t stream) {
    if (int e = validate_xslot_desc(desc)) return e;
    SC_CHECK_ARG(packed, SCOUTER_E_INVALID, "xslot_pack: packed is NULL");
    SC_CHECK_ARG(desc->initial_slots && desc->gru_w_ih && desc->gru_w_hh && desc->gru_b_ih && desc->gru_b_hh,
                 SCOUTER_E_INVALID, "xslot_pack: NULL parameter pointer");
    for (int l = 0; l < desc->to_k_layers; ++l)
        SC_CHECK_ARG(desc->to_k_w[l] && desc->to_k_b[l], SCOUTER_E_INVALID, "xslot_pack: to_k layer %d is NULL", l);
    xslot_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(*desc, (float*)packed);
    SC_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t scouter_xslot_workspace_bytes(const scouter_xslot_desc_t*, int, int) { return 0; }

extern "C" int scouter_xslot_forward(const scouter_xslot_desc_t* desc, const void* packed, const scouter_xslot_io_t* io,
                                     void*, size_t, scouter_stream_t stream) {
    if (int e = validate_xslot_desc(desc)) return e;
    SC_CHECK_ARG(packed && io, SCOUTER_E_INVALID, "xslot_forward: NULL packed/io");
    SC_CHECK_ARG(io->batch > 0 && io->n > 0, SCOUTER_E_INVALID, "xslot_forward: batch=%d n=%d", io->batch, io->n);
    SC_CHECK_ARG(io->x && io->logits, SCOUTER_E_INVALID, "xslot_forward: x / logits is NULL");
    SC_CHECK_ARG(io->x_pe || io->pe, SCOUTER_E_INVALID, "xslot_forward: give x_pe or the pe table");
    if (!io->x_pe && io->x_sd == 1 && io->x_sn == XD && io->x_sb == (int64_t)io->n * XD && xslot_fast_supported(desc, io->n)) {
        XSlotFastIO f;
        f.batch = io->batch; f.n = io->n; f.x = io->x; f.xpart = nullptr; f.conv_bias = nullptr; f.split_stride = 0; f.nsplit = 0;
        f.pe = io->pe; f.x_out = nullptr; f.logits = io->logits; f.attn = io->attn; f.attn_sum = io->attn_sum;
        return xslot_fast_launch(desc, packed, f, (cudaStream_t)stream);
    }
    return xslot_loop_launch(desc, packed, io, (cudaStream_t)stream);
}

extern "C" int scouter_head_finalize(const float* logits, const float* attn_sum, const int64_t* target, int batch,
                                     int num_classes, int num_slots, int n, float power, float lambda_value,
                                     float* log_probs, float* losses, scouter_stream_t stream) {
    SC_CHECK_ARG(logits && log_probs && batch > 0 && num_classes > 0, SCOUTER_E_INVALID, "head_finalize: bad arguments");
    head_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, attn_sum, target, batch, num_classes, num_slots, n,
                                                             power, lambda_value, log_probs, losses);
    SC_LAUNCH_CHECK();
    return 0;
}

extern "C" int scouter_vis_maps_u8(const float* attn, int batch, int num_classes, int slots_per_class, int n, int vis_id,
                                   uint8_t* maps, scouter_stream_t stream) {
    SC_CHECK_ARG(attn && maps && vis_id >= 0 && vis_id < batch, SCOUTER_E_INVALID, "vis_maps: vis_id=%d batch=%d", vis_id, batch);
    const float* a = attn + (long long)vis_id * num_classes * slots_per_class * n;
    vis_maps_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a, num_classes, slots_per_class, n, maps);
    SC_LAUNCH_CHECK();
    return 0;
}
