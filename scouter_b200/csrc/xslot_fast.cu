// xSlot loop, throughput version for the common regime (S <= 32 slots, to_k_layers <= 5): IMG images per CTA processed
// side by side, 512 threads, every weight matrix staged in shared memory (to_k first; the GRU block replaces it
// with cp.async while the first attention pass runs), fp32 FMA throughout (the sum-normalisation needs
// fp32-accurate to_k / QK^T / GRU), fixed-order reductions (bit-reproducible).
//
// Same math as xslot_loop_kernel (xslot.cu), i.e. sloter/utils/slot_attention.py:44-96; that kernel stays the
// general path (S up to ~600 via slot chunking).  When the projection ran split-K, the token load also finishes it:
// x = relu(sum_s partial_s + conv1x1.bias)  (sloter/slot_model.py:108-109).
#include <cstdlib>

#include "xslot.cuh"

namespace scouter {
namespace {

constexpr int FT = 512, FW = FT / 32;
constexpr int LDX = XD + 4;
constexpr int W_FLOATS = 2 * XD * XG + 2 * XG;   // GRU block: WihT[64][192], WhhT[64][192], b_ih[192], b_hh[192]
constexpr int TOK_FLOATS = XD * XD + XD;         // one to_k layer: WT[64][64], b[64]

struct FastArgs {
    const float* packed;
    const float* x;          // (B, n, 64) contiguous, or null when xpart is given
    const float* xpart;      // (nsplit, B*n, 64) partial projections, or null
    const float* conv_bias;  // (64) with xpart
    long long split_stride;  // floats between partial slabs
    int nsplit;
    const float* pe;         // (n, 64)
    float* x_out;            // optional (B, n, 64)
    float* logits;
    float* attn;
    float* attn_sum;
    int B, n, S, C, spc, L, iters, loss_status;
};

// Kb holds the MLP ping-pong buffer first, {gates[SR][384], attnT[IMG][n][SP]} afterwards.
__host__ __device__ inline int kb_floats(int img, int n, int S) {
    const int a = img * n * LDX, b = img * S * 2 * XG + img * n * ((S + 3) & ~3);
    return ((a > b ? a : b) + 3) & ~3;
}

__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + expf(-v)); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int IMG>
__global__ void __launch_bounds__(FT, 1) xslot_fast_kernel(FastArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int n = a.n, S = a.S;
    const int R = IMG * n;          // token rows in this CTA
    const int SR = IMG * S;         // slot rows
    const int tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int b0 = blockIdx.x * IMG;
    const int nimg = min(IMG, a.B - b0);
    const XSlotPacked pk{S, a.L};

    const int SP = (S + 3) & ~3;
    float* Xs = sm;                          // [R][LDX]
    float* Ka = Xs + R * LDX;                // [R][LDX]
    float* Kb = Ka + R * LDX;                // [R][LDX]  MLP ping-pong; afterwards gates / attention scratch
    float* Wsm = Kb + kb_floats(IMG, n, S);  // [W_FLOATS] to_k layers, then the GRU block
    float* slots = Wsm + W_FLOATS;           // [SR][64]
    float* upd = slots + SR * XD;            // [SR][64]
    float* rsum = upd + SR * XD;             // [SR]
    float* usum = rsum + SR;                 // [SR]
    float* misc = usum + SR;                 // [IMG + FW + 4]
    // views into Kb after the MLP:
    float* gates = Kb;                       // [SR][2*192]   (gi | gh)
    float* attnT = Kb + SR * 2 * XG;         // [IMG][n][SP]  attention, slot index fastest (SP = S rounded up to 4)

    // ---- stage to_k weights, load tokens ------------------------------------------------------------------------
    for (int i = tid; i < a.L * TOK_FLOATS / 4; i += FT)
        cp_async16(Wsm + i * 4, a.packed + pk.tok_wt(0) + i * 4);
    float* kin = (a.L & 1) ? Kb : Ka;        // after L ping-pong layers the keys end up in Ka
    float* kout = (a.L & 1) ? Ka : Kb;
    for (int idx = tid; idx < R * (XD / 4); idx += FT) {
        const int r = idx / (XD / 4), e4 = idx - r * (XD / 4);
        const int img = r / n, j = r - img * n;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (img < nimg) {
            const size_t off = ((size_t)(b0 + img) * n + j) * XD + e4 * 4;
            if (a.xpart) {
                for (int s = 0; s < a.nsplit; ++s) {   // fixed order
                    const float4 p = __ldg(reinterpret_cast<const float4*>(a.xpart + (size_t)s * a.split_stride + off));
                    xv.x += p.x; xv.y += p.y; xv.z += p.z; xv.w += p.w;
                }
                const float4 bv = __ldg(reinterpret_cast<const float4*>(a.conv_bias + e4 * 4));
                xv.x = fmaxf(xv.x + bv.x, 0.f); xv.y = fmaxf(xv.y + bv.y, 0.f);
                xv.z = fmaxf(xv.z + bv.z, 0.f); xv.w = fmaxf(xv.w + bv.w, 0.f);
                if (a.x_out) *reinterpret_cast<float4*>(a.x_out + off) = xv;
            } else {
                xv = __ldg(reinterpret_cast<const float4*>(a.x + off));
            }
        }
        const float4 pv = __ldg(reinterpret_cast<const float4*>(a.pe + j * XD + e4 * 4));
        *reinterpret_cast<float4*>(Xs + r * LDX + e4 * 4) = xv;
        *reinterpret_cast<float4*>(kin + r * LDX + e4 * 4) = make_float4(xv.x + pv.x, xv.y + pv.y, xv.z + pv.z, xv.w + pv.w);
    }
    for (int idx = tid; idx < SR * XD; idx += FT) slots[idx] = __ldg(a.packed + pk.slots() + (idx % (S * XD)));
    cp_async_wait_all();
    __syncthreads();

    // ---- to_k MLP: thread = output feature o (weights of its column in registers), 4 token rows at a time ---------
    {
        const int o = tid % XD, rg = tid / XD;   // 8 row groups
        for (int l = 0; l < a.L; ++l) {
            const float* WT = Wsm + l * TOK_FLOATS;
            float wreg[XD];
#pragma unroll
            for (int e = 0; e < XD; ++e) wreg[e] = WT[e * XD + o];
            const float bias = WT[XD * XD + o];
            const bool relu = l + 1 < a.L;
            for (int j0 = rg * 4; j0 < R; j0 += (FT / XD) * 4) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                const float* r0 = kin + min(j0 + 0, R - 1) * LDX;
                const float* r1 = kin + min(j0 + 1, R - 1) * LDX;
                const float* r2 = kin + min(j0 + 2, R - 1) * LDX;
                const float* r3 = kin + min(j0 + 3, R - 1) * LDX;
#pragma unroll
                for (int e4 = 0; e4 < XD / 4; ++e4) {
                    const float4 v0 = *reinterpret_cast<const float4*>(r0 + e4 * 4);
                    const float4 v1 = *reinterpret_cast<const float4*>(r1 + e4 * 4);
                    const float4 v2 = *reinterpret_cast<const float4*>(r2 + e4 * 4);
                    const float4 v3 = *reinterpret_cast<const float4*>(r3 + e4 * 4);
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
                for (int r = 0; r < 4; ++r)
                    if (j0 + r < R) {
                        const float v = acc[r] + bias;
                        kout[(j0 + r) * LDX + o] = relu ? fmaxf(v, 0.f) : v;
                    }
            }
            __syncthreads();
            float* t = kin; kin = kout; kout = t;
        }
    }
    const float* K = Ka;
    // the to_k weights are dead: bring in the GRU block while the first attention pass runs
    if (a.iters > 1)
        for (int i = tid; i < W_FLOATS / 4; i += FT) cp_async16(Wsm + i * 4, a.packed + pk.gru_wih_t() + i * 4);

    for (int it = 0; it < a.iters; ++it) {
        const bool last = it == a.iters - 1;
        // dots[img][i][j] = scale * <slot, key>, kept in attnT[img][j][i]
        for (int idx = tid; idx < SR * n; idx += FT) {
            const int sr = idx / n, j = idx - sr * n;
            const int img = sr / S, i = sr - img * S;
            const float4* sp = reinterpret_cast<const float4*>(slots + sr * XD);
            const float4* kp = reinterpret_cast<const float4*>(K + (img * n + j) * LDX);
            float acc = 0.f;
#pragma unroll
            for (int e4 = 0; e4 < XD / 4; ++e4) {
                const float4 s4 = sp[e4], k4 = kp[e4];
                acc = fmaf(s4.x, k4.x, acc); acc = fmaf(s4.y, k4.y, acc);
                acc = fmaf(s4.z, k4.z, acc); acc = fmaf(s4.w, k4.w, acc);
            }
            attnT[(img * n + j) * SP + i] = acc * 0.125f;
        }
        __syncthreads();
        for (int sr = warp; sr < SR; sr += FW) {          // row sums r_bi, lane-strided then a fixed shuffle tree
            const int img = sr / S, i = sr - img * S;
            float s = 0.f;
            for (int j = lane; j < n; j += 32) s += attnT[(img * n + j) * SP + i];
            s = warp_sum(s);
            if (lane == 0) rsum[sr] = s;
        }
        __syncthreads();
        if (warp < IMG) {                                   // per-image totals t_b
            float s = 0.f;
            for (int i = lane; i < S; i += 32) s += rsum[warp * S + i];
            s = warp_sum(s);
            if (lane == 0) misc[warp] = s;
        }
        __syncthreads();
        for (int idx = tid; idx < SR * n; idx += FT) {      // attn = sigmoid(D / r * t)
            const int sr = idx / n, j = idx - sr * n;
            const int img = sr / S, i = sr - img * S;
            float* p = attnT + (img * n + j) * SP + i;
            const float at = sigm(*p / rsum[sr] * misc[img]);
            *p = at;
            if (last && img < nimg && a.attn) a.attn[((size_t)(b0 + img) * S + i) * n + j] = at;
        }
        __syncthreads();
        // updates[sr][e] = sum_j attn * X / d : thread = (image, e), all S slots of the image in registers (4 at a time)
        for (int w = tid; w < IMG * XD * ((S + 3) / 4); w += FT) {
            const int e = w % XD;
            const int rest = w / XD;
            const int img = rest % IMG, i0 = (rest / IMG) * 4;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const float* xp = Xs + img * n * LDX + e;
            const float* ap = attnT + img * n * SP + i0;
            for (int j = 0; j < n; ++j) {
                const float xv = xp[j * LDX];
                const float4 a4 = *reinterpret_cast<const float4*>(ap + j * SP);
                acc[0] = fmaf(a4.x, xv, acc[0]); acc[1] = fmaf(a4.y, xv, acc[1]);
                acc[2] = fmaf(a4.z, xv, acc[2]); acc[3] = fmaf(a4.w, xv, acc[3]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (i0 + r < S) upd[(img * S + i0 + r) * XD + e] = acc[r] * (1.0f / XD);
        }
        __syncthreads();
        if (last) {
            for (int sr = warp; sr < SR; sr += FW) {
                float s = upd[sr * XD + lane] + upd[sr * XD + 32 + lane];
                s = warp_sum(s);
                if (lane == 0) usum[sr] = s;
            }
        } else {
            if (it == 0) {
                cp_async_wait_all();
                __syncthreads();
            }
            // gate pre-activations: thread = (which in {ih,hh}, gate column g); 8 slot rows at a time
            if (tid < 2 * XG) {
                const int which = tid / XG, g = tid - which * XG;
                const float* WT = Wsm + which * XD * XG;                 // [e][192]
                const float bias = Wsm[2 * XD * XG + which * XG + g];
                const float* src = which ? slots : upd;
                for (int r0 = 0; r0 < SR; r0 += 8) {
                    float acc[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
#pragma unroll 4
                    for (int e4 = 0; e4 < XD / 4; ++e4) {
                        const float w0 = WT[(e4 * 4 + 0) * XG + g], w1 = WT[(e4 * 4 + 1) * XG + g];
                        const float w2 = WT[(e4 * 4 + 2) * XG + g], w3 = WT[(e4 * 4 + 3) * XG + g];
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const float4 v = *reinterpret_cast<const float4*>(src + min(r0 + r, SR - 1) * XD + e4 * 4);
                            acc[r] = fmaf(v.x, w0, acc[r]); acc[r] = fmaf(v.y, w1, acc[r]);
                            acc[r] = fmaf(v.z, w2, acc[r]); acc[r] = fmaf(v.w, w3, acc[r]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 8; ++r)
                        if (r0 + r < SR) gates[(r0 + r) * 2 * XG + which * XG + g] = acc[r] + bias;
                }
            }
            __syncthreads();
            for (int idx = tid; idx < SR * XD; idx += FT) {   // GRU cell, gate order [r|z|n]
                const int sr = idx / XD, e = idx - sr * XD;
                const float* gi = gates + sr * 2 * XG;
                const float* gh = gi + XG;
                const float rg_ = sigm(gi[e] + gh[e]);
                const float zg = sigm(gi[XD + e] + gh[XD + e]);
                const float ng = tanhf(gi[2 * XD + e] + rg_ * gh[2 * XD + e]);
                const float hp = slots[idx];
                slots[idx] = (hp - ng) * zg + ng;              // ATen's form of (1-z)*n + z*h
            }
        }
        __syncthreads();
    }

    for (int idx = tid; idx < nimg * a.C; idx += FT) {
        const int img = idx / a.C, c = idx - img * a.C;
        float s = 0.f;
        for (int m = 0; m < a.spc; ++m) s += usum[img * S + c * a.spc + m];
        a.logits[(size_t)(b0 + img) * a.C + c] = (float)a.loss_status * s;
    }
    if (a.attn_sum) {
        // per-image sum of the final attention (area loss), fixed order, from the stored map
        __syncthreads();
        for (int img = warp; img < nimg; img += FW) {
            float s = 0.f;
            for (int k = lane; k < n * S; k += 32) {
                const int j = k / S, i = k - j * S;
                s += attnT[(img * n + j) * SP + i];
            }
            s = warp_sum(s);
            if (lane == 0) a.attn_sum[b0 + img] = s;
        }
    }
}

size_t fast_smem_bytes(int img, int n, int S) {
    const int R = img * n, SR = img * S;
    size_t fl = 2 * (size_t)R * LDX + kb_floats(img, n, S) + W_FLOATS + 2 * (size_t)SR * XD + 2 * SR + 64;
    return fl * sizeof(float);
}

}  // namespace

bool xslot_fast_supported(const scouter_xslot_desc_t* d, int n) {
    static bool off = getenv("SCOUTER_NO_FAST_XSLOT") != nullptr;
    const int S = d->num_classes * d->slots_per_class;
    if (off || S > 32 || d->to_k_layers * TOK_FLOATS > W_FLOATS) return false;
    return fast_smem_bytes(1, n, S) <= 220 * 1024;
}

int xslot_fast_launch(const scouter_xslot_desc_t* d, const void* packed, const XSlotFastIO& io, cudaStream_t s) {
    const int S = d->num_classes * d->slots_per_class;
    FastArgs a;
    a.packed = (const float*)packed;
    a.x = io.x; a.xpart = io.xpart; a.conv_bias = io.conv_bias; a.split_stride = io.split_stride; a.nsplit = io.nsplit;
    a.pe = io.pe; a.x_out = io.x_out; a.logits = io.logits; a.attn = io.attn; a.attn_sum = io.attn_sum;
    a.B = io.batch; a.n = io.n; a.S = S; a.C = d->num_classes; a.spc = d->slots_per_class; a.L = d->to_k_layers;
    a.iters = d->iters; a.loss_status = d->loss_status;
    const bool two = io.batch > 1 && fast_smem_bytes(2, a.n, S) <= 220 * 1024;
    if (two) {
        const size_t smem = fast_smem_bytes(2, a.n, S);
        SC_CUDA(cudaFuncSetAttribute(xslot_fast_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xslot_fast_kernel<2><<<cdiv(io.batch, 2), FT, smem, s>>>(a);
    } else {
        const size_t smem = fast_smem_bytes(1, a.n, S);
        SC_CUDA(cudaFuncSetAttribute(xslot_fast_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xslot_fast_kernel<1><<<io.batch, FT, smem, s>>>(a);
    }
    SC_LAUNCH_CHECK();
    return 0;
}

}  // namespace scouter
