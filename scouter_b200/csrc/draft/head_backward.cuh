// DRAFT -- row f1, NOT part of libscouter_b200.so and never run on a GPU yet.
//
// Backward of the xSlot head (slot_model.py:108-116 + slot_attention.py:44-96 under loss.backward(), engine.py:31):
// a first, correctness-first transcription of oracle/head_backward.py.  One CTA per image; every intermediate lives in
// a per-image scratch slab in global memory (L1/L2 resident, a few hundred KB), every phase is a thread-strided loop
// over independent outputs followed by a block barrier, parameter gradients are accumulated across images with fp32
// atomics.  Because of that structure the body compiles unchanged as host code (one "thread", barriers are no-ops),
// which is how tests/test_head_backward_draft.py checks it against the oracle without a GPU.  What it is not yet:
// fused with anything, tensor-core based, or bit-reproducible (atomic order).
//
// Gradients of  sum(g_logits * logits) + g_attn_loss * attn_loss  w.r.t. the backbone features (d_feat) and every
// head parameter; `attn_coef` = g_attn_loss * power * m^(power-1) / (B*S*n) with m = sum(attn_last) / (B*S*n) comes
// from the forward's attn_sum (tiny prep kernel in head_backward.cu).
#pragma once
#include <math.h>
#include <stddef.h>

#ifdef __CUDACC__
#define HB_HD __device__ __forceinline__
#define HB_HOST_DEVICE __host__ __device__ inline
#define HB_SYNC() __syncthreads()
#define HB_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define HB_HD static inline
#define HB_HOST_DEVICE static inline
#define HB_SYNC() ((void)0)
#define HB_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

namespace scouter_draft {

constexpr int HD = 64;          // hidden_dim (train.py:60); the library supports only 64 (xslot.cuh XD)
constexpr int HB_MAX_L = 8;

struct HeadBwdArgs {
    int B, n, ch, S, C, spc, L, iters, loss_status;
    const float *feat;                         // (B, n, ch) token-major backbone output
    const float *conv_w, *conv_b, *pe;         // (64, ch), (64), (n, 64)
    const float *to_k_w[HB_MAX_L], *to_k_b[HB_MAX_L];
    const float *w_ih, *w_hh, *b_ih, *b_hh;    // (192, 64) x2, (192) x2, gate order [r | z | n]
    const float *slots0;                       // (S, 64) initial_slots
    const float *g_logits;                     // (B, C)
    const float *attn_coef;                    // scalar on the device
    float *d_feat;                             // (B, n, ch), or NULL when d_pre is requested instead
    float *d_pre;                              // (B, n, 64) gradient at the conv1x1 pre-activation, or NULL.  With it the
                                               // two ch-sized products leave this kernel: d_feat = d_pre W is a 1x1
                                               // conv (64 -> ch) the forward tcgen05 kernel already runs, and
                                               // g_conv_w = d_pre^T feat is a wgrad GEMM over K = B*n rows
    float *g_conv_w, *g_conv_b, *g_to_k_w[HB_MAX_L], *g_to_k_b[HB_MAX_L];
    float *g_w_ih, *g_w_hh, *g_b_ih, *g_b_hh, *g_slots0;     // zero-initialised by the caller, accumulated here
    float *scratch;                            // B * scratch_floats(...)
    size_t scratch_per_image;
};

struct HeadBwdLayout {      // offsets (floats) into one image's scratch slab
    size_t X, A, SL, DOTS, ATT, UPD, GR, GZ, GN, GHN, GI, GH, ROW, DROW, SCAL, DUPD, DS, DSN, DDOT, DX, DK, DA0, DA1, total;
};

HB_HOST_DEVICE HeadBwdLayout head_bwd_layout(int n, int S, int L, int iters) {
    HeadBwdLayout o;
    size_t p = 0;
    const size_t nd = (size_t)n * HD, sd = (size_t)S * HD, sn = (size_t)S * n;
    o.X = p; p += nd;
    o.A = p; p += (size_t)(L + 1) * nd;            // A[0] = x + pe, A[l] = output of Linear l-1 (ReLU'd unless last)
    o.SL = p; p += (size_t)iters * sd;             // s_0 .. s_{iters-1}
    o.DOTS = p; p += (size_t)iters * sn;
    o.ATT = p; p += (size_t)iters * sn;
    o.UPD = p; p += (size_t)iters * sd;
    o.GR = p; p += (size_t)iters * sd;             // GRU caches of step t (r, z, n, gh_n); the last step is never run
    o.GZ = p; p += (size_t)iters * sd;
    o.GN = p; p += (size_t)iters * sd;
    o.GHN = p; p += (size_t)iters * sd;
    o.GI = p; p += 3 * sd;
    o.GH = p; p += 3 * sd;
    o.ROW = p; p += S;
    o.DROW = p; p += S;
    o.SCAL = p; p += 4;                            // [0] tot, [1] d_tot
    o.DUPD = p; p += sd;
    o.DS = p; p += sd;
    o.DSN = p; p += sd;
    o.DDOT = p; p += sn;
    o.DX = p; p += nd;
    o.DK = p; p += nd;
    o.DA0 = p; p += nd;
    o.DA1 = p; p += nd;
    o.total = p;
    return o;
}

HB_HD float hb_sigmoid(float v) { return 1.0f / (1.0f + expf(-v)); }

// thread-strided loop over N independent outputs + block barrier
#define HB_PHASE(N) for (long long idx = tid; idx < (long long)(N); idx += nthreads)

HB_HD void head_backward_image(const HeadBwdArgs& a, int b, int tid, int nthreads) {
    const int n = a.n, S = a.S, ch = a.ch, L = a.L, T = a.iters;
    const HeadBwdLayout o = head_bwd_layout(n, S, L, T);
    float* w = a.scratch + (size_t)b * a.scratch_per_image;
    const float* feat = a.feat + (size_t)b * n * ch;
    const size_t nd = (size_t)n * HD, sd = (size_t)S * HD, sn = (size_t)S * n;
    const float scale = 0.125f;                    // 64^-1/2
    const float inv_d = 1.0f / HD;

    // ================================ forward (recomputed) =======================================
    HB_PHASE(nd) {                                 // x = relu(conv1x1(feat)); A[0] = x + pe
        const int j = (int)(idx / HD), e = (int)(idx % HD);
        const float* f = feat + (size_t)j * ch;
        const float* cw = a.conv_w + (size_t)e * ch;
        float s = a.conv_b[e];
        for (int c = 0; c < ch; ++c) s = fmaf(f[c], cw[c], s);
        s = s > 0.f ? s : 0.f;
        w[o.X + idx] = s;
        w[o.A + idx] = s + a.pe[idx];
    }
    HB_SYNC();
    for (int l = 0; l < L; ++l) {                  // to_k MLP (slot_attention.py:30-37)
        const float* in = w + o.A + (size_t)l * nd;
        float* out = w + o.A + (size_t)(l + 1) * nd;
        HB_PHASE(nd) {
            const int j = (int)(idx / HD), e = (int)(idx % HD);
            float s = a.to_k_b[l][e];
            for (int i = 0; i < HD; ++i) s = fmaf(in[(size_t)j * HD + i], a.to_k_w[l][e * HD + i], s);
            out[idx] = (l + 1 < L && s < 0.f) ? 0.f : s;
        }
        HB_SYNC();
    }
    const float* X = w + o.X;
    const float* K = w + o.A + (size_t)L * nd;
    HB_PHASE(sd) w[o.SL + idx] = a.slots0[idx];
    HB_SYNC();
    for (int t = 0; t < T; ++t) {
        const float* s_t = w + o.SL + (size_t)t * sd;
        float* dots = w + o.DOTS + (size_t)t * sn;
        float* att = w + o.ATT + (size_t)t * sn;
        float* upd = w + o.UPD + (size_t)t * sd;
        HB_PHASE(sn) {
            const int i = (int)(idx / n), j = (int)(idx % n);
            float s = 0.f;
            for (int e = 0; e < HD; ++e) s = fmaf(s_t[i * HD + e], K[(size_t)j * HD + e], s);
            dots[idx] = s * scale;
        }
        HB_SYNC();
        HB_PHASE(S) {
            float s = 0.f;
            for (int j = 0; j < n; ++j) s += dots[idx * n + j];
            w[o.ROW + idx] = s;
        }
        HB_SYNC();
        HB_PHASE(1) {
            float s = 0.f;
            for (int i = 0; i < S; ++i) s += w[o.ROW + i];
            w[o.SCAL] = s;
        }
        HB_SYNC();
        HB_PHASE(sn) {
            const int i = (int)(idx / n);
            att[idx] = hb_sigmoid(dots[idx] / w[o.ROW + i] * w[o.SCAL]);
        }
        HB_SYNC();
        HB_PHASE(sd) {
            const int i = (int)(idx / HD), e = (int)(idx % HD);
            float s = 0.f;
            for (int j = 0; j < n; ++j) s = fmaf(att[i * n + j], X[(size_t)j * HD + e], s);
            upd[idx] = s * inv_d;
        }
        HB_SYNC();
        if (t + 1 < T) {                           // s_{t+1} = GRU(upd_t, s_t); the last step is dead
            HB_PHASE(3 * sd) {
                const int i = (int)(idx / (3 * HD)), g = (int)(idx % (3 * HD));
                float gi = a.b_ih[g], gh = a.b_hh[g];
                for (int e = 0; e < HD; ++e) {
                    gi = fmaf(upd[i * HD + e], a.w_ih[g * HD + e], gi);
                    gh = fmaf(s_t[i * HD + e], a.w_hh[g * HD + e], gh);
                }
                w[o.GI + idx] = gi;
                w[o.GH + idx] = gh;
            }
            HB_SYNC();
            HB_PHASE(sd) {
                const int i = (int)(idx / HD), e = (int)(idx % HD);
                const float* gi = w + o.GI + (size_t)i * 3 * HD;
                const float* gh = w + o.GH + (size_t)i * 3 * HD;
                const float r = hb_sigmoid(gi[e] + gh[e]);
                const float z = hb_sigmoid(gi[HD + e] + gh[HD + e]);
                const float nn = tanhf(gi[2 * HD + e] + r * gh[2 * HD + e]);
                w[o.GR + (size_t)t * sd + idx] = r;
                w[o.GZ + (size_t)t * sd + idx] = z;
                w[o.GN + (size_t)t * sd + idx] = nn;
                w[o.GHN + (size_t)t * sd + idx] = gh[2 * HD + e];
                w[o.SL + (size_t)(t + 1) * sd + idx] = (1.f - z) * nn + z * s_t[idx];
            }
            HB_SYNC();
        }
    }

    // ================================ backward ===================================================
    const float attn_coef = *a.attn_coef;
    HB_PHASE(sd) {                                 // d upd_{T-1} = loss_status * g_logits[class of the slot]
        const int i = (int)(idx / HD);
        w[o.DUPD + idx] = (float)a.loss_status * a.g_logits[(size_t)b * a.C + i / a.spc];
        w[o.DSN + idx] = 0.f;
    }
    HB_PHASE(nd) { w[o.DX + idx] = 0.f; w[o.DK + idx] = 0.f; }
    HB_SYNC();
    for (int t = T - 1; t >= 0; --t) {
        const float* s_t = w + o.SL + (size_t)t * sd;
        const float* dots = w + o.DOTS + (size_t)t * sn;
        const float* att = w + o.ATT + (size_t)t * sn;
        const float* upd = w + o.UPD + (size_t)t * sd;
        if (t + 1 < T) {                           // through s_{t+1} = GRU(upd_t, s_t), given d s_{t+1} in DSN
            const float* R = w + o.GR + (size_t)t * sd;
            const float* Z = w + o.GZ + (size_t)t * sd;
            const float* NN = w + o.GN + (size_t)t * sd;
            const float* GHN = w + o.GHN + (size_t)t * sd;
            HB_PHASE(sd) {
                const int i = (int)(idx / HD), e = (int)(idx % HD);
                const float dsn = w[o.DSN + idx], r = R[idx], z = Z[idx], nn = NN[idx];
                const float da_n = dsn * (1.f - z) * (1.f - nn * nn);
                const float da_r = da_n * GHN[idx] * r * (1.f - r);
                const float da_z = dsn * (s_t[idx] - nn) * z * (1.f - z);
                float* dgi = w + o.GI + (size_t)i * 3 * HD;
                float* dgh = w + o.GH + (size_t)i * 3 * HD;
                dgi[e] = da_r; dgi[HD + e] = da_z; dgi[2 * HD + e] = da_n;
                dgh[e] = da_r; dgh[HD + e] = da_z; dgh[2 * HD + e] = da_n * r;
                w[o.DS + idx] = dsn * z;           // direct path s_t -> s_{t+1}
            }
            HB_SYNC();
            HB_PHASE(sd) {                         // d upd_t = dgi W_ih ; d s_t += dgh W_hh
                const int i = (int)(idx / HD), e = (int)(idx % HD);
                const float* dgi = w + o.GI + (size_t)i * 3 * HD;
                const float* dgh = w + o.GH + (size_t)i * 3 * HD;
                float du = 0.f, ds = 0.f;
                for (int g = 0; g < 3 * HD; ++g) {
                    du = fmaf(dgi[g], a.w_ih[g * HD + e], du);
                    ds = fmaf(dgh[g], a.w_hh[g * HD + e], ds);
                }
                w[o.DUPD + idx] = du;
                w[o.DS + idx] += ds;
            }
            HB_PHASE((size_t)3 * HD * HD) {        // weight gradients of this step (reads GI/GH/upd/s_t only)
                const int g = (int)(idx / HD), e = (int)(idx % HD);
                float gi = 0.f, gh = 0.f;
                for (int i = 0; i < S; ++i) {
                    gi = fmaf(w[o.GI + (size_t)i * 3 * HD + g], upd[i * HD + e], gi);
                    gh = fmaf(w[o.GH + (size_t)i * 3 * HD + g], s_t[i * HD + e], gh);
                }
                HB_ATOMIC_ADD(a.g_w_ih + idx, gi);
                HB_ATOMIC_ADD(a.g_w_hh + idx, gh);
            }
            HB_PHASE(3 * HD) {
                float gi = 0.f, gh = 0.f;
                for (int i = 0; i < S; ++i) { gi += w[o.GI + (size_t)i * 3 * HD + idx]; gh += w[o.GH + (size_t)i * 3 * HD + idx]; }
                HB_ATOMIC_ADD(a.g_b_ih + idx, gi);
                HB_ATOMIC_ADD(a.g_b_hh + idx, gh);
            }
            HB_SYNC();
        } else {
            HB_PHASE(sd) w[o.DS + idx] = 0.f;
            HB_SYNC();
        }
        const float* dupd = w + o.DUPD;
        HB_PHASE(sn) {                             // d attn = d upd x^T / d (+ area-loss term on the last iteration); d u
            const int i = (int)(idx / n), j = (int)(idx % n);
            float s = 0.f;
            for (int e = 0; e < HD; ++e) s = fmaf(dupd[i * HD + e], X[(size_t)j * HD + e], s);
            s *= inv_d;
            if (t == T - 1) s += attn_coef;
            const float at = att[idx];
            w[o.DDOT + idx] = s * at * (1.f - at); // = d u
        }
        HB_PHASE(nd) {                             // d x += attn^T d upd / d
            const int j = (int)(idx / HD), e = (int)(idx % HD);
            float s = 0.f;
            for (int i = 0; i < S; ++i) s = fmaf(att[i * n + j], dupd[i * HD + e], s);
            w[o.DX + idx] += s * inv_d;
        }
        HB_PHASE(S) {                              // row sums of this iteration (ROW was overwritten by later ones)
            float s = 0.f;
            for (int j = 0; j < n; ++j) s += dots[idx * n + j];
            w[o.ROW + idx] = s;
        }
        HB_SYNC();
        HB_PHASE(1) {
            float s = 0.f;
            for (int i = 0; i < S; ++i) s += w[o.ROW + i];
            w[o.SCAL] = s;
        }
        HB_SYNC();
        HB_PHASE(S) {                              // u = dots * tot / row:  d row_i, and the per-row part of d tot
            const float row = w[o.ROW + idx], tot = w[o.SCAL];
            float s = 0.f;
            for (int j = 0; j < n; ++j) s = fmaf(w[o.DDOT + idx * n + j], dots[idx * n + j], s);
            w[o.DROW + idx] = -s * tot / (row * row);
            w[o.GI + idx] = s / row;               // GI is free here: partial of d tot
        }
        HB_SYNC();
        HB_PHASE(1) {
            float s = 0.f;
            for (int i = 0; i < S; ++i) s += w[o.GI + i];
            w[o.SCAL + 1] = s;
        }
        HB_SYNC();
        HB_PHASE(sn) {
            const int i = (int)(idx / n);
            w[o.DDOT + idx] = w[o.DDOT + idx] * w[o.SCAL] / w[o.ROW + i] + w[o.DROW + i] + w[o.SCAL + 1];   // d dots
        }
        HB_SYNC();
        HB_PHASE(sd) {                             // d s_t += d dots k * scale  -> becomes d s_{t+1} of the next (earlier) step
            const int i = (int)(idx / HD), e = (int)(idx % HD);
            float s = 0.f;
            for (int j = 0; j < n; ++j) s = fmaf(w[o.DDOT + i * n + j], K[(size_t)j * HD + e], s);
            w[o.DSN + idx] = w[o.DS + idx] + s * scale;
        }
        HB_PHASE(nd) {                             // d k += d dots^T s_t * scale
            const int j = (int)(idx / HD), e = (int)(idx % HD);
            float s = 0.f;
            for (int i = 0; i < S; ++i) s = fmaf(w[o.DDOT + i * n + j], s_t[i * HD + e], s);
            w[o.DK + idx] += s * scale;
        }
        HB_SYNC();
    }
    HB_PHASE(sd) HB_ATOMIC_ADD(a.g_slots0 + idx, w[o.DSN + idx]);

    // to_k MLP backward: d_a walks from d k to d (x + pe)
    const float* d_in = w + o.DK;
    for (int l = L - 1; l >= 0; --l) {
        const float* in = w + o.A + (size_t)l * nd;               // input of Linear l
        const float* outp = w + o.A + (size_t)(l + 1) * nd;       // its (ReLU'd) output
        float* d_a = w + ((L - 1 - l) % 2 ? o.DA1 : o.DA0);
        HB_PHASE(nd) d_a[idx] = (l + 1 < L && !(outp[idx] > 0.f)) ? 0.f : d_in[idx];
        HB_SYNC();
        HB_PHASE((size_t)HD * HD) {
            const int e = (int)(idx / HD), i = (int)(idx % HD);
            float s = 0.f;
            for (int j = 0; j < n; ++j) s = fmaf(d_a[(size_t)j * HD + e], in[(size_t)j * HD + i], s);
            HB_ATOMIC_ADD(a.g_to_k_w[l] + idx, s);
        }
        HB_PHASE(HD) {
            float s = 0.f;
            for (int j = 0; j < n; ++j) s += d_a[(size_t)j * HD + idx];
            HB_ATOMIC_ADD(a.g_to_k_b[l] + idx, s);
        }
        float* d_prev = w + ((L - 1 - l) % 2 ? o.DA0 : o.DA1);    // the other buffer
        HB_PHASE(nd) {
            const int j = (int)(idx / HD), i = (int)(idx % HD);
            float s = 0.f;
            for (int e = 0; e < HD; ++e) s = fmaf(d_a[(size_t)j * HD + e], a.to_k_w[l][e * HD + i], s);
            d_prev[idx] = s;
        }
        HB_SYNC();
        d_in = d_prev;
    }
    HB_PHASE(nd) {                                 // d pre-activation of conv1x1: (d x + d(x+pe)) * [x > 0]
        const float v = w[o.DX + idx] + d_in[idx];
        w[o.DX + idx] = X[idx] > 0.f ? v : 0.f;
    }
    HB_SYNC();
    const float* dpre = w + o.DX;
    HB_PHASE(HD) {                                 // conv1x1.bias gradient (always here: it is 64 sums)
        float s = 0.f;
        for (int j = 0; j < n; ++j) s += dpre[(size_t)j * HD + idx];
        HB_ATOMIC_ADD(a.g_conv_b + idx, s);
    }
    if (a.d_pre) {                                 // hand the ch-sized products to the GEMM kernels
        HB_PHASE(nd) a.d_pre[(size_t)b * nd + idx] = dpre[idx];
        return;
    }
    HB_PHASE((size_t)n * ch) {                     // d feat = d pre W
        const int j = (int)(idx / ch), c = (int)(idx % ch);
        float s = 0.f;
        for (int e = 0; e < HD; ++e) s = fmaf(dpre[(size_t)j * HD + e], a.conv_w[(size_t)e * ch + c], s);
        a.d_feat[(size_t)b * n * ch + idx] = s;
    }
    HB_PHASE((size_t)HD * ch) {                    // d W = d pre^T feat
        const int e = (int)(idx / ch), c = (int)(idx % ch);
        float s = 0.f;
        for (int j = 0; j < n; ++j) s = fmaf(dpre[(size_t)j * HD + e], feat[(size_t)j * ch + c], s);
        HB_ATOMIC_ADD(a.g_conv_w + idx, s);
    }
}

}  // namespace scouter_draft
