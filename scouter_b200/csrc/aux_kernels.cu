// Memory-bound helpers of the backbone: pools, split-attention reductions and re-weighting, global average pool.
// All are NHWC, float4 per thread, indexed through a 3-D grid (x = column*channel-quad, y = row, z = image) so that no
// 64-bit division sits on the address path, and written to keep several independent 16-byte loads in flight per
// thread (HBM needs ~90 KB outstanding per SM to reach peak).
//
// Reference ops: nn.MaxPool2d(3,2,1) resnet.py:420; AvgPool2d(2,2,ceil_mode,count_include_pad=False) resnet.py:300;
// AvgPool2d(3,2,1) resnest.py:101; the radix-sum / GAP / r-softmax / re-weighting of split_attn.py:62-79.
#include <algorithm>

#include "common.cuh"

namespace scouter {
namespace {

// Each thread owns ROWS consecutive output rows (same column / channel quad): the loads of all rows are issued before
// any is consumed, which is what keeps enough bytes in flight.
#define SC_PIXEL_INDEX(ROWS)                                        \
    const int cq = C >> 2;                                          \
    const int i_ = blockIdx.x * blockDim.x + threadIdx.x;           \
    if (i_ >= Wo * cq) return;                                      \
    const int wo = i_ / cq, q = i_ - wo * cq;                       \
    const int ho0 = blockIdx.y * (ROWS), b = blockIdx.z

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void add4(float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
__device__ __forceinline__ float4 round4(float4 v) { return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w)); }

template <int ROWS>
__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int C,
                                                      int Ho, int Wo, int k, int stride, int pad) {
    SC_PIXEL_INDEX(ROWS);
    const float* base = in + (size_t)b * H * W * C + q * 4;
    float4 m[ROWS];
    if (k == 3 && stride == 2 && pad == 1) {
        // the stem pool (resnet.py:420): all nine loads of a row are issued before the first max (static window, indices clamped
        // into the map -- a clamped index repeats an element of the same window, which a maximum does not notice)
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            const int ho = ho0 + j;
            if (ho >= Ho) continue;
            float4 v[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int hi = min(max(ho * 2 - 1 + t / 3, 0), H - 1), wi = min(max(wo * 2 - 1 + t % 3, 0), W - 1);
                v[t] = ld4(base + ((size_t)hi * W + wi) * C);
            }
            float4 a = v[0];
#pragma unroll
            for (int t = 1; t < 9; ++t) { a.x = fmaxf(a.x, v[t].x); a.y = fmaxf(a.y, v[t].y); a.z = fmaxf(a.z, v[t].z); a.w = fmaxf(a.w, v[t].w); }
            *reinterpret_cast<float4*>(out + (((size_t)b * Ho + ho) * Wo + wo) * C + q * 4) = a;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) m[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int r = 0; r < k; ++r)
        for (int s = 0; s < k; ++s) {
            const int wi = wo * stride - pad + s;
            if (wi < 0 || wi >= W) continue;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const int hi = (ho0 + j) * stride - pad + r;
                if (hi < 0 || hi >= H) continue;
                const float4 v = ld4(base + ((size_t)hi * W + wi) * C);
                m[j].x = fmaxf(m[j].x, v.x); m[j].y = fmaxf(m[j].y, v.y); m[j].z = fmaxf(m[j].z, v.z); m[j].w = fmaxf(m[j].w, v.w);
            }
        }
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
        if (ho0 + j < Ho) *reinterpret_cast<float4*>(out + (((size_t)b * Ho + ho0 + j) * Wo + wo) * C + q * 4) = m[j];
}

// PyTorch avg_pool2d semantics: the window is first clipped to the padded extent (that size is the divisor when
// count_include_pad), then to the real extent (that size is the divisor otherwise).
template <int ROWS>
__global__ void __launch_bounds__(256) avgpool_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int C,
                                                      int Ho, int Wo, int k, int stride, int pad, int count_include_pad,
                                                      int round_out) {
    SC_PIXEL_INDEX(ROWS);
    const float* base = in + (size_t)b * H * W * C + q * 4;
    int ws = wo * stride - pad;
    const int we_p = min(ws + k, W + pad);
    const int wspan_p = we_p - ws;
    ws = max(ws, 0);
    const int we = min(we_p, W);
    float4 a[ROWS];
    float div[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        int hs = (ho0 + j) * stride - pad;
        int he = min(hs + k, H + pad);
        const int pool = (he - hs) * wspan_p;
        hs = max(hs, 0);
        he = min(he, H);
        div[j] = (float)(count_include_pad ? pool : (he - hs) * (we - ws));
        if (ho0 + j < Ho)
            for (int hi = hs; hi < he; ++hi)
                for (int wi = ws; wi < we; ++wi) add4(a[j], ld4(base + ((size_t)hi * W + wi) * C));
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        if (ho0 + j >= Ho) continue;
        float4 v = a[j];
        v.x /= div[j]; v.y /= div[j]; v.z /= div[j]; v.w /= div[j];
        if (round_out) v = round4(v);
        *reinterpret_cast<float4*>(out + (((size_t)b * Ho + ho0 + j) * Wo + wo) * C + q * 4) = v;
    }
}

// The shortcut pool AvgPool2d(2, 2, ceil_mode=True, count_include_pad=False) (resnet.py:300) with every load of the
// thread's ROWS windows issued before the first add: 4*ROWS independent 16-byte loads in flight per thread.  Same
// summation order and divisor as the generic kernel (bit-identical results).
template <int ROWS>
__global__ void __launch_bounds__(256) avgpool2x2_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int C,
                                                         int Ho, int Wo, int round_out) {
    SC_PIXEL_INDEX(ROWS);
    const float* base = in + (size_t)b * H * W * C + q * 4;
    const int wi0 = 2 * wo;
    const bool w1 = wi0 + 1 < W;
    float4 v[ROWS][4];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const int hi0 = 2 * (ho0 + j);
        const bool r0 = ho0 + j < Ho, r1 = r0 && hi0 + 1 < H;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        v[j][0] = r0 ? ld4(base + ((size_t)hi0 * W + wi0) * C) : z;
        v[j][1] = (r0 && w1) ? ld4(base + ((size_t)hi0 * W + wi0 + 1) * C) : z;
        v[j][2] = r1 ? ld4(base + ((size_t)(hi0 + 1) * W + wi0) * C) : z;
        v[j][3] = (r1 && w1) ? ld4(base + ((size_t)(hi0 + 1) * W + wi0 + 1) * C) : z;
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        if (ho0 + j >= Ho) continue;
        const int hi0 = 2 * (ho0 + j);
        const float div = (float)((hi0 + 1 < H ? 2 : 1) * (w1 ? 2 : 1));
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        add4(a, v[j][0]);
        if (w1) add4(a, v[j][1]);
        if (hi0 + 1 < H) { add4(a, v[j][2]); if (w1) add4(a, v[j][3]); }
        a.x /= div; a.y /= div; a.z /= div; a.w /= div;
        if (round_out) a = round4(a);
        *reinterpret_cast<float4*>(out + (((size_t)b * Ho + ho0 + j) * Wo + wo) * C + q * 4) = a;
    }
}

// ---- split attention --------------------------------------------------------------------------------------------
// Stage 1: per (image, pixel slice) partial sums of all 2C channels: thread = one float4 channel group x one pixel
// lane, four independent accumulators.  Stage 2 adds the slices in a fixed order and the two radix halves.
__global__ void __launch_bounds__(256) splat_gap_partial_kernel(const float* __restrict__ in, float* __restrict__ part, int HW,
                                                                int C2 /* = 2C */, int nsplit) {
    __shared__ float4 red[256];
    const int G = C2 >> 2;                    // float4 groups per pixel (<= 256)
    const int P = 256 / G;                    // pixel lanes
    const int g = threadIdx.x % G, pl = threadIdx.x / G;
    const int b = blockIdx.y, sp = blockIdx.x;
    const int per = (HW + nsplit - 1) / nsplit;
    const int hw0 = sp * per, hw1 = min(HW, hw0 + per);
    const float* base = in + (size_t)b * HW * C2 + g * 4;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    int hw = hw0 + pl;
    if (pl < P) {
        for (; hw + 3 * P < hw1; hw += 4 * P) {
            const float4 v0 = ld4(base + (size_t)hw * C2), v1 = ld4(base + (size_t)(hw + P) * C2);
            const float4 v2 = ld4(base + (size_t)(hw + 2 * P) * C2), v3 = ld4(base + (size_t)(hw + 3 * P) * C2);
            add4(a0, v0); add4(a1, v1); add4(a2, v2); add4(a3, v3);
        }
        for (; hw < hw1; hw += P) add4(a0, ld4(base + (size_t)hw * C2));
    }
    add4(a0, a1); add4(a2, a3); add4(a0, a2);
    red[threadIdx.x] = a0;
    __syncthreads();
    if (pl == 0) {
        float4 t = red[g];
        for (int i = 1; i < P; ++i) add4(t, red[i * G + g]);
        *reinterpret_cast<float4*>(part + ((size_t)b * nsplit + sp) * C2 + g * 4) = t;
    }
}

__global__ void splat_gap_finish_kernel(const float* __restrict__ part, float* __restrict__ gap, int B, int C, int nsplit, float inv_hw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i - b * C;
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) {
        const float* p = part + ((size_t)b * nsplit + sp) * 2 * C;
        s += p[c] + p[C + c];                 // radix sum (split_attn.py:64-65)
    }
    gap[i] = s * inv_hw;
}

// Plain global average pool (B,HW,C) -> (B,C); one CTA = 32 channels x 8 pixel lanes.
__global__ void __launch_bounds__(256) gap_kernel(const float* __restrict__ in, float* __restrict__ out, int HW, int C) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x % 32, hy = threadIdx.x / 32;
    const int b = blockIdx.y, c = blockIdx.x * 32 + cx;
    const float* base = in + (size_t)b * HW * C;
    float s = 0.f;
    if (c < C)
        for (int hw = hy; hw < HW; hw += 8) s += __ldg(base + (size_t)hw * C + c);
    red[hy][cx] = s;
    __syncthreads();
    if (hy == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][cx];
        out[(size_t)b * C + c] = t / (float)HW;
    }
}

// out[b,ho,wo,c] = pool3x3s2p1?( x[b,h,w,c]*a0[b,c] + x[b,h,w,C+c]*a1[b,c] ), (a0,a1) = softmax over the radix pair of
// the fc2 output `logit` (B, 2C) (split_attn.py:14-28,74-79).  The pool divisor follows avg_pool2d with
// count_include_pad=True (resnest.py:101): 9 wherever the padded window is full.
template <int ROWS, bool AVD>
__global__ void __launch_bounds__(256) splat_apply_kernel(const float* __restrict__ in, const float* __restrict__ logit,
                                                          float* __restrict__ out, int H, int W, int C, int Ho, int Wo, int round_out) {
    constexpr bool avd = AVD;   // compile-time: the 18-register-quad window of the pooled form must not cost the plain form occupancy
    SC_PIXEL_INDEX(ROWS);
    const float4 l0 = ld4(logit + (size_t)b * 2 * C + q * 4), l1 = ld4(logit + (size_t)b * 2 * C + C + q * 4);
    const float* base = in + (size_t)b * H * W * 2 * C + q * 4;
    float4 x0[ROWS], x1[ROWS];
    if (!avd) {   // issue every row's loads before the softmax arithmetic
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            const int ho = min(ho0 + j, Ho - 1);
            const float* p = base + ((size_t)ho * W + wo) * 2 * C;
            x0[j] = ld4(p);
            x1[j] = ld4(p + C);
        }
    }
    float4 a0, a1;
    {
        const float y0[4] = {l0.x, l0.y, l0.z, l0.w}, y1[4] = {l1.x, l1.y, l1.z, l1.w};
        float r0[4], r1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float mx = fmaxf(y0[j], y1[j]);
            const float e0 = expf(y0[j] - mx), e1 = expf(y1[j] - mx);
            const float inv = 1.f / (e0 + e1);
            r0[j] = e0 * inv;
            r1[j] = e1 * inv;
        }
        a0 = make_float4(r0[0], r0[1], r0[2], r0[3]);
        a1 = make_float4(r1[0], r1[1], r1[2], r1[3]);
    }
    // the reference multiplies, then sums over the radix axis: x0*a0 + x1*a1 (two roundings + add)
    auto mix = [&](const float4& u0, const float4& u1) {
        return make_float4(u0.x * a0.x + u1.x * a1.x, u0.y * a0.y + u1.y * a1.y, u0.z * a0.z + u1.z * a1.z,
                           u0.w * a0.w + u1.w * a1.w);
    };
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const int ho = ho0 + j;
        if (ho >= Ho) continue;
        float4 r;
        if (!avd) {
            r = mix(x0[j], x1[j]);
        } else {
            int hs = ho * 2 - 1, ws = wo * 2 - 1;
            int he = min(hs + 3, H + 1), we = min(ws + 3, W + 1);
            const float div = (float)((he - hs) * (we - ws));
            hs = max(hs, 0); ws = max(ws, 0);
            he = min(he, H); we = min(we, W);
            r = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int hi = hs; hi < he; ++hi)
                for (int wi = ws; wi < we; ++wi) {
                    const float* p = base + ((size_t)hi * W + wi) * 2 * C;
                    add4(r, mix(ld4(p), ld4(p + C)));
                }
            r.x /= div; r.y /= div; r.z /= div; r.w /= div;
        }
        if (round_out) r = round4(r);
        *reinterpret_cast<float4*>(out + (((size_t)b * Ho + ho) * Wo + wo) * C + q * 4) = r;
    }
}

// The pooled form (first block of layers 2-4: AvgPool2d(3, 2, padding=1) after the re-weighting, resnest.py:101,131) as a
// rolling window down the rows: thread = (output column, channel quad) walks RC output rows; per step it loads the two new
// input rows of the window (3 columns x 2 radix halves = 12 quads issued together) and keeps the column-summed previous
// row, so every input element is read once per column window instead of once per overlapping 3x3 window (the window form
// above ran at 3.4-3.8 TB/s).  Sums run column-first, then rows (the reference's order inside a window is row-major; the
// difference is fp32 rounding of a 9-term sum).
template <int RC>
__global__ void __launch_bounds__(256) splat_apply_pool_kernel(const float* __restrict__ in, const float* __restrict__ logit,
                                                               float* __restrict__ out, int H, int W, int C, int Ho, int Wo, int round_out) {
    const int cq = C >> 2;
    const int i_ = blockIdx.x * blockDim.x + threadIdx.x;
    if (i_ >= Wo * cq) return;
    const int wo = i_ / cq, q = i_ - wo * cq;
    const int b = blockIdx.z;
    const int ho_begin = blockIdx.y * RC, ho_end = min(ho_begin + RC, Ho);
    const float4 l0 = ld4(logit + (size_t)b * 2 * C + q * 4), l1 = ld4(logit + (size_t)b * 2 * C + C + q * 4);
    const float* base = in + (size_t)b * H * W * 2 * C + q * 4;
    float a0[4], a1[4];
    {
        const float y0[4] = {l0.x, l0.y, l0.z, l0.w}, y1[4] = {l1.x, l1.y, l1.z, l1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float mx = fmaxf(y0[j], y1[j]);
            const float e0 = expf(y0[j] - mx), e1 = expf(y1[j] - mx);
            const float inv = 1.f / (e0 + e1);
            a0[j] = e0 * inv;
            a1[j] = e1 * inv;
        }
    }
    // columns of the window: 2wo-1, 2wo, 2wo+1, clipped (a clipped column is loaded from a valid address and weighted 0)
    int wc[3];
    float wk[3];
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3) {
        const int wi = 2 * wo - 1 + s3;
        wk[s3] = (wi >= 0 && wi < W) ? 1.f : 0.f;
        wc[s3] = min(max(wi, 0), W - 1);
    }
    const int ws = 2 * wo - 1, we_p = min(ws + 3, W + 1);
    const float wspan = (float)(we_p - ws);            // count_include_pad=True: padded columns count, columns past W+pad do not
    // column sum of input row hi (zeros when the row is outside the map); x0*a0 + x1*a1 per pixel like the reference
    auto load_row = [&](int hi, float4 (&u)[6]) {
        const int hc = min(max(hi, 0), H - 1);
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) {
            const float* p = base + ((size_t)hc * W + wc[s3]) * 2 * C;
            u[2 * s3] = ld4(p);
            u[2 * s3 + 1] = ld4(p + C);
        }
    };
    auto row_sum = [&](int hi, const float4 (&u)[6]) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hi < 0 || hi >= H) return r;
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) {
            const float4 &x0 = u[2 * s3], &x1 = u[2 * s3 + 1];
            r.x += wk[s3] * (x0.x * a0[0] + x1.x * a1[0]);
            r.y += wk[s3] * (x0.y * a0[1] + x1.y * a1[1]);
            r.z += wk[s3] * (x0.z * a0[2] + x1.z * a1[2]);
            r.w += wk[s3] * (x0.w * a0[3] + x1.w * a1[3]);
        }
        return r;
    };
    float4 prev;
    {
        float4 u[6];
        load_row(2 * ho_begin - 1, u);
        prev = row_sum(2 * ho_begin - 1, u);
    }
    for (int ho = ho_begin; ho < ho_end; ++ho) {
        float4 ua[6], ub[6];
        load_row(2 * ho, ua);
        load_row(2 * ho + 1, ub);
        const float4 ra = row_sum(2 * ho, ua), rb = row_sum(2 * ho + 1, ub);
        const int hs = 2 * ho - 1, he_p = min(hs + 3, H + 1);
        const float div = (float)(he_p - hs) * wspan;
        float4 r = make_float4((prev.x + ra.x + rb.x) / div, (prev.y + ra.y + rb.y) / div, (prev.z + ra.z + rb.z) / div,
                               (prev.w + ra.w + rb.w) / div);
        prev = rb;
        if (round_out) r = round4(r);
        *reinterpret_cast<float4*>(out + (((size_t)b * Ho + ho) * Wo + wo) * C + q * 4) = r;
    }
}

// f2: uint8 HWC -> normalised fp32 NCHW, fp64 arithmetic with a single rounding (reference: float64 ToTensor/Normalize,
// then .to(float32)).  One thread = one pixel (all channels), coalesced planar stores.
struct NormConst { double mean[4], inv255, std[4]; };
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ img, float* __restrict__ out,
                                                            long long npix_per_img, int C, NormConst k) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= npix_per_img) return;
    const uint8_t* src = img + ((long long)b * npix_per_img + p) * C;
    for (int c = 0; c < C; ++c) {
        const double v = ((double)src[c] / 255.0 - k.mean[c]) / k.std[c];
        out[((long long)b * C + c) * npix_per_img + p] = (float)v;
    }
}

int pixel_grid(int B, int Ho, int Wo, int C, int rows, dim3& grid) {
    SC_CHECK_ARG(C % 4 == 0, SCOUTER_E_UNSUPPORTED, "C = %d not a multiple of 4", C);
    SC_CHECK_ARG(B <= 65535 && Ho <= 65535, SCOUTER_E_UNSUPPORTED, "batch %d / height %d exceed the grid limits", B, Ho);
    grid = dim3(cdiv(Wo * (C / 4), 256), cdiv(Ho, rows), B);
    return 0;
}

}  // namespace

int launch_maxpool(const float* in, float* out, int B, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad,
                   cudaStream_t s) {
    dim3 grid;
    if (int e = pixel_grid(B, Ho, Wo, C, 2, grid)) return e;
    maxpool_kernel<2><<<grid, 256, 0, s>>>(in, out, H, W, C, Ho, Wo, k, stride, pad);
    SC_LAUNCH_CHECK();
    return 0;
}

int launch_avgpool(const float* in, float* out, int B, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad,
                   int count_include_pad, int round_out, cudaStream_t s) {
    dim3 grid;
    if (int e = pixel_grid(B, Ho, Wo, C, 4, grid)) return e;
    if (k == 2 && stride == 2 && pad == 0)
        avgpool2x2_kernel<4><<<grid, 256, 0, s>>>(in, out, H, W, C, Ho, Wo, round_out);
    else
        avgpool_kernel<4><<<grid, 256, 0, s>>>(in, out, H, W, C, Ho, Wo, k, stride, pad, count_include_pad, round_out);
    SC_LAUNCH_CHECK();
    return 0;
}

int splat_gap_splits(int B, int HW) {
    // enough CTAs for ~8 per SM, but never slices shorter than 32 pixels
    int n = std::max(1, std::min(cdiv(148 * 8, B), HW / 32));
    return std::min(n, 64);
}

// Many slots (the conv epilogue writes one per tile and epilogue warp: 112 on the 56x56 maps): CTA = image, the slots are spread
// over thread lanes and merged through shared memory in lane order (fixed order) -- the thread-per-channel kernel above walked
// them serially with 64 CTAs in flight (35 us per call on layer 1).
__global__ void __launch_bounds__(512) splat_gap_finish_wide_kernel(const float* __restrict__ part, float* __restrict__ gap, int C, int nslots,
                                                                    float inv_hw) {
    __shared__ float red[512];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int cw = C < 512 ? C : 512;          // channels handled per pass (C is a power of two here: 64..512)
    const int lanes = 512 / cw;
    const int cl = tid % cw, sl = tid / cw;
    for (int c0 = 0; c0 < C; c0 += cw) {
        const int c = c0 + cl;
        float s = 0.f;
        for (int sp = sl; sp < nslots; sp += lanes) {
            const float* p = part + ((size_t)b * nslots + sp) * 2 * C;
            s += __ldg(p + c) + __ldg(p + C + c);       // radix sum (split_attn.py:64-65)
        }
        red[tid] = s;
        __syncthreads();
        if (sl == 0) {
            float t = red[cl];
            for (int l = 1; l < lanes; ++l) t += red[l * cw + cl];
            gap[(size_t)b * C + c] = t * inv_hw;
        }
        __syncthreads();
    }
}

int launch_splat_gap_finish(const float* part, float* gap, int B, int HW, int C, int nslots, cudaStream_t s) {
    if (nslots >= 16 && (C & (C - 1)) == 0 && C >= 32 && 512 % (C < 512 ? C : 512) == 0) {
        splat_gap_finish_wide_kernel<<<B, 512, 0, s>>>(part, gap, C, nslots, 1.0f / (float)HW);
        SC_LAUNCH_CHECK();
        return 0;
    }
    splat_gap_finish_kernel<<<cdiv(B * C, 256), 256, 0, s>>>(part, gap, B, C, nslots, 1.0f / (float)HW);
    SC_LAUNCH_CHECK();
    return 0;
}

int launch_splat_gap(const float* in, float* part, float* gap, int B, int HW, int C, cudaStream_t s) {
    SC_CHECK_ARG(C % 2 == 0 && (2 * C / 4) <= 256 && 256 % (2 * C / 4) == 0, SCOUTER_E_UNSUPPORTED,
                 "splat gap: C = %d (2C/4 must divide 256)", C);
    SC_CHECK_ARG(B <= 65535, SCOUTER_E_UNSUPPORTED, "splat gap: batch %d", B);
    const int nsplit = splat_gap_splits(B, HW);
    splat_gap_partial_kernel<<<dim3(nsplit, B), 256, 0, s>>>(in, part, HW, 2 * C, nsplit);
    SC_LAUNCH_CHECK();
    splat_gap_finish_kernel<<<cdiv(B * C, 256), 256, 0, s>>>(part, gap, B, C, nsplit, 1.0f / (float)HW);
    SC_LAUNCH_CHECK();
    return 0;
}

int launch_gap(const float* in, float* out, int B, int HW, int C, cudaStream_t s) {
    dim3 grid(cdiv(C, 32), B);
    gap_kernel<<<grid, 256, 0, s>>>(in, out, HW, C);
    SC_LAUNCH_CHECK();
    return 0;
}

int launch_splat_apply(const float* in, const float* logit, float* out, int B, int H, int W, int C, int Ho, int Wo, int avd,
                       int round_out, cudaStream_t s) {
    dim3 grid;
    if (avd) {
        static bool window_form = getenv("SCOUTER_SPLAT_POOL_WINDOW") != nullptr;   // the round-1 kernel, for A/B
        if (window_form) {
            if (int e = pixel_grid(B, Ho, Wo, C, 2, grid)) return e;
            splat_apply_kernel<2, true><<<grid, 256, 0, s>>>(in, logit, out, H, W, C, Ho, Wo, round_out);
        } else {
            if (int e = pixel_grid(B, Ho, Wo, C, 7, grid)) return e;
            splat_apply_pool_kernel<7><<<grid, 256, 0, s>>>(in, logit, out, H, W, C, Ho, Wo, round_out);
        }
    } else {
        if (int e = pixel_grid(B, Ho, Wo, C, 4, grid)) return e;
        splat_apply_kernel<4, false><<<grid, 256, 0, s>>>(in, logit, out, H, W, C, Ho, Wo, round_out);
    }
    SC_LAUNCH_CHECK();
    return 0;
}

int launch_preprocess_u8(const uint8_t* img, int B, int H, int W, int C, const double* mean, const double* stdv, float* out,
                         cudaStream_t s) {
    SC_CHECK_ARG(C >= 1 && C <= 4 && B <= 65535, SCOUTER_E_UNSUPPORTED, "preprocess_u8: channels=%d batch=%d", C, B);
    NormConst k;
    for (int c = 0; c < 4; ++c) { k.mean[c] = c < C ? mean[c] : 0.0; k.std[c] = c < C ? stdv[c] : 1.0; }
    k.inv255 = 1.0 / 255.0;
    const long long npix = (long long)H * W;
    preprocess_u8_kernel<<<dim3((unsigned)((npix + 255) / 256), B), 256, 0, s>>>(img, out, npix, C, k);
    SC_LAUNCH_CHECK();
    return 0;
}

}  // namespace scouter
