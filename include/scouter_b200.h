/*
 * scouter_b200.h -- C ABI of libscouter_b200.so: the SCOUTER forward hot path on B200 (sm_100a).
 *
 * The reference (wbw520/scouter) has no FFI / operator layer: the boundary its hot path sits behind
 * is the Python nn.Module API (sloter/slot_model.py:18-127, sloter/utils/slot_attention.py:9-96) and
 * every numeric primitive below it is a PyTorch library call.  This header is therefore the set of
 * entry points a binding for that path would need; each one cites the reference code it replaces.
 * The Python mirror of the reference modules (scouter_b200/) calls these through ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - All tensor pointers are DEVICE pointers owned by the caller unless a name ends in _host.
 *     The library never frees or retains caller buffers beyond the lifetime documented per call.
 *   - Every launch is stream-ordered on `stream` (a cudaStream_t passed as void*); no call
 *     synchronises the device except scouter_forward_host.
 *   - Return value: 0 = OK; < 0 = invalid argument / unsupported shape (SCOUTER_E_*); > 0 = a
 *     cudaError_t.  scouter_last_error() returns a thread-local message for the last failure.
 *   - No CPU fallback exists anywhere: an unsupported request is an error.
 *   - Activations are fp32.  Internal activations are NHWC; the network input is NCHW like the
 *     reference's (engine.py:25).
 */
#ifndef SCOUTER_B200_H_
#define SCOUTER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCOUTER_ABI_VERSION 3

#define SCOUTER_OK 0
#define SCOUTER_E_INVALID (-1)     /* bad argument (null pointer, non-positive size, ...) */
#define SCOUTER_E_UNSUPPORTED (-2) /* well-formed but outside what the kernels implement */
#define SCOUTER_E_STATE (-3)       /* call order violated (plan not bound, ...) */

typedef void* scouter_stream_t; /* cudaStream_t */

int scouter_abi_version(void);
const char* scouter_last_error(void);
/* Compiled-for architecture (100 for sm_100a) and whether the running device matches. */
int scouter_device_check(int device);

/* ------------------------------------------------------------------------------------------------
 * a7  PositionEmbeddingSine.forward  (sloter/utils/position_encode.py:26-46 as built by :77-81:
 *     num_pos_feats = d/2, normalize=True, scale 2*pi, temperature 1e4).
 *     Writes the input-independent table token-major: pe[j*d + c], j = y*w + x -- the layout
 *     slot_model.py:113-115 produces with reshape(b,d,-1).permute(0,2,1).
 * ---------------------------------------------------------------------------------------------- */
int scouter_pe_sine(float* pe /* (h*w, d) */, int d, int h, int w, scouter_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a8/a9  SlotAttention  (sloter/utils/slot_attention.py:10-96)
 * ---------------------------------------------------------------------------------------------- */
#define SCOUTER_MAX_TO_K_LAYERS 8

typedef struct scouter_xslot_desc {
    int32_t d;               /* hidden dim; only 64 is implemented (train.py default hidden_dim) */
    int32_t num_classes;     /* C */
    int32_t slots_per_class; /* spc; S = C * spc slots, consecutive slots belong to one class */
    int32_t to_k_layers;     /* number of Linear layers in to_k (slot_attention.py:30-37) */
    int32_t iters;           /* attention evaluations; reference hard default 3 (:10) */
    int32_t loss_status;     /* +1 / -1 (:96) */
    float power;             /* attn-loss exponent (:96) */
    /* Parameters in the reference's own (PyTorch) layouts: */
    const float* initial_slots;                     /* (S, d)   slot.initial_slots */
    const float* to_k_w[SCOUTER_MAX_TO_K_LAYERS];   /* (d, d) [out][in]  slot.to_k.{0,2,4,..}.weight */
    const float* to_k_b[SCOUTER_MAX_TO_K_LAYERS];   /* (d)               slot.to_k.{0,2,4,..}.bias   */
    const float* gru_w_ih;                          /* (3d, d) rows [r|z|n]  slot.gru.weight_ih_l0 */
    const float* gru_w_hh;                          /* (3d, d)               slot.gru.weight_hh_l0 */
    const float* gru_b_ih;                          /* (3d)                  slot.gru.bias_ih_l0   */
    const float* gru_b_hh;                          /* (3d)                  slot.gru.bias_hh_l0   */
} scouter_xslot_desc_t;

/* Kernel-friendly copy of the parameters (transposed to_k / GRU matrices, hi/lo tf32 splits).
 * Caller allocates `scouter_xslot_packed_bytes` bytes on the device and keeps the buffer alive
 * while it is used; re-pack after the parameters change (optimizer step, load_state_dict). */
size_t scouter_xslot_packed_bytes(const scouter_xslot_desc_t* desc);
int scouter_xslot_pack(const scouter_xslot_desc_t* desc, void* packed, scouter_stream_t stream);

typedef struct scouter_xslot_io {
    int32_t batch;           /* B */
    int32_t n;               /* tokens per image (h*w) */
    /* inputs_x: post-ReLU features, element (b,j,e) at x[b*x_sb + j*x_sn + e*x_sd] (strides in
     * floats; the reference passes permuted views, slot_model.py:113-115). */
    const float* x;
    int64_t x_sb, x_sn, x_sd;
    /* inputs (= x + PE).  Either give x_pe with its own strides, or leave x_pe NULL and give the
     * (n, d) table `pe` from scouter_pe_sine: the kernel then forms x + pe itself. */
    const float* x_pe;
    int64_t xpe_sb, xpe_sn, xpe_sd;
    const float* pe;
    /* outputs (any may be NULL except logits) */
    float* logits;           /* (B, C)  = loss_status * sum_d updates           (:96) */
    float* attn;             /* (B, S, n) final-iteration attention             (:57) */
    float* attn_sum;         /* (B) per-image sum of the final attention -- the cross-image part of
                                the area loss is finished by scouter_head_finalize */
} scouter_xslot_io_t;

/* slot_attention.py:44-96 up to (not including) the batch mean of the loss. */
size_t scouter_xslot_workspace_bytes(const scouter_xslot_desc_t* desc, int batch, int n);
int scouter_xslot_forward(const scouter_xslot_desc_t* desc, const void* packed, const scouter_xslot_io_t* io,
                          void* workspace, size_t workspace_bytes, scouter_stream_t stream);

/* a11 + the batch reduction of a9:  log_softmax (slot_model.py:117), attn_loss =
 * (sum_b attn_sum / (B*S*n))^power (slot_attention.py:93-96), nll_loss and
 * loss = nll + lambda * attn_loss (slot_model.py:119-122).  `target` (int64) may be NULL, then
 * nll/loss are not written.  Scalars are written to losses[0..2] = {loss, nll, attn_loss}. */
int scouter_head_finalize(const float* logits, const float* attn_sum, const int64_t* target, int batch,
                          int num_classes, int num_slots, int n, float power, float lambda_value,
                          float* log_probs /* (B,C) */, float* losses /* (3) or NULL */,
                          scouter_stream_t stream);

/* a10  the vis branch (slot_attention.py:68-80): per-class sum of attn[vis_id], joint min-max over
 * (C, n), *255, truncation to uint8.  Output (C, n) uint8 on the device. */
int scouter_vis_maps_u8(const float* attn /* (B,S,n) */, int batch, int num_classes, int slots_per_class, int n,
                        int vis_id, uint8_t* maps, scouter_stream_t stream);

/* f3  the explanation output path (test.py:33-35, 40-44): the PNG round trip
 *     np.array(Image.open('sloter/vis/slot_{id}.png').resize(image_raw.size, resample=Image.BILINEAR))
 * on the device.  `maps` = `count` uint8 maps of h x w (the output of scouter_vis_maps_u8, h*w = n); `out` =
 * (count, out_h, out_w) uint8, bit-identical to Pillow's 8-bit bilinear resampler (Resample.c: triangle filter of
 * support max(1, in/out), 22-bit fixed-point coefficients, horizontal then vertical pass with a uint8 image in
 * between); `ratios` = (count) float64 attention ratios sum(map)/(h*w*255) of the un-resized maps (test.py:43).
 * Either output may be NULL.  Down-scaling works too (same filter); no workspace, graph-capturable. */
int scouter_vis_upsample_u8(const uint8_t* maps /* (count,h,w) */, int count, int h, int w, int out_h, int out_w,
                            uint8_t* out, double* ratios, scouter_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a6 + a7 + a9  the fused xSlot head: conv1x1 + ReLU (slot_model.py:108-109), + PE (:110-111),
 * SlotAttention (:116).  `feat` is the backbone output.
 * ---------------------------------------------------------------------------------------------- */
#define SCOUTER_LAYOUT_NHWC 0 /* (B, h*w, ch): this library's backbone output */
#define SCOUTER_LAYOUT_NCHW 1 /* (B, ch, h*w): the reference backbone's flattened output (slot_model.py:108) */

#define SCOUTER_MATH_FP32 0    /* CUDA-core fp32 FMA everywhere (exact mode) */
#define SCOUTER_MATH_TC 1      /* tcgen05 tensor cores, error-compensated products on fp32 data (convs: fp16 main + two 16-bit
                                  corrections; head: tf32 main + bf16 corrections) -- fp32-class results; default */
#define SCOUTER_MATH_TC_FAST 2 /* tcgen05 tensor cores, single tf32 pass on tf32-rounded activations/weights
                                  (cuDNN-TF32 class accuracy: ~3e-3 on the log-probs; opt-in) */

typedef struct scouter_head_io {
    int32_t batch, h, w;     /* feature map is h x w (feature_size; derived from the input, D6) */
    int32_t channel;         /* ch: backbone channels (args.channel) */
    int32_t layout;          /* SCOUTER_LAYOUT_* of feat */
    int32_t math;            /* SCOUTER_MATH_* */
    const float* feat;
    const float* conv_w;     /* (d, ch)  conv1x1.weight viewed 2-D */
    const float* conv_b;     /* (d)      conv1x1.bias */
    const float* pe;         /* (h*w, d) from scouter_pe_sine */
    float* logits;           /* (B, C) */
    float* attn;             /* (B, S, n) or NULL */
    float* attn_sum;         /* (B) or NULL */
    float* x_out;            /* (B, n, d) projected features, or NULL (debug / tests) */
    const void* conv_w_split;/* optional (2d, ch) bfloat16: rows [0,d) = bf16(W), rows [d,2d) = bf16(W - trunc19(W)), the
                                correction operands of the error-compensated product; NULL = derived per call */
} scouter_head_io_t;

size_t scouter_head_workspace_bytes(const scouter_xslot_desc_t* desc, const scouter_head_io_t* io);
/* Kernels one scouter_head_forward call launches for this geometry: 1 when the whole head runs as the fused kernel
 * (a unit of images fits one 128-row tensor-core tile and S <= 32), else projection + loop (+ layout copy). */
int scouter_head_launch_count(const scouter_xslot_desc_t* desc, const scouter_head_io_t* io);
int scouter_head_forward(const scouter_xslot_desc_t* desc, const void* packed, const scouter_head_io_t* io,
                         void* workspace, size_t workspace_bytes, scouter_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a2-a5  backbone: an op program executed by the library.
 *
 * The Python side walks the module tree (the mirror of timm's ResNet / ResNestBottleneck /
 * SplitAttnConv2d / BasicBlock), folds eval-mode BatchNorm into each conv and
 * describes the network as a flat list of ops over numbered activation buffers.  The library does
 * shape inference, buffer placement in one caller-provided arena, kernel selection and launch.
 * Buffer 0 is the network input (NCHW, caller pointer given at run time).
 * ---------------------------------------------------------------------------------------------- */
enum scouter_op_kind {
    SCOUTER_OP_STEM_CONV = 1,  /* NCHW in (cin <= 4) -> NHWC out; k x k, stride, pad; +bias, optional ReLU.
                                  timm/models/resnet.py:401 (deep stem conv1.0) / :410 / slot_model.py:23-24 */
    SCOUTER_OP_CONV = 2,       /* NHWC conv, groups >= 1, +bias [+residual src2] [ReLU].
                                  resnet.py:403-408, resnest.py:84,103, split_attn.py:43-45, resnet.py:153-161,304 */
    SCOUTER_OP_MAXPOOL = 3,    /* k/stride/pad max-pool (resnet.py:420) */
    SCOUTER_OP_AVGPOOL = 4,    /* avg-pool; flags select ceil_mode / count_include_pad (resnest.py:101, resnet.py:300) */
    SCOUTER_OP_SPLAT_GAP = 5,  /* (B,H,W,2C) -> (B,C): mean_hw of the radix sum (split_attn.py:62-68) */
    /* 6 is retired: fc1(+bn1)+ReLU and fc2 (split_attn.py:69-73) are ordinary SCOUTER_OP_CONVs on (B,1,1,C) */
    SCOUTER_OP_SPLAT_APPLY = 7,/* r-softmax of the fc2 output over the radix pair (split_attn.py:14-28,74), then
                                  sum_r x_r * a_r [+ fused avd AvgPool(3,2,1)] (split_attn.py:76-79, resnest.py:101) */
    SCOUTER_OP_GAP = 8,        /* global average pool (B,H,W,C) -> (B,1,1,C)  (no-slot classifier path) */
    SCOUTER_OP_TO_NCHW = 9     /* NHWC -> NCHW copy for callers that want the reference's flattened layout */
};

#define SCOUTER_F_RELU 1
#define SCOUTER_F_RESIDUAL 2          /* add buffer src2 before the ReLU */
#define SCOUTER_F_CEIL_MODE 4
#define SCOUTER_F_COUNT_INCLUDE_PAD 8
#define SCOUTER_F_AVD_POOL 16         /* SPLAT_APPLY: follow with AvgPool2d(3, 2, padding=1) */
#define SCOUTER_F_TF32_1PASS 32      /* per-stage precision policy: in a SCOUTER_MATH_TC plan run THIS op as SCOUTER_MATH_TC_FAST
                                        would (single tf32 pass, tf32-rounded output; the caller passes tf32-rounded weights and
                                        no w2).  Ignored by the other math modes. */

typedef struct scouter_op {
    int32_t kind;
    int32_t src, src2, dst;   /* buffer ids; src2 = residual (CONV) / fc2 output (B,2C) (SPLAT_APPLY) / -1 */
    int32_t cin, cout;
    int32_t kh, kw, stride, pad, groups;
    int32_t flags;
    int32_t mid;              /* unused (0) */
    int32_t reserved;
    /* Folded parameters, device pointers, fp32:
     *   CONV / STEM_CONV: w = (cout, kh, kw, cin/groups) "OHWI", b = (cout); for SCOUTER_MATH_TC an optional
     *     w2 = 16-bit [fp16(W) ; bf16(W - fp16(W))] ((2*cout, kh, kw, cin/groups) halves: the pre-split operand of the
     *     error-compensated fp16 product, scouter_b200/plan.py split_weights_f16) lets the kernels skip the on-the-fly
     *     weight split and keep the activation operand in tensor memory
     */
    const float* w;
    const float* b;
    const float* w2;
    const float* b2;
} scouter_op_t;

typedef struct scouter_plan scouter_plan_t;

/* Copies the op list (not the weights: those stay caller-owned and must outlive the plan). */
int scouter_plan_create(const scouter_op_t* ops, int n_ops, int n_buffers, int math, scouter_plan_t** out);
void scouter_plan_destroy(scouter_plan_t* plan);

/* Shape inference + arena layout for input (batch, cin, h, w).  After a successful bind,
 * scouter_plan_arena_bytes / scouter_plan_buffer_shape are valid.  Re-binding is allowed. */
int scouter_plan_bind(scouter_plan_t* plan, int batch, int cin, int h, int w);
size_t scouter_plan_arena_bytes(const scouter_plan_t* plan);
/* shape[4] = {B, H, W, C} of a buffer (NHWC logical dims; buffer 0 reports the NCHW input as B,H,W,C too). */
int scouter_plan_buffer_shape(const scouter_plan_t* plan, int buffer, int32_t shape[4]);
/* Byte offset of a buffer inside the arena (buffer 0 has none: returns (size_t)-1). */
size_t scouter_plan_buffer_offset(const scouter_plan_t* plan, int buffer);

/* Runs every op in order.  `arena` must hold scouter_plan_arena_bytes bytes, 1024-byte aligned. */
int scouter_plan_run(scouter_plan_t* plan, const float* input_nchw, void* arena, size_t arena_bytes,
                     scouter_stream_t stream);
/* Number of kernel launches one scouter_plan_run issues (for bench.py's gpu_launches). */
int scouter_plan_launch_count(const scouter_plan_t* plan);

/* One SCOUTER_OP_CONV outside a plan (unit tests of the kernels; same dispatch as scouter_plan_run):
 * in (B,H,W,cin) NHWC, res (B,Ho,Wo,cout) or NULL, out (B,Ho,Wo,cout).  scouter_conv_path reports which
 * kernel family the dispatch picks: 2 = tcgen05 halo 3x3, 1 = tcgen05 flat / tap-reload, 0 = CUDA-core fp32. */
int scouter_conv_forward(const scouter_op_t* op, const float* in, const float* res, float* out, int batch, int h, int w,
                         int math, scouter_stream_t stream);
int scouter_conv_path(const scouter_op_t* op, int batch, int h, int w, int math);
/* The network's first conv (SCOUTER_OP_STEM_CONV geometry: NCHW input with 1..4 channels -> NHWC) outside a plan; used
 * by the train-mode forward, where BatchNorm is not folded (resnet.py:401, slot_model.py:23-24). */
int scouter_stem_conv_forward(const scouter_op_t* op, const float* in_nchw, float* out, int batch, int h, int w,
                              scouter_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * f2  input pipeline boundary (engine.py:25, dataset/transform_func.py:52-67 ToTensor, :87-94 Normalize,
 *     normalisation constants :101-106): uint8 HWC images -> the fp32 NCHW batch the model consumes.
 *     out[b,c,y,x] = (float)(((double)img[b,y,x,c] / 255 - mean[c]) / std[c])  -- evaluated in fp64 and rounded once,
 *     exactly like the reference (its ToTensor yields float64; engine.py casts to float32 on the device).
 *     `mean`/`std` are HOST arrays of `c` doubles (c <= 4).  4x fewer H2D bytes than shipping fp32 images.
 * ---------------------------------------------------------------------------------------------- */
int scouter_preprocess_u8(const uint8_t* img_nhwc, int batch, int h, int w, int c, const double* mean_host,
                          const double* std_host, float* out_nchw, scouter_stream_t stream);


/* ------------------------------------------------------------------------------------------------
 * f1  the training step (engine.py:28-35 `loss.backward(); optimizer.step()`, train.py:140-148): train-mode
 *     BatchNorm, the backward of every op on the path, AdamW.  Each entry is one stream-ordered launch group over
 *     caller-owned NHWC fp32 buffers; the step itself (op order, buffer lifetimes, gradient accumulation flags) is host
 *     logic in scouter_b200/train.py, driven by the same op program as the forward.  These are the correctness-first
 *     CUDA-core kernels of csrc/draft/ (validated on B200 against the reference's loss.backward(), memcheck + racecheck
 *     clean); the tensor-core versions of the conv gradients replace them behind the same entries.
 *     Gradient outputs marked "accumulated" must be zeroed (or hold a running sum) on entry.
 * ---------------------------------------------------------------------------------------------- */
typedef struct scouter_bn_train_args {   /* nn.BatchNorm2d in train mode (eps 1e-5, momentum 0.1): timm/models/resnet.py:401-420 */
    long long M; int C;                  /* x viewed as (M = B*H*W, C); C % 4 == 0 */
    const float* x;
    double* sums;                        /* (C, 2) workspace, zeroed by the call */
    const float *gamma, *beta;
    float *running_mean, *running_var;   /* updated in place (unbiased variance, momentum) */
    float *scale, *shift;                /* (C) workspace */
    float *save_mean, *save_rstd;        /* (C) batch statistics kept for the backward, or NULL */
    float eps, momentum;
    const float* residual;               /* (M, C) added before the ReLU, or NULL */
    float* y;                            /* (M, C); may alias x */
    int relu;
} scouter_bn_train_args_t;
int scouter_train_bn_forward(const scouter_bn_train_args_t* a, scouter_stream_t stream);

typedef struct scouter_bn_bwd_args {     /* backward of out = [relu](bn(x) [+ residual]) */
    long long M; int C;
    const float *x, *out, *d_out;        /* conv output, block output (ReLU mask; may be NULL when !relu), its gradient */
    const float *gamma, *save_mean, *save_rstd;
    double* sums;                        /* (C, 2) workspace, zeroed by the call */
    float *d_gamma, *d_beta;             /* (C) accumulated */
    float* coef;                         /* (C, 3) workspace */
    float* dx;                           /* (M, C); may alias d_out */
    float* d_residual;                   /* (M, C) or NULL: receives the masked gradient */
    int relu;
} scouter_bn_bwd_args_t;
int scouter_train_bn_backward(const scouter_bn_bwd_args_t* a, scouter_stream_t stream);

typedef struct scouter_wgrad_args {      /* nn.Conv2d weight / bias gradient; dW in this library's OHWI layout */
    int B, H, W, Cin, Ho, Wo, Cout, k, stride, pad, groups;
    const float *x, *dy;                 /* (B,H,W,Cin), (B,Ho,Wo,Cout) */
    float* dw;                           /* (Cout,k,k,Cin/groups) accumulated */
    float* db;                           /* (Cout) accumulated, or NULL */
} scouter_wgrad_args_t;
int scouter_train_conv_wgrad(const scouter_wgrad_args_t* a, scouter_stream_t stream);

typedef struct scouter_dgrad_args {      /* nn.Conv2d data gradient, any stride (gather form) */
    int B, H, W, Cin, Ho, Wo, Cout, k, stride, pad, groups;
    const float *dy, *w;                 /* (B,Ho,Wo,Cout), (Cout,k,k,Cin/groups) */
    float* dx;                           /* (B,H,W,Cin) overwritten */
} scouter_dgrad_args_t;
int scouter_train_conv_dgrad(const scouter_dgrad_args_t* a, scouter_stream_t stream);

typedef struct scouter_pool_bwd_args {   /* kind 0: MaxPool2d(3,2,1) resnet.py:420; 1: AvgPool2d(2,2,ceil,count_include_pad=False) */
    int B, H, W, C, Ho, Wo;              /*      resnet.py:300; 2: AvgPool2d(3,2,1) resnest.py:101 */
    const float* x;                      /* forward input (kind 0 only) */
    const float* dy; float* dx;
} scouter_pool_bwd_args_t;
int scouter_train_pool_backward(const scouter_pool_bwd_args_t* a, int kind, scouter_stream_t stream);
/* forward of the same three pools outside a plan (the train-mode forward keeps every buffer alive for the backward) */
int scouter_pool_forward(int kind, const float* in, float* out, int batch, int h, int w, int c, scouter_stream_t stream);

typedef struct scouter_splat_bwd_args {  /* split attention, split_attn.py:62-79 (radix 2) */
    int B, HW, C;
    const float *x2, *d_out, *att;       /* (B,HW,2C), (B,HW,C), softmax over the radix (B,2,C) */
    float *d_att, *d_logit;              /* stage 0 outputs (B,2,C) */
    const float* d_gap;                  /* stage 1 input (B,C) */
    float* d_x2;                         /* stage 1 output (B,HW,2C) */
} scouter_splat_bwd_args_t;
int scouter_train_splat_backward(const scouter_splat_bwd_args_t* a, int stage, scouter_stream_t stream);
/* forward halves outside a plan: gap = mean_hw(x_0 + x_1) (scratch: scouter_splat_gap_scratch_floats floats);
 * out = x_0 * a_0 + x_1 * a_1 with a = softmax over the radix pair of `logits` (B, 2C) */
size_t scouter_splat_gap_scratch_floats(int batch, int hw, int c);
int scouter_splat_gap_forward(const float* in, float* scratch, float* gap, int batch, int hw, int c, scouter_stream_t stream);
int scouter_splat_apply_forward(const float* in, const float* logits, float* out, int batch, int h, int w, int c,
                                scouter_stream_t stream);

typedef struct scouter_head_bwd_args {   /* backward of slot_model.py:108-116 + slot_attention.py:44-96 (recomputes the forward) */
    int B, n, ch, S, C, spc, L, iters, loss_status;
    const float* feat;                   /* (B, n, ch) token-major backbone output */
    const float *conv_w, *conv_b, *pe;   /* (64, ch), (64), (n, 64) */
    const float *to_k_w[SCOUTER_MAX_TO_K_LAYERS], *to_k_b[SCOUTER_MAX_TO_K_LAYERS];
    const float *w_ih, *w_hh, *b_ih, *b_hh;
    const float* slots0;                 /* (S, 64) */
    const float* g_logits;               /* (B, C) gradient at the logits */
    const float* attn_coef;              /* device scalar: g_attn_loss * power * mean(attn)^(power-1) / (B*S*n) */
    float* d_feat;                       /* (B, n, ch) or NULL */
    float* d_pre;                        /* (B, n, 64) or NULL */
    float *g_conv_w, *g_conv_b, *g_to_k_w[SCOUTER_MAX_TO_K_LAYERS], *g_to_k_b[SCOUTER_MAX_TO_K_LAYERS];
    float *g_w_ih, *g_w_hh, *g_b_ih, *g_b_hh, *g_slots0;     /* accumulated */
    float* scratch;                      /* B * scouter_train_head_backward_scratch_floats(...) floats */
    size_t scratch_per_image;
} scouter_head_bwd_args_t;
size_t scouter_train_head_backward_scratch_floats(int n, int s, int to_k_layers, int iters);
int scouter_train_head_backward(const scouter_head_bwd_args_t* a, scouter_stream_t stream);

typedef struct scouter_adamw_args {      /* torch.optim.AdamW single-tensor step (train.py:146, engine.py:34); scalars from the host */
    float decay;                         /* 1 - lr * weight_decay */
    float one_minus_beta1, beta2, one_minus_beta2, eps;
    float step_size;                     /* lr / (1 - beta1^t) */
    float bias_correction2_sqrt;         /* sqrt(1 - beta2^t) */
    size_t n;
    float* p; const float* g; float *m, *v;
} scouter_adamw_args_t;
int scouter_train_adamw_step(const scouter_adamw_args_t* a, scouter_stream_t stream);

/* The pre-split weight operand `w2` of scouter_op_t for SCOUTER_MATH_TC, made on the device in ONE launch: for n fp32 weights
 * out16[0..n) = fp16(w) (round to nearest, saturating) and out16[n..2n) = bf16(w - fp16(w)).  The training step re-splits every
 * conv's weights each step (the parameters change); scouter_b200/plan.py split_weights_f16 is the torch-op equivalent the
 * eval path runs once per parameter version. */
int scouter_split_weights_f16(const float* w, void* out16, size_t n, scouter_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a1  SlotModel.forward from HOST buffers in one call (sloter/slot_model.py:105-127 as driven by
 *     engine.py:25-30: H2D copy of the batch, forward, read-back of the result).
 *     cudaMemcpyAsync(input_host -> input_dev), scouter_plan_run, scouter_head_forward on the plan's
 *     `feat_buffer`, scouter_head_finalize, cudaMemcpyAsync(log_probs/losses -> host), then
 *     cudaStreamSynchronize(stream): on return the host outputs are valid.  Host buffers should be
 *     page-locked.  This is the call bench.py times for its end-to-end number.
 * ---------------------------------------------------------------------------------------------- */
typedef struct scouter_forward_host_args {
    scouter_plan_t* plan;             /* bound backbone plan */
    const scouter_xslot_desc_t* desc;
    const void* packed;
    scouter_head_io_t head;           /* .feat is ignored (taken from the plan); the rest as for scouter_head_forward */
    int32_t feat_buffer;              /* plan buffer holding the NHWC features */
    int32_t reserved;
    const float* input_host;          /* (B, Cin, H, W) fp32 */
    float* input_dev;                 /* device staging buffer of the same size */
    size_t input_bytes;
    void* arena; size_t arena_bytes;
    void* head_workspace; size_t head_workspace_bytes;
    const int64_t* target_dev;        /* (B) or NULL */
    float lambda_value;
    float reserved2;
    float* log_probs_dev;             /* (B, C) */
    float* losses_dev;                /* (3) or NULL */
    float* log_probs_host;            /* (B, C) */
    float* losses_host;               /* (3) or NULL */
    scouter_stream_t stream;
} scouter_forward_host_args_t;

int scouter_forward_host(const scouter_forward_host_args_t* args);

#ifdef __cplusplus
}
#endif
#endif /* SCOUTER_B200_H_ */
