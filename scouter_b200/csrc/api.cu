// C-ABI glue: error reporting, the backbone op-program executor (shape inference, arena placement,
// kernel selection) and the fused-head entry point.  See include/scouter_b200.h for the contract.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"
#include "umma.cuh"
#include "xslot.cuh"

namespace scouter {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int xslot_loop_launch(const scouter_xslot_desc_t* d, const void* packed, const scouter_xslot_io_t* io, cudaStream_t s);

struct Buf {
    int B = 0, H = 0, W = 0, C = 0;
    bool defined = false;
    size_t offset = (size_t)-1, bytes = 0;
    int first_def = -1, last_use = -1;
};

}  // namespace scouter

using namespace scouter;

struct scouter_plan {
    std::vector<scouter_op_t> ops;
    std::vector<Buf> bufs;
    int math = SCOUTER_MATH_FP32;
    bool bound = false;
    size_t arena_bytes = 0;
    size_t scratch_off = 0;  // shared scratch at the end of the arena (split-attention GAP partial sums)
    int launches = 0;
    std::vector<UmmaConvPlan> umma;  // per op; .valid says whether the tcgen05 kernel takes it
    std::vector<int> gap_slots;      // per op: > 0 on a conv whose epilogue writes the GAP partial sums of the SPLAT_GAP op that
                                     // follows it, and on that SPLAT_GAP op (which then only runs the finish kernel)
    std::vector<std::vector<float>> host_w, host_b;  // per op: host copies of small stem filter banks
};

extern "C" int scouter_abi_version(void) { return SCOUTER_ABI_VERSION; }
extern "C" const char* scouter_last_error(void) { return g_err; }

extern "C" int scouter_device_check(int device) {
    cudaDeviceProp prop;
    SC_CUDA(cudaGetDeviceProperties(&prop, device));
    SC_CHECK_ARG(prop.major == 10, SCOUTER_E_UNSUPPORTED,
                 "device %d is sm_%d%d; libscouter_b200 carries sm_100a code only", device, prop.major, prop.minor);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Plan
// ------------------------------------------------------------------------------------------------
extern "C" int scouter_plan_create(const scouter_op_t* ops, int n_ops, int n_buffers, int math, scouter_plan_t** out) {
    SC_CHECK_ARG(ops && out && n_ops > 0 && n_buffers > 1, SCOUTER_E_INVALID, "plan_create: bad arguments");
    SC_CHECK_ARG(math >= SCOUTER_MATH_FP32 && math <= SCOUTER_MATH_TC_FAST, SCOUTER_E_INVALID, "plan_create: math=%d", math);
    for (int i = 0; i < n_ops; ++i) {
        const scouter_op_t& o = ops[i];
        SC_CHECK_ARG(o.kind >= SCOUTER_OP_STEM_CONV && o.kind <= SCOUTER_OP_TO_NCHW && o.kind != 6, SCOUTER_E_INVALID,
                     "plan_create: op %d has unknown kind %d", i, o.kind);
        SC_CHECK_ARG(o.src >= 0 && o.src < n_buffers && o.dst > 0 && o.dst < n_buffers && o.src2 < n_buffers,
                     SCOUTER_E_INVALID, "plan_create: op %d references a buffer outside [0,%d)", i, n_buffers);
        if (o.kind == SCOUTER_OP_STEM_CONV || o.kind == SCOUTER_OP_CONV)
            SC_CHECK_ARG(o.w != nullptr, SCOUTER_E_INVALID, "plan_create: op %d (conv) has no weights", i);
    }
    scouter_plan* p = new (std::nothrow) scouter_plan();
    SC_CHECK_ARG(p, SCOUTER_E_INVALID, "plan_create: out of host memory");
    p->ops.assign(ops, ops + n_ops);
    p->bufs.resize(n_buffers);
    p->math = math;
    p->host_w.resize(n_ops);
    p->host_b.resize(n_ops);
    for (int i = 0; i < n_ops; ++i) {
        const scouter_op_t& o = ops[i];
        // stem filter banks (<= 864 floats) are kept on the host too: they are passed by value in the kernel parameters
        if (o.kind == SCOUTER_OP_STEM_CONV && o.kh == 3 && o.kw == 3 && (size_t)o.cout * 9 * o.cin <= 1024) {
            p->host_w[i].resize((size_t)o.cout * 9 * o.cin);
            cudaError_t e = cudaMemcpy(p->host_w[i].data(), o.w, p->host_w[i].size() * sizeof(float), cudaMemcpyDeviceToHost);
            if (e == cudaSuccess && o.b) {
                p->host_b[i].resize(o.cout);
                e = cudaMemcpy(p->host_b[i].data(), o.b, o.cout * sizeof(float), cudaMemcpyDeviceToHost);
            }
            if (e != cudaSuccess) {  // e.g. no device (CPU-only shape inference): fall back to the generic stem kernel
                (void)cudaGetLastError();
                p->host_w[i].clear();
                p->host_b[i].clear();
            }
        }
    }
    *out = p;
    return 0;
}

extern "C" void scouter_plan_destroy(scouter_plan_t* plan) { delete plan; }

static int pool_out(int H, int k, int s, int p, bool ceil_mode) {
    if (!ceil_mode) return (H + 2 * p - k) / s + 1;
    int o = (H + 2 * p - k + s - 1) / s + 1;
    if ((o - 1) * s >= H + p) --o;  // the last window must start inside the input or left padding
    return o;
}

extern "C" int scouter_plan_bind(scouter_plan_t* plan, int batch, int cin, int h, int w) {
    SC_CHECK_ARG(plan && batch > 0 && cin > 0 && h > 0 && w > 0, SCOUTER_E_INVALID, "plan_bind: bad arguments");
    plan->bound = false;
    // the kernels index elements with 32-bit integers and put the image index in gridDim.z: refuse (loudly) what
    // would overflow instead of computing garbage -- the caller splits the batch
    SC_CHECK_ARG(batch <= 65535, SCOUTER_E_UNSUPPORTED, "plan_bind: batch %d > 65535 images per call", batch);
    SC_CHECK_ARG((long long)batch * cin * h * w < (1ll << 31), SCOUTER_E_UNSUPPORTED,
                 "plan_bind: input of %d x %d x %d x %d has 2^31 or more elements", batch, cin, h, w);
    for (auto& b : plan->bufs) b = Buf();
    Buf& in = plan->bufs[0];
    in.B = batch; in.H = h; in.W = w; in.C = cin; in.defined = true; in.first_def = -1;
    const int n_ops = (int)plan->ops.size();
    for (int i = 0; i < n_ops; ++i) {
        const scouter_op_t& o = plan->ops[i];
        const Buf s = plan->bufs[o.src];
        SC_CHECK_ARG(s.defined, SCOUTER_E_STATE, "plan_bind: op %d reads buffer %d before it is written", i, o.src);
        Buf d;
        d.B = s.B;
        switch (o.kind) {
            case SCOUTER_OP_STEM_CONV:
            case SCOUTER_OP_CONV:
                SC_CHECK_ARG(s.C == o.cin, SCOUTER_E_INVALID, "plan_bind: op %d expects %d input channels, buffer %d has %d",
                             i, o.cin, o.src, s.C);
                SC_CHECK_ARG(o.stride >= 1 && o.kh >= 1 && o.kw >= 1 && o.groups >= 1, SCOUTER_E_INVALID, "plan_bind: op %d geometry", i);
                d.H = (s.H + 2 * o.pad - o.kh) / o.stride + 1;
                d.W = (s.W + 2 * o.pad - o.kw) / o.stride + 1;
                d.C = o.cout;
                break;
            case SCOUTER_OP_MAXPOOL:
                d.H = pool_out(s.H, o.kh, o.stride, o.pad, false);
                d.W = pool_out(s.W, o.kw, o.stride, o.pad, false);
                d.C = s.C;
                break;
            case SCOUTER_OP_AVGPOOL:
                d.H = pool_out(s.H, o.kh, o.stride, o.pad, o.flags & SCOUTER_F_CEIL_MODE);
                d.W = pool_out(s.W, o.kw, o.stride, o.pad, o.flags & SCOUTER_F_CEIL_MODE);
                d.C = s.C;
                break;
            case SCOUTER_OP_SPLAT_GAP:
                SC_CHECK_ARG(s.C == 2 * o.cout, SCOUTER_E_INVALID, "plan_bind: op %d (splat gap) wants %d channels, got %d", i, 2 * o.cout, s.C);
                d.H = d.W = 1; d.C = o.cout;
                break;
            case SCOUTER_OP_SPLAT_APPLY:
                SC_CHECK_ARG(s.C == 2 * o.cout && o.src2 >= 0, SCOUTER_E_INVALID, "plan_bind: op %d (splat apply) channel mismatch", i);
                if (o.flags & SCOUTER_F_AVD_POOL) { d.H = pool_out(s.H, 3, 2, 1, false); d.W = pool_out(s.W, 3, 2, 1, false); }
                else { d.H = s.H; d.W = s.W; }
                d.C = o.cout;
                break;
            case SCOUTER_OP_GAP:
                d.H = d.W = 1; d.C = s.C;
                break;
            case SCOUTER_OP_TO_NCHW:
                d.H = s.H; d.W = s.W; d.C = s.C;
                break;
        }
        SC_CHECK_ARG(d.H > 0 && d.W > 0, SCOUTER_E_INVALID, "plan_bind: op %d produces an empty %dx%d map (input too small)", i, d.H, d.W);
        SC_CHECK_ARG((long long)d.B * d.H * d.W * d.C < (1ll << 31), SCOUTER_E_UNSUPPORTED,
                     "plan_bind: op %d produces %d x %d x %d x %d = 2^31 or more elements (split the batch)", i, d.B, d.H, d.W, d.C);
        if (o.src2 >= 0) {
            const Buf& r = plan->bufs[o.src2];
            SC_CHECK_ARG(r.defined, SCOUTER_E_STATE, "plan_bind: op %d reads buffer %d before it is written", i, o.src2);
            if (o.kind == SCOUTER_OP_CONV && (o.flags & SCOUTER_F_RESIDUAL))
                SC_CHECK_ARG(r.H == d.H && r.W == d.W && r.C == d.C, SCOUTER_E_INVALID,
                             "plan_bind: op %d residual buffer %d is %dx%dx%d, output is %dx%dx%d", i, o.src2, r.H, r.W, r.C, d.H, d.W, d.C);
        }
        SC_CHECK_ARG(!plan->bufs[o.dst].defined, SCOUTER_E_INVALID, "plan_bind: buffer %d written twice (op %d)", o.dst, i);
        d.defined = true;
        d.first_def = i;
        d.bytes = align_up((size_t)d.B * d.H * d.W * d.C * sizeof(float), 1024);
        plan->bufs[o.dst] = d;
        plan->bufs[o.src].last_use = i;
        if (o.src2 >= 0) plan->bufs[o.src2].last_use = i;
    }
    // Arena placement: first-fit over live intervals; buffers nobody reads are outputs and stay live.
    struct Live { size_t off, bytes; int buf; };
    std::vector<Live> live;
    size_t high = 0;
    for (int i = 0; i < n_ops; ++i) {
        const int dst = plan->ops[i].dst;
        Buf& d = plan->bufs[dst];
        std::sort(live.begin(), live.end(), [](const Live& a, const Live& b) { return a.off < b.off; });
        size_t off = 0;
        for (const Live& l : live) {
            if (off + d.bytes <= l.off) break;
            off = std::max(off, l.off + l.bytes);
        }
        d.offset = off;
        high = std::max(high, off + d.bytes);
        live.push_back({off, d.bytes, dst});
        // release buffers whose last reader is this op (never the one just written)
        live.erase(std::remove_if(live.begin(), live.end(), [&](const Live& l) {
                       const Buf& b = plan->bufs[l.buf];
                       return l.buf != dst && b.last_use >= 0 && b.last_use <= i;
                   }), live.end());
    }
    size_t scratch = 0;
    int launches = 0;
    plan->gap_slots.assign(n_ops, 0);
    static const bool no_gap_fuse = getenv("SCOUTER_NO_GAP_FUSE") != nullptr;
    for (int i = 0; i < n_ops; ++i) {
        const scouter_op_t& o = plan->ops[i];
        ++launches;
        if (o.kind == SCOUTER_OP_SPLAT_GAP) {
            const Buf& sb = plan->bufs[o.src];
            // the conv right before it writes this buffer: if the halo kernel takes it, its epilogue delivers the partial sums
            int slots = 0;
            if (i > 0 && !no_gap_fuse && plan->math == SCOUTER_MATH_TC && plan->ops[i - 1].kind == SCOUTER_OP_CONV &&
                plan->ops[i - 1].dst == o.src && !(plan->ops[i - 1].flags & SCOUTER_F_TF32_1PASS) && plan->ops[i - 1].w2) {
                const scouter_op_t& c = plan->ops[i - 1];
                const Buf& cb = plan->bufs[c.src];
                ConvArgs a{nullptr, c.w, c.b, nullptr, nullptr, cb.B, cb.H, cb.W, cb.C, sb.H, sb.W, sb.C, c.kh, c.kw, c.stride, c.pad,
                           c.groups, (c.flags & SCOUTER_F_RELU) ? 1 : 0, 0, 1, c.w2};
                slots = halo_gap_slots(a);
            }
            if (slots > 0) {
                plan->gap_slots[i - 1] = plan->gap_slots[i] = slots;
                scratch = std::max(scratch, (size_t)sb.B * slots * 2 * o.cout * sizeof(float));
            } else {
                scratch = std::max(scratch, (size_t)sb.B * splat_gap_splits(sb.B, sb.H * sb.W) * 2 * o.cout * sizeof(float));
                ++launches;  // partial + finish
            }
        }
    }
    plan->scratch_off = high;
    plan->arena_bytes = high + align_up(scratch, 1024);
    plan->launches = launches;
    plan->umma.assign(n_ops, UmmaConvPlan());
    plan->bound = true;
    return 0;
}

extern "C" size_t scouter_plan_arena_bytes(const scouter_plan_t* plan) { return plan && plan->bound ? plan->arena_bytes : 0; }

extern "C" int scouter_plan_buffer_shape(const scouter_plan_t* plan, int buffer, int32_t shape[4]) {
    SC_CHECK_ARG(plan && plan->bound, SCOUTER_E_STATE, "plan_buffer_shape: plan is not bound");
    SC_CHECK_ARG(buffer >= 0 && buffer < (int)plan->bufs.size() && plan->bufs[buffer].defined, SCOUTER_E_INVALID,
                 "plan_buffer_shape: buffer %d undefined", buffer);
    const Buf& b = plan->bufs[buffer];
    shape[0] = b.B; shape[1] = b.H; shape[2] = b.W; shape[3] = b.C;
    return 0;
}

extern "C" size_t scouter_plan_buffer_offset(const scouter_plan_t* plan, int buffer) {
    if (!plan || !plan->bound || buffer <= 0 || buffer >= (int)plan->bufs.size()) return (size_t)-1;
    return plan->bufs[buffer].offset;
}

extern "C" int scouter_plan_launch_count(const scouter_plan_t* plan) { return plan && plan->bound ? plan->launches : 0; }

extern "C" int scouter_plan_run(scouter_plan_t* plan, const float* input_nchw, void* arena, size_t arena_bytes,
                                scouter_stream_t stream) {
    SC_CHECK_ARG(plan && plan->bound, SCOUTER_E_STATE, "plan_run: plan is not bound");
    SC_CHECK_ARG(input_nchw && arena, SCOUTER_E_INVALID, "plan_run: NULL input / arena");
    SC_CHECK_ARG(arena_bytes >= plan->arena_bytes, SCOUTER_E_INVALID, "plan_run: arena of %zu bytes, need %zu", arena_bytes, plan->arena_bytes);
    SC_CHECK_ARG(((uintptr_t)arena & 1023) == 0, SCOUTER_E_INVALID, "plan_run: arena is not 1024-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    char* base = (char*)arena;
    auto ptr = [&](int id) -> float* { return id == 0 ? const_cast<float*>(input_nchw) : (float*)(base + plan->bufs[id].offset); };
    for (size_t i = 0; i < plan->ops.size(); ++i) {
        const scouter_op_t& o = plan->ops[i];
        const Buf& sb = plan->bufs[o.src];
        const Buf& db = plan->bufs[o.dst];
        int rc = 0;
        // per-op precision: SCOUTER_F_TF32_1PASS turns one op of a compensated plan into its single-pass form
        const bool fast = plan->math == SCOUTER_MATH_TC_FAST || (plan->math == SCOUTER_MATH_TC && (o.flags & SCOUTER_F_TF32_1PASS));
        const int rnd = fast ? 1 : 0;                                       // tf32-representable activations for the 1-pass tcgen05 convs
        const int split = (plan->math == SCOUTER_MATH_TC && !fast) ? 1 : 0; // error-compensated 3xTF32
        switch (o.kind) {
            case SCOUTER_OP_STEM_CONV: {
                SC_CHECK_ARG(o.kh == o.kw && o.groups == 1, SCOUTER_E_UNSUPPORTED, "stem conv: square, ungrouped kernels only");
                StemArgs a{ptr(o.src), o.w, o.b, ptr(o.dst), sb.B, sb.H, sb.W, sb.C, db.H, db.W, db.C, o.kh, o.stride, o.pad,
                           (o.flags & SCOUTER_F_RELU) ? 1 : 0, rnd,
                           plan->host_w[i].empty() ? nullptr : plan->host_w[i].data(),
                           plan->host_b[i].empty() ? nullptr : plan->host_b[i].data(), split};
                rc = launch_stem_conv(a, s);
                break;
            }
            case SCOUTER_OP_CONV: {
                ConvArgs a{ptr(o.src), o.w, o.b, (o.flags & SCOUTER_F_RESIDUAL) ? ptr(o.src2) : nullptr, ptr(o.dst),
                           sb.B, sb.H, sb.W, sb.C, db.H, db.W, db.C, o.kh, o.kw, o.stride, o.pad, o.groups,
                           (o.flags & SCOUTER_F_RELU) ? 1 : 0, (rnd && db.H * db.W > 1) ? 1 : 0, split, split ? o.w2 : nullptr};
                if (plan->gap_slots[i] > 0) {   // this conv's epilogue also writes the GAP partial sums of the next op
                    a.gap_part = (float*)(base + plan->scratch_off);
                    a.gap_slots = plan->gap_slots[i];
                }
                if (plan->math != SCOUTER_MATH_FP32 && tc_conv_supported(a)) rc = launch_conv_tc(a, plan->umma[i], s);
                else rc = launch_conv_simt(a, s);
                break;
            }
            case SCOUTER_OP_MAXPOOL:
                rc = launch_maxpool(ptr(o.src), ptr(o.dst), sb.B, sb.H, sb.W, sb.C, db.H, db.W, o.kh, o.stride, o.pad, s);
                break;
            case SCOUTER_OP_AVGPOOL:
                rc = launch_avgpool(ptr(o.src), ptr(o.dst), sb.B, sb.H, sb.W, sb.C, db.H, db.W, o.kh, o.stride, o.pad,
                                    (o.flags & SCOUTER_F_COUNT_INCLUDE_PAD) ? 1 : 0, rnd, s);
                break;
            case SCOUTER_OP_SPLAT_GAP:
                if (plan->gap_slots[i] > 0)
                    rc = launch_splat_gap_finish((float*)(base + plan->scratch_off), ptr(o.dst), sb.B, sb.H * sb.W, o.cout, plan->gap_slots[i], s);
                else
                    rc = launch_splat_gap(ptr(o.src), (float*)(base + plan->scratch_off), ptr(o.dst), sb.B, sb.H * sb.W, o.cout, s);
                break;
            case SCOUTER_OP_SPLAT_APPLY:
                rc = launch_splat_apply(ptr(o.src), ptr(o.src2), ptr(o.dst), sb.B, sb.H, sb.W, o.cout, db.H, db.W,
                                        (o.flags & SCOUTER_F_AVD_POOL) ? 1 : 0, rnd, s);
                break;
            case SCOUTER_OP_GAP:
                rc = launch_gap(ptr(o.src), ptr(o.dst), sb.B, sb.H * sb.W, sb.C, s);
                break;
            case SCOUTER_OP_TO_NCHW:
                rc = launch_nhwc_to_nchw(ptr(o.src), ptr(o.dst), sb.B, sb.H * sb.W, sb.C, s);
                break;
        }
        if (rc) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Fused head
// ------------------------------------------------------------------------------------------------
static int validate_head(const scouter_xslot_desc_t* desc, const scouter_head_io_t* io) {
    if (int e = validate_xslot_desc(desc)) return e;
    SC_CHECK_ARG(io, SCOUTER_E_INVALID, "head: io is NULL");
    SC_CHECK_ARG(io->batch > 0 && io->h > 0 && io->w > 0 && io->channel > 0, SCOUTER_E_INVALID,
                 "head: batch=%d h=%d w=%d channel=%d", io->batch, io->h, io->w, io->channel);
    SC_CHECK_ARG(io->channel % 16 == 0, SCOUTER_E_UNSUPPORTED, "head: channel=%d is not a multiple of 16", io->channel);
    SC_CHECK_ARG(io->layout == SCOUTER_LAYOUT_NHWC || io->layout == SCOUTER_LAYOUT_NCHW, SCOUTER_E_INVALID, "head: layout=%d", io->layout);
    SC_CHECK_ARG(io->math >= SCOUTER_MATH_FP32 && io->math <= SCOUTER_MATH_TC_FAST, SCOUTER_E_INVALID, "head: math=%d", io->math);
    return 0;
}

// workspace = [4 split-K partial slabs][NHWC copy of NCHW features][bf16 split of conv1x1.weight (fused kernel)]
static size_t head_ws_base(const scouter_head_io_t* io) {
    size_t n = (size_t)io->h * io->w;
    size_t bytes = align_up((size_t)4 * io->batch * n * XD * sizeof(float), 1024);
    if (io->layout == SCOUTER_LAYOUT_NCHW) bytes += align_up((size_t)io->batch * n * io->channel * sizeof(float), 1024);
    return bytes;
}

extern "C" size_t scouter_head_workspace_bytes(const scouter_xslot_desc_t* desc, const scouter_head_io_t* io) {
    if (validate_head(desc, io)) return 0;
    return head_ws_base(io) + head_fused_workspace_bytes(io->channel);
}

extern "C" int scouter_head_launch_count(const scouter_xslot_desc_t* desc, const scouter_head_io_t* io) {
    if (validate_head(desc, io)) return 0;
    const int n = io->h * io->w;
    const int copy = io->layout == SCOUTER_LAYOUT_NCHW ? 1 : 0;
    if (io->math != SCOUTER_MATH_FP32 && head_fused_supported(desc, io->batch, n, io->channel))
        return copy + 1 + (io->conv_w_split ? 0 : 1);
    return copy + 2;
}

extern "C" int scouter_head_forward(const scouter_xslot_desc_t* desc, const void* packed, const scouter_head_io_t* io,
                                    void* workspace, size_t workspace_bytes, scouter_stream_t stream) {
    if (int e = validate_head(desc, io)) return e;
    SC_CHECK_ARG(packed && io->feat && io->conv_w && io->conv_b && io->pe && io->logits, SCOUTER_E_INVALID, "head: NULL pointer argument");
    const size_t need = scouter_head_workspace_bytes(desc, io);
    SC_CHECK_ARG(workspace && workspace_bytes >= need, SCOUTER_E_INVALID, "head: workspace of %zu bytes, need %zu", workspace_bytes, need);
    SC_CHECK_ARG(((uintptr_t)workspace & 1023) == 0, SCOUTER_E_INVALID, "head: workspace is not 1024-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = io->h * io->w;
    float* xbuf = (float*)workspace;
    const float* feat = io->feat;
    if (io->layout == SCOUTER_LAYOUT_NCHW) {
        float* t = (float*)((char*)workspace + align_up((size_t)4 * io->batch * n * XD * sizeof(float), 1024));
        if (int e = launch_nchw_to_nhwc(io->feat, t, io->batch, io->channel, n, s)) return e;
        feat = t;
    }
    // one kernel for the whole head when a unit of images fits a 128-row UMMA tile (head_fused.cu)
    if (io->math != SCOUTER_MATH_FP32 && head_fused_supported(desc, io->batch, n, io->channel))
        return head_fused_launch(desc, packed, io, feat, (char*)workspace + head_ws_base(io), s);
    const int M = io->batch * n;
    const bool fast = xslot_fast_supported(desc, n);
    // conv1x1 + bias + ReLU (slot_model.py:108-109).  On the tensor cores the projection always runs error-compensated
    // (it is HBM-bound: the extra MMAs are free).
    ConvArgs c{feat, io->conv_w, io->conv_b, nullptr, nullptr, io->batch, io->h, io->w, io->channel, io->h, io->w, XD,
               1, 1, 1, 0, 1, 1, 0, 1, nullptr, 1};
    const bool tc = io->math != SCOUTER_MATH_FP32 && umma_conv_supported(c);
    const int kblocks = io->channel / 32;
    // split-K so that the long serial K loop (K = ch) spreads over all SMs; the partial sums are finished (fixed order,
    // + bias, ReLU) by the loop kernel while it loads its tokens
    const int ksplit = (tc && fast && kblocks % 4 == 0 && kblocks >= 16) ? 4 : 1;
    XSlotFastIO f;
    f.batch = io->batch; f.n = n; f.pe = io->pe; f.x_out = nullptr;
    f.logits = io->logits; f.attn = io->attn; f.attn_sum = io->attn_sum;
    f.x = nullptr; f.xpart = nullptr; f.conv_bias = nullptr; f.split_stride = 0; f.nsplit = 0;
    int rc;
    if (ksplit > 1) {
        c.out = xbuf;            // ksplit slabs of (M, 64)
        c.ksplit = ksplit;
        UmmaConvPlan tmp;
        if ((rc = launch_conv_umma(c, tmp, s))) return rc;
        f.xpart = xbuf; f.conv_bias = io->conv_b; f.split_stride = (long long)M * XD; f.nsplit = ksplit;
        f.x_out = io->x_out;
        return xslot_fast_launch(desc, packed, f, s);
    }
    float* x = io->x_out ? io->x_out : xbuf;
    c.out = x;
    if (tc) {
        UmmaConvPlan tmp;
        rc = launch_conv_umma(c, tmp, s);
    } else {
        rc = launch_conv_simt(c, s);
    }
    if (rc) return rc;
    if (fast) {
        f.x = x;
        return xslot_fast_launch(desc, packed, f, s);
    }
    scouter_xslot_io_t xi;
    memset(&xi, 0, sizeof(xi));
    xi.batch = io->batch; xi.n = n;
    xi.x = x; xi.x_sb = (int64_t)n * XD; xi.x_sn = XD; xi.x_sd = 1;
    xi.pe = io->pe;
    xi.logits = io->logits; xi.attn = io->attn; xi.attn_sum = io->attn_sum;
    return xslot_loop_launch(desc, packed, &xi, s);
}

// ------------------------------------------------------------------------------------------------
// a1 from host buffers (the end-to-end call)
// ------------------------------------------------------------------------------------------------
extern "C" int scouter_forward_host(const scouter_forward_host_args_t* a) {
    SC_CHECK_ARG(a && a->plan && a->desc && a->packed, SCOUTER_E_INVALID, "forward_host: NULL plan/desc/packed");
    SC_CHECK_ARG(a->input_host && a->input_dev && a->input_bytes > 0, SCOUTER_E_INVALID, "forward_host: NULL input");
    SC_CHECK_ARG(a->log_probs_dev && a->log_probs_host, SCOUTER_E_INVALID, "forward_host: NULL log_probs");
    SC_CHECK_ARG(a->plan->bound, SCOUTER_E_STATE, "forward_host: plan is not bound");
    SC_CHECK_ARG(a->feat_buffer > 0 && a->feat_buffer < (int)a->plan->bufs.size() && a->plan->bufs[a->feat_buffer].defined,
                 SCOUTER_E_INVALID, "forward_host: feat_buffer=%d undefined", a->feat_buffer);
    cudaStream_t s = (cudaStream_t)a->stream;
    SC_CUDA(cudaMemcpyAsync(a->input_dev, a->input_host, a->input_bytes, cudaMemcpyHostToDevice, s));
    if (int e = scouter_plan_run(a->plan, a->input_dev, a->arena, a->arena_bytes, a->stream)) return e;
    scouter_head_io_t h = a->head;
    h.feat = (const float*)((char*)a->arena + a->plan->bufs[a->feat_buffer].offset);
    h.layout = SCOUTER_LAYOUT_NHWC;
    if (int e = scouter_head_forward(a->desc, a->packed, &h, a->head_workspace, a->head_workspace_bytes, a->stream)) return e;
    const int S = a->desc->num_classes * a->desc->slots_per_class;
    if (int e = scouter_head_finalize(h.logits, h.attn_sum, a->target_dev, h.batch, a->desc->num_classes, S, h.h * h.w,
                                      a->desc->power, a->lambda_value, a->log_probs_dev,
                                      h.attn_sum ? a->losses_dev : nullptr, a->stream))
        return e;
    SC_CUDA(cudaMemcpyAsync(a->log_probs_host, a->log_probs_dev, (size_t)h.batch * a->desc->num_classes * sizeof(float),
                            cudaMemcpyDeviceToHost, s));
    if (a->losses_host && a->losses_dev && h.attn_sum)
        SC_CUDA(cudaMemcpyAsync(a->losses_host, a->losses_dev, 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    SC_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// single-op test entries
// ------------------------------------------------------------------------------------------------
static int conv_args_from_op(const scouter_op_t* o, const float* in, const float* res, float* out, int B, int H, int W, int math,
                             ConvArgs* a) {
    SC_CHECK_ARG(o && o->kind == SCOUTER_OP_CONV && B > 0 && H > 0 && W > 0, SCOUTER_E_INVALID, "conv_forward: bad arguments");
    SC_CHECK_ARG(o->stride >= 1 && o->kh >= 1 && o->kw >= 1 && o->groups >= 1, SCOUTER_E_INVALID, "conv_forward: geometry");
    int Ho = (H + 2 * o->pad - o->kh) / o->stride + 1, Wo = (W + 2 * o->pad - o->kw) / o->stride + 1;
    SC_CHECK_ARG(Ho > 0 && Wo > 0, SCOUTER_E_INVALID, "conv_forward: empty output");
    *a = ConvArgs{in, o->w, o->b, (o->flags & SCOUTER_F_RESIDUAL) ? res : nullptr, out, B, H, W, o->cin, Ho, Wo, o->cout,
                  o->kh, o->kw, o->stride, o->pad, o->groups, (o->flags & SCOUTER_F_RELU) ? 1 : 0,
                  math == SCOUTER_MATH_TC_FAST ? 1 : 0, math == SCOUTER_MATH_TC ? 1 : 0,
                  math == SCOUTER_MATH_TC ? o->w2 : nullptr};
    return 0;
}

extern "C" int scouter_conv_path(const scouter_op_t* op, int batch, int h, int w, int math) {
    ConvArgs a;
    if (int e = conv_args_from_op(op, nullptr, nullptr, nullptr, batch, h, w, math, &a)) return e;
    if (math == SCOUTER_MATH_FP32) return 0;
    return halo_conv_supported(a) ? 2 : (umma_conv_supported(a) ? 1 : 0);
}

extern "C" int scouter_conv_forward(const scouter_op_t* op, const float* in, const float* res, float* out, int batch, int h, int w,
                                    int math, scouter_stream_t stream) {
    ConvArgs a;
    if (int e = conv_args_from_op(op, in, res, out, batch, h, w, math, &a)) return e;
    SC_CHECK_ARG(in && out && op->w, SCOUTER_E_INVALID, "conv_forward: NULL pointer");
    if (math != SCOUTER_MATH_FP32 && tc_conv_supported(a)) {
        UmmaConvPlan tmp;
        return launch_conv_tc(a, tmp, (cudaStream_t)stream);
    }
    return launch_conv_simt(a, (cudaStream_t)stream);
}

// The network's first conv outside a plan: NCHW input (1..4 channels) -> NHWC output, CUDA-core kernel (the train-mode
// forward of scouter_b200/train.py; in train mode BatchNorm is a separate op, so no bias / ReLU is fused here).
extern "C" int scouter_stem_conv_forward(const scouter_op_t* op, const float* in_nchw, float* out, int batch, int h, int w,
                                         scouter_stream_t stream) {
    SC_CHECK_ARG(op && in_nchw && out && op->w && batch > 0 && h > 0 && w > 0, SCOUTER_E_INVALID, "stem_conv_forward: bad arguments");
    SC_CHECK_ARG(op->kh == op->kw && op->groups == 1 && op->stride >= 1, SCOUTER_E_UNSUPPORTED, "stem_conv_forward: square, ungrouped kernels only");
    const int Ho = (h + 2 * op->pad - op->kh) / op->stride + 1, Wo = (w + 2 * op->pad - op->kw) / op->stride + 1;
    SC_CHECK_ARG(Ho > 0 && Wo > 0, SCOUTER_E_INVALID, "stem_conv_forward: empty output");
    StemArgs a{in_nchw, op->w, op->b, out, batch, h, w, op->cin, Ho, Wo, op->cout, op->kh, op->stride, op->pad,
               (op->flags & SCOUTER_F_RELU) ? 1 : 0, 0, nullptr, nullptr, 0};
    return launch_stem_conv(a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// f2: input pipeline boundary
// ------------------------------------------------------------------------------------------------
extern "C" int scouter_preprocess_u8(const uint8_t* img_nhwc, int batch, int h, int w, int c, const double* mean_host,
                                     const double* std_host, float* out_nchw, scouter_stream_t stream) {
    SC_CHECK_ARG(img_nhwc && out_nchw && mean_host && std_host && batch > 0 && h > 0 && w > 0, SCOUTER_E_INVALID,
                 "preprocess_u8: bad arguments");
    for (int i = 0; i < c && i < 4; ++i) SC_CHECK_ARG(std_host[i] != 0.0, SCOUTER_E_INVALID, "preprocess_u8: std[%d] is 0", i);
    return launch_preprocess_u8(img_nhwc, batch, h, w, c, mean_host, std_host, out_nchw, (cudaStream_t)stream);
}
