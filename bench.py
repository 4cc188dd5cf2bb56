#!/usr/bin/env python
"""bench.py -- images/sec of SlotModel.forward (resnest26d, 10 slots) on N B200s (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                   # the reference algorithm on the host CPU

A "step" is one eval-mode forward of one synthetic batch: cfg 3 of BASELINE.json -- ImageNet-10
resnest26d + negative xSlot (loss_status -1, to_k_layer 3), batch 256 per GPU, 3x224x224 fp32 (the
batch north_star's roofline target is quoted on; override with --batch/--size).  One JSON line on stdout:

  value        whole-job images/s, inputs resident in HBM, whole forward replayed from a CUDA graph,
               CUDA-event timed, max over ranks (weak scaling: every rank runs its own batch of B)
  e2e          same metric through the C-ABI host entry (scouter_forward_host): pinned host batch ->
               H2D -> forward -> D2H of the log-probs, every step
  roofline     the fused xSlot head (conv1x1 + PE + to_k + 3x attention/GRU + logits): algorithmic
               bytes (B*ch*n*4 feature read + weights + logits) / its CUDA-event time, vs measured HBM GB/s
  roofline_backbone   the backbone program: algorithmic conv FLOPs / its time vs the measured tensor peak
  cpu_baseline the oracle port of the reference forward on the host cores (bounded sample)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec SlotModel forward (resnest26d, 10 slots) @1/2/4/8 B200"     # BASELINE.json's metric string, verbatim
ARGS = dict(model="resnest26d", dataset="ImageNet", channel=2048, num_classes=10, slots_per_class=1, power=2,
            to_k_layer=3, loss_status=-1, lambda_value=1.0)
# 2*MACs of the 47 backbone convs per image
BACKBONE_GFLOP = {224: 7.24, 260: 10.30}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                clk, mxv = float(f[0]), float(f[1])
            except ValueError:
                continue
            mx = mxv
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def cpu_reference_rate(batch, size, steps, warmup):
    """The reference algorithm (oracle port: oracle/backbone.py + oracle/head.py) on the host cores."""
    import scouter_b200 as sb
    from oracle import backbone as ob
    from scouter_b200.synth import fill_state_dict, make_args, synth_images
    torch.set_num_threads(os.cpu_count())
    m = sb.SlotModel(make_args(**ARGS))
    sd = fill_state_dict(m.state_dict(), seed=0)
    x = synth_images(batch, 3, size, size)
    fwd = lambda: ob.slot_model_forward("resnest26d", sd, x, num_classes=10, slots_per_class=1, loss_status=-1, power=2)
    with torch.no_grad():
        for _ in range(warmup):
            fwd()
        t0 = time.perf_counter()
        for _ in range(steps):
            fwd()
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, torch.get_num_threads()


# dram__bytes_read.sum + dram__bytes_write.sum of one head_fused_kernel launch at B=256, 224^2 (ncu --set full)
HEAD_DRAM_TRAFFIC = 104132608 + 4102656


_REAL_STDOUT = None


def emit(line: str):
    """Print the result line on the real stdout (see main)."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--math", default=os.environ.get("SCOUTER_MATH", "tc"), choices=["tc", "fp32", "tc_fast"])
    ap.add_argument("--cpu-sample", type=int, default=32, help="images per CPU-baseline step")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that write banners to file descriptor 1 (NCCL prints its version
    # there at communicator creation) go to stderr until the line is printed
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"cfg3: ImageNet-10 resnest26d + negative xSlot (loss_status -1, to_k_layer 3, 10x1 slots), "
                f"batch {a.batch}/GPU, 3x{a.size}x{a.size} fp32, eval forward")

    if a.impl == "reference":
        if rank != 0:
            return
        rate, spb, cores = cpu_reference_rate(a.cpu_sample, a.size, a.steps, min(a.warmup, 1))
        sample = f"{a.cpu_sample} images/step x {a.steps} steps of the same workload (oracle port of the reference, torch CPU ops)"
        emit(json.dumps({
            "impl": "reference", "metric": METRIC, "value": rate, "unit": "images/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": min(a.warmup, 1), "ms_per_step": spb * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "note": "host CPU only; bounded sample per step"},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch.distributed as dist
    import scouter_b200 as sb
    from scouter_b200 import _lib as L
    from scouter_b200.synth import fill_state_dict, make_args     # nothing under oracle/ is imported on this arm

    from scouter_b200 import dist as sdist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sdist.init_from_env("nccl", dev)
    L.check(L.lib().scouter_device_check(local))

    m = sb.SlotModel(make_args(**ARGS))
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
    m = m.to(dev).eval()
    m.math = {"tc": L.MATH_TC, "fp32": L.MATH_FP32, "tc_fast": L.MATH_TC_FAST}[a.math]
    m.use_cuda_graph = True
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(a.batch, 3, a.size, a.size, device=dev, generator=g)
    x_host = x.cpu().pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    with torch.no_grad():
        # ---- whole forward, inputs resident, CUDA graph ------------------------------------------
        out = m(x)                                   # builds plan, captures the graph
        st = next(iter(m._states.values()))
        step = lambda: st.graph.replay()
        for _ in range(a.warmup):
            step()
        sampler = ClockSampler(local) if rank == 0 else None
        t0 = time.time()
        ms = timed(step, a.steps)
        t1 = time.time()
        clocks = sampler.stop(t0, t1) if sampler else None
        value = world * a.batch * a.steps / (ms / 1e3)

        # ---- end to end from host memory ----------------------------------------------------------------
        # (a) one synchronous C call per step (scouter_forward_host: H2D, forward, D2H, sync);
        # (b) the streaming API: same per-step copies, but batch i+1 uploads while batch i computes
        for _ in range(2):
            m.forward_host(x_host, dev)
        ms_e2e_sync = timed(lambda: m.forward_host(x_host, dev), a.steps)
        for _ in m.forward_host_stream([x_host] * 3, dev):
            pass

        def stream_steps():
            n = 0
            for out in m.forward_host_stream([x_host] * a.steps, dev):
                n += 1
            assert n == a.steps

        barrier()
        t_w0 = time.perf_counter()
        stream_steps()
        torch.cuda.synchronize()
        ms_e2e = (time.perf_counter() - t_w0) * 1e3
        if world > 1:
            t = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t)
        e2e_value = world * a.batch * a.steps / (ms_e2e / 1e3)
        e2e_sync_value = world * a.batch * a.steps / (ms_e2e_sync / 1e3)

        # ---- per-kernel: the fused head and the backbone program, CUDA events on the launch stream ---
        import ctypes as C
        desc, packed = m._head_params(st, dev)
        lib = L.lib()
        head = lambda: L.check(lib.scouter_head_forward(C.byref(desc), packed.data_ptr(), C.byref(st.io),
                                                        st.ws.data_ptr() + st.ws_off, st.ws_bytes, L.stream_ptr()))
        bb = lambda: st.cp.run(st.static_in)
        for _ in range(3):
            head(); bb()
        # the head reads 103 MB of features, less than the 126 MB L2: every timed call is preceded by a 256 MB write so
        # that the features come from HBM as they do inside the forward (the flush is outside the events)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed_flushed(fn, steps):
            tot = 0.0
            for _ in range(steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            ms = tot / steps
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
            return ms

        ms_head = timed_flushed(head, a.steps)
        ms_bb = timed(bb, a.steps) / a.steps
        head_launches = lib.scouter_head_launch_count(C.byref(desc), C.byref(st.io))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    n = st.n
    S, Lk = 10, 3
    w_bytes = 2048 * 64 * 4 + 64 * 4 + Lk * (64 * 64 + 64) * 4 + S * 64 * 4 + 2 * (192 * 64 + 192) * 4 + 64 * n * 4
    head_bytes = a.batch * 2048 * n * 4 + a.batch * 10 * 4 + w_bytes
    head_gbs = head_bytes / (ms_head / 1e3) / 1e9
    bb_tflops = BACKBONE_GFLOP.get(a.size, 7.24 * (a.size / 224.0) ** 2) * a.batch / (ms_bb / 1e3) / 1e3
    tf32_peak = pk["bf16_sustained"] / 2
    # bounded CPU sample: ~10 s of host work at N=1 (the reported baseline), a token 3 steps on multi-GPU lines
    cpu_steps = 24 if world == 1 else 3
    cpu_rate, cpu_spb, cores = cpu_reference_rate(a.cpu_sample, a.size, cpu_steps, 1)
    launches = m.launches_per_forward(tuple(x.shape), dev)
    emit(json.dumps({
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tc": "tf32 + 2 bf16 correction products (error-compensated, fp32-class; fp32 accumulate)", "tc_fast": "tf32", "fp32": "f32"}[a.math], "data": "synthetic",
        "config": {"workload": workload, "global_batch": world * a.batch, "parallelism": f"dp{world}",
                   "math": a.math, "timing": "CUDA events, max over ranks; whole forward = one CUDA-graph replay; inputs "
                   "(154 MB batch + GBs of activations) larger than the 126 MB L2, no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": a.batch * 10 * 4,
                "api": "SlotModel.forward_host_stream (pinned host batches in, pinned log-probs out; the H2D of batch i+1 "
                       "overlaps the compute of batch i; wall-clock over all steps incl. first upload and last read-back)",
                "synchronous_call_value": e2e_sync_value,
                "synchronous_call": "scouter_forward_host: H2D + forward + D2H + sync in one C call per step"},
        "gpu_launches": launches * a.steps,
        "roofline": {"kernel": "xSlot head: scouter_head_forward (conv1x1+ReLU+PE+to_k+3x{QK^T,normalise,sigmoid,attn.V,GRU}+logits)",
                     "bound": "hbm", "achieved": head_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": head_gbs / pk["hbm"],
                     "traffic": HEAD_DRAM_TRAFFIC if (a.batch, a.size) == (256, 224) else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of head_fused_kernel, one ncu --set full capture "
                                       "(profiles/r01_ncu_head_fused.md)",
                     "ms": ms_head, "algorithmic_bytes": head_bytes, "peak_source": pk["src"] + " copy bandwidth",
                     "launches": head_launches, "timing": "CUDA events around each call, 256 MB L2 flush before each call"},
        "roofline_backbone": {"kernel": "backbone op program (47 convs + pools + split attention)", "bound": "tensor",
                              "achieved": bb_tflops, "peak": tf32_peak, "unit": "TFLOP/s", "frac": bb_tflops / tf32_peak,
                              "ms": ms_bb, "peak_source": pk["src"] + " bf16 sustained / 2 (tf32 assumed half of bf16)"},
        "cpu_baseline": {"value": cpu_rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{a.cpu_sample} images/step x {cpu_steps} steps after 1 warm-up (oracle port of the reference "
                                   "forward, torch CPU ops, all host threads)"},
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
