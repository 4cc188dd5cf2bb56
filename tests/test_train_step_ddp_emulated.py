"""CPU, 2 gloo ranks: ONE data-parallel training step of the row-f1 groundwork end to end -- per-rank backward through the
op program and the emulated draft kernels, gradients written into ``GradientBuckets`` and all-reduced bucket by bucket
as they become ready (the step's only collective, train.py:140), then the emulated AdamW step over the flat buffers --
against the reference semantics: mean of the two ranks' oracle gradients and ``torch.optim.AdamW(params, lr)``."""
import ctypes as C
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ARGS = dict(model="resnet18", dataset="MNIST", channel=512, num_classes=10, slots_per_class=1, power=1, to_k_layer=1,
            loss_status=1, lambda_value=1.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    torch.set_num_threads(2)
    import torch.distributed as dist
    import draft_emu as E
    import scouter_b200 as sb
    from oracle.train import train_step
    from scouter_b200 import dist as sd_
    from scouter_b200.synth import fill_state_dict, make_args, synth_images, synth_labels
    from test_adamw_draft import Args as AdamArgs
    from test_train_program import program_gradients
    sd_.init_from_env(backend="gloo")
    m = sb.SlotModel(make_args(**ARGS))
    sd = fill_state_dict(m.state_dict(), seed=0)
    batches = [(synth_images(3, 1, 64, 64, seed=40 + r), synth_labels(3, 10, seed=40 + r)) for r in range(world)]
    x, tgt = batches[rank]                                              # every rank trains on its own images

    # ---- backward on this rank, gradients straight into the buckets, reduced as they become ready -----------------------
    grads, _ = program_gradients(m, ARGS, sd, x, tgt)
    names = [(n, p.shape) for n, p in m.named_parameters()]
    gb = sd_.GradientBuckets(names, "cpu", bucket_bytes=8 << 20)
    gb.zero_()
    for n, _ in reversed(names):                                        # back to front, like the backward produces them
        if n in grads:                                                  # slot.to_q.* is never reached: stays zero
            gb.grad(n).copy_(grads[n].reshape(gb.grad(n).shape))
            gb.mark_ready(n)
    gb.finish()

    # ---- reference: the mean over ranks of the oracle's gradients (what DDP hands to the optimizer) -----------------------
    kw = dict(num_classes=10, slots_per_class=1, loss_status=1, power=1, lambda_value=1.0)
    refs = [train_step("resnet18", sd, bx, bt, **kw)["grads"] for bx, bt in batches]
    scale = max(float(g.abs().max()) for g in refs[0].values() if g is not None)
    for n, _ in names:
        if refs[0][n] is None:
            assert float(gb.grad(n).abs().max()) == 0.0
            continue
        mean = sum(r[n] for r in refs) / world
        den = max(float(mean.abs().max()), 1e-4 * scale)
        assert float((gb.grad(n) - mean).abs().max()) / den < 2e-3, n

    # ---- optimizer: emulated AdamW over flat parameter buffers laid out like the gradient buckets ------------------------
    pb = sd_.GradientBuckets(names, "cpu", bucket_bytes=8 << 20)        # same layout, holds the parameters
    params = {n: torch.nn.Parameter(sd[n].clone()) for n, _ in names}
    for n, _ in names:
        pb.grad(n).copy_(sd[n])
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4)             # train.py:146
    for n, _ in names:                                                  # identical gradients on both sides; parameters the
        params[n].grad = gb.grad(n).clone() if refs[0][n] is not None else None   # backward never reached keep grad = None
    opt.step()
    lib = E.lib()
    d = opt.defaults
    b1, b2 = d["betas"]
    _f = C.POINTER(C.c_float)
    skipped = 0
    for b, (flat_p, flat_g) in enumerate(zip(pb.flat, gb.flat)):
        mom_b, var_b = np.zeros(flat_p.numel(), np.float32), np.zeros(flat_p.numel(), np.float32)
        skipped += flat_p.numel() - sum(k for _, k in gb.ready_ranges(b))
        for off, k in gb.ready_ranges(b):                              # only what received a gradient (train.py:145, AdamW: grad is None)
            pn, gn = flat_p.numpy()[off:off + k], flat_g.numpy()[off:off + k]
            mom, var = mom_b[off:off + k], var_b[off:off + k]
            a = AdamArgs(decay=1 - 1e-4 * d["weight_decay"], one_minus_beta1=1 - b1, beta2=b2, one_minus_beta2=1 - b2, eps=d["eps"],
                         step_size=1e-4 / (1 - b1), bias_correction2_sqrt=math.sqrt(1 - b2), n=k, p=pn.ctypes.data_as(_f),
                         g=gn.ctypes.data_as(_f), m=mom.ctypes.data_as(_f), v=var.ctypes.data_as(_f))
            lib.adamw_host(C.byref(a))
    assert skipped == sum(sd[n].numel() for n, _ in names if refs[0][n] is None) > 0     # slot.to_q.*: no decay, no Adam state
    for n, _ in names:
        if refs[0][n] is None:
            assert torch.equal(pb.grad(n), sd[n]) and torch.equal(params[n].detach(), sd[n])
    worst = max(float((pb.grad(n) - params[n].detach()).abs().max()) for n, _ in names)
    assert worst <= 2e-7, worst
    moved = max(float((pb.grad(n) - sd[n]).abs().max()) for n, _ in names)
    assert 0.5e-4 < moved < 2.5e-4                                      # first Adam step moves by about lr
    # both ranks hold the same parameters after the step
    digest = [None] * world
    dist.all_gather_object(digest, float(sum(f.double().sum() for f in pb.flat)))
    assert digest[0] == digest[1]
    out[rank] = True
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_rank_training_step_emulated():
    import shutil
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(800)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(out) == {0: True, 1: True}
