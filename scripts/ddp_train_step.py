"""Row f1 + row e: ONE data-parallel training step of SlotModel on N B200s, exactly as train.py:139-148 / engine.py:28-35
drive it -- DistributedDataParallel(find_unused_parameters=True) around the module, AdamW, loss.backward() -- with this
repo's CUDA path underneath (train-mode forward + backward through the C ABI; the only collective is DDP's NCCL gradient
all-reduce over NVLink, 60.8 MB of fp32 gradients for resnest26d).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/ddp_train_step.py [--batch B] [--steps K]

Checks (rank 0 prints one JSON line): after backward every rank holds the MEAN over ranks of the local gradients (computed
again without DDP and all-gathered); after optimizer.step() the parameters are bit-identical on all ranks; BatchNorm running
statistics stay per rank (the reference does not sync them).  Also reports the step time (CUDA events, max over ranks)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scouter_b200 as sb  # noqa: E402
from scouter_b200 import _lib as L  # noqa: E402
from scouter_b200.synth import fill_state_dict, make_args  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--freeze", type=int, default=0)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29511")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

args = dict(model="resnest26d", num_classes=10, slots_per_class=1, power=2, to_k_layer=3, loss_status=-1, channel=2048)
m = sb.SlotModel(make_args(**args))
m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
m = m.to(dev).train()
if a.freeze:
    m.dfs_freeze(m.backbone, a.freeze)
g = torch.Generator(device=dev).manual_seed(100 + rank)           # DistributedSampler: every rank sees its own images
x = torch.randn(a.batch, 3, a.size, a.size, device=dev, generator=g)
y = torch.randint(0, 10, (a.batch,), device=dev, generator=g)

# local gradients without DDP (reference for the all-reduce), from a copy of the BatchNorm state
state0 = {k: v.clone() for k, v in m.state_dict().items()}
out, (loss, nll, attn) = m(x, y)
loss.backward()
named = [(k, p) for k, p in m.named_parameters() if p.grad is not None]
local_g = torch.cat([p.grad.reshape(-1) for _, p in named])
mean_g = local_g.clone()
dist.all_reduce(mean_g)
mean_g /= world
m.load_state_dict(state0)
m.zero_grad(set_to_none=True)

ddp = DDP(m, device_ids=[local], find_unused_parameters=True)      # train.py:140
opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4)   # train.py:145-146
out, (loss, nll, attn) = ddp(x, y)
loss.backward()
got = torch.cat([p.grad.reshape(-1) for _, p in named])
scale = float(mean_g.abs().max())
# the backward merges conv gradients with fp32 atomics: two runs of the same step agree to rounding, not bitwise
grad_err = float((got - mean_g).abs().max()) / scale
opt.step()
flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
ref = flat.clone()
dist.broadcast(ref, 0)
same_params = bool(torch.equal(flat, ref))
rm = m.backbone.bn1.running_mean.clone()
rm0 = rm.clone()
dist.broadcast(rm0, 0)
bn_differs = bool(world == 1 or not torch.equal(rm, rm0) or rank == 0)

# step time: forward + backward (with the all-reduce) + optimizer, CUDA events, max over ranks
for _ in range(1):
    opt.zero_grad(set_to_none=True)
    ddp(x, y)[1][0].backward()
    opt.step()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    opt.zero_grad(set_to_none=True)
    ddp(x, y)[1][0].backward()
    opt.step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
flags = torch.tensor([float(same_params), float(bn_differs), grad_err], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
gmax = torch.tensor([grad_err], device=dev)
dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"what": "one DDP training step of SlotModel (resnest26d + negative xSlot) through scouter_b200, NCCL gradient all-reduce",
                      "n_gpus": world, "batch_per_gpu": a.batch, "size": a.size, "freeze_layers": a.freeze,
                      "grad_allreduce_err_vs_mean_of_local_grads": float(gmax), "params_identical_after_step": bool(flags[0] > 0),
                      "bn_running_stats_per_rank": bool(flags[1] > 0), "gradient_floats": int(local_g.numel()),
                      "ms_per_step": float(ms), "images_per_s": world * a.batch / (float(ms) / 1e3), "loss": float(loss)}))
    assert float(gmax) < 2e-3 and flags[0] > 0, "DDP gradient exchange does not reproduce the mean of the local gradients"
dist.destroy_process_group()
