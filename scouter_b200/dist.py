"""Data-parallel plumbing for the forward path (reference: train.py:139-160, tools/prepare_things.py:9-31).

SlotModel.forward shards over the batch with no collective: every rank holds a full replica and runs its own images.
The only cross-rank operations are the rendezvous, a barrier, and reductions of *measurements* (max time over ranks,
sum of images).  `torch.distributed` with NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None, device: torch.device | None = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's env; initialises the process group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend in (None, "nccl") and device is not None and device.type == "cuda":
            kw["device_id"] = device
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of `total` independent images for `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_images_per_second(images_this_rank: int, ms_this_rank: float, device: torch.device | str = "cpu") -> float:
    """Whole-job throughput: all images of all ranks / the slowest rank's time."""
    return sum_over_ranks(images_this_rank, device) / (max_over_ranks(ms_this_rank, device) / 1e3)


class GradientBuckets:
    """Flat fp32 gradient storage, bucketed for an all-reduce that overlaps the backward -- the one exchange step of the
    training path (the job ``DistributedDataParallel`` does in the reference, train.py:140, 60.8 MB of fp32
    gradients per step for resnest26d).  Row f1 groundwork: the CUDA backward will accumulate straight into these
    views through raw pointers (there is no autograd graph for DDP's hooks to attach to).

    Parameters are packed in *reverse* registration order -- gradients become ready back to front -- into buckets of
    about ``bucket_bytes``; ``mark_ready(name)`` launches the asynchronous all-reduce of a bucket as soon as its last
    gradient is in, ``finish()`` launches what is left, waits and turns the sums into means.  With NCCL each all-reduce
    is stream-ordered after the kernels already enqueued on the current stream; with gloo (CPU tests) it is a host
    call.  Unused parameters (``slot.to_q.*``; ``find_unused_parameters=True`` in the reference) and frozen ones never get
    ``mark_ready``: their gradient stays zero for the exchange, and ``ready_ranges`` leaves them OUT of the optimizer pass --
    the reference builds ``params = [p for p in ... if p.requires_grad]`` (train.py:145) and ``torch.optim.AdamW`` skips
    ``grad is None``, so those tensors receive neither weight decay nor Adam state."""

    def __init__(self, named_shapes, device="cpu", bucket_bytes: int = 25 << 20):
        self.device = torch.device(device)
        self.names = [n for n, _ in named_shapes]
        if len(set(self.names)) != len(self.names):
            raise ValueError("GradientBuckets: duplicate parameter names")
        plan, cur, cur_bytes = [], [], 0
        for name, shape in reversed(list(named_shapes)):
            numel = 1
            for s in shape:
                numel *= int(s)
            if cur and cur_bytes + 4 * numel > bucket_bytes:
                plan.append(cur)
                cur, cur_bytes = [], 0
            cur.append((name, tuple(shape), numel))
            cur_bytes += 4 * numel
        if cur:
            plan.append(cur)
        self.flat, self.views, self.bucket_of, self._members, self._layout = [], {}, {}, [], []
        for b, members in enumerate(plan):
            flat = torch.zeros(sum(m[2] for m in members), dtype=torch.float32, device=self.device)
            off = 0
            for name, shape, numel in members:
                self.views[name] = flat[off:off + numel].view(shape)
                self.bucket_of[name] = b
                off += numel
            self.flat.append(flat)
            self._members.append({m[0] for m in members})
            self._layout.append([(m[0], m[2]) for m in members])
        self._pending = [set(m) for m in self._members]
        self._work = [None] * len(self.flat)
        self._launched = [False] * len(self.flat)
        self._written = set()

    def grad(self, name: str) -> torch.Tensor:
        """The gradient view of one parameter (aliases the bucket's flat buffer)."""
        return self.views[name]

    def zero_(self):
        """Start of a step: clear the gradients and the readiness bookkeeping."""
        for f in self.flat:
            f.zero_()
        self._pending = [set(m) for m in self._members]
        self._work = [None] * len(self.flat)
        self._launched = [False] * len(self.flat)
        self._written = set()

    def _launch(self, b: int):
        if self._launched[b]:
            return
        self._launched[b] = True
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._work[b] = dist.all_reduce(self.flat[b], op=dist.ReduceOp.SUM, async_op=True)

    def mark_ready(self, name: str):
        """The backward has finished writing this gradient; all-reduce its bucket once every member is in."""
        b = self.bucket_of[name]
        if self._launched[b]:
            raise RuntimeError(f"GradientBuckets: {name} marked ready after its bucket was reduced (missing zero_()?)")
        self._pending[b].discard(name)
        self._written.add(name)
        if not self._pending[b]:
            self._launch(b)

    def finish(self):
        """End of the backward: reduce the buckets that are still open (unused parameters), wait, average."""
        for b in range(len(self.flat)):
            self._launch(b)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        for b, w in enumerate(self._work):
            if w is not None:
                w.wait()
            if world > 1:
                self.flat[b].mul_(1.0 / world)

    def ready_ranges(self, b: int):
        """[(offset, numel)] of bucket ``b``: maximal runs of parameters that received a gradient this step (``mark_ready``).
        The optimizer pass runs over these ranges only -- parameters without a gradient (unused, frozen) are not decayed and
        get no Adam state, like ``torch.optim.AdamW`` with ``grad is None``.  Every rank marks the same set."""
        out, off = [], 0
        for name, numel in self._layout[b]:
            if name in self._written:
                if out and out[-1][0] + out[-1][1] == off:
                    out[-1] = (out[-1][0], out[-1][1] + numel)
                else:
                    out.append((off, numel))
            off += numel
        return out

    @property
    def total_bytes(self) -> int:
        return 4 * sum(f.numel() for f in self.flat)
