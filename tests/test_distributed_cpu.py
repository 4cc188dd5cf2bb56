"""CPU, world_size 2, gloo: the data-parallel plumbing of the forward path (no collective touches the data path)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from scouter_b200 import dist as sd
    r, w, l = sd.init_from_env(backend="gloo")
    assert (r, w, l) == (rank, world, rank)
    # 1. shards are disjoint, ordered and cover the batch (uneven split included)
    b, e = sd.shard_range(257, r, w)
    sizes = [None] * w
    dist.all_gather_object(sizes, (b, e))
    assert sizes[0][0] == 0 and sizes[-1][1] == 257 and all(sizes[i][1] == sizes[i + 1][0] for i in range(w - 1))
    assert max(x[1] - x[0] for x in sizes) - min(x[1] - x[0] for x in sizes) <= 1
    # 2. each rank "processes" its shard independently (a replica: same weights, own images); results only meet in a gather
    torch.manual_seed(0)
    weight = torch.randn(16, 4)
    images = torch.arange(257 * 16, dtype=torch.float32).reshape(257, 16)
    mine = images[b:e] @ weight
    gathered = [None] * w
    dist.all_gather_object(gathered, mine)
    assert torch.equal(torch.cat(gathered), images @ weight)
    # 3. timing reductions: max over ranks, whole-job throughput
    sd.barrier()
    ms = 10.0 * (rank + 1)
    assert sd.max_over_ranks(ms) == 10.0 * world
    ips = sd.aggregate_images_per_second(e - b, ms)
    assert abs(ips - 257 / (10.0 * world / 1e3)) < 1e-6
    out[rank] = True
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_sharding_and_reductions():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(100)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(out) == {0: True, 1: True}


def test_shard_range_properties():
    from scouter_b200.dist import shard_range
    for total in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _bucket_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import scouter_b200 as sb
    from scouter_b200 import dist as sd
    from scouter_b200.synth import make_args
    sd.init_from_env(backend="gloo")
    m = sb.SlotModel(make_args(model="resnet18", dataset="MNIST", channel=512, to_k_layer=1, power=1))
    shapes = [(n, p.shape) for n, p in m.named_parameters()]
    gb = sd.GradientBuckets(shapes, "cpu", bucket_bytes=4 << 20)
    assert gb.total_bytes == 4 * sum(p.numel() for p in m.parameters()) == 44937728      # 44.9 MB
    assert len(gb.flat) > 5 and all(f.numel() * 4 <= (4 << 20) or len(ms) == 1 for f, ms in zip(gb.flat, gb._members))
    assert gb.bucket_of[shapes[-1][0]] == 0                      # last registered parameter -> first bucket (ready first)
    for step in range(2):                                        # two steps: zero_() resets the bookkeeping
        gb.zero_()
        order = [n for n, _ in shapes if not n.startswith("slot.to_q.")][::-1]      # back to front, to_q never ready
        g = torch.Generator().manual_seed(100 * step + rank)
        mine = {}
        for n in order:
            v = gb.grad(n)
            v.copy_(torch.randn(v.shape, generator=g))
            mine[n] = v.clone()
            gb.mark_ready(n)
        gb.finish()
        for n in order[:: max(1, len(order) // 12)]:             # every rank ends with the mean over ranks
            both = [None] * world
            dist.all_gather_object(both, mine[n])
            assert torch.allclose(gb.grad(n), sum(both) / world, rtol=0, atol=1e-6), n
        assert float(gb.grad("slot.to_q.0.weight").abs().max()) == 0.0               # unused parameter stays zero
    out[rank] = True
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_gradient_buckets():
    """f1 groundwork: bucketed, readiness-driven gradient all-reduce (the training path's only collective) on 2 gloo ranks."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(150)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(out) == {0: True, 1: True}


def test_gradient_buckets_single_process():
    from scouter_b200.dist import GradientBuckets
    gb = GradientBuckets([("a", (3, 4)), ("b", (5,)), ("c", (2, 2, 2))], bucket_bytes=40)
    assert [sorted(m) for m in gb._members] == [["c"], ["b"], ["a"]]      # reverse order; 32 B, 20 B, 48 B against a 40 B cap
    gb.grad("a").fill_(2.0)
    assert gb.flat[gb.bucket_of["a"]].sum() == 24.0              # the view aliases the flat buffer
    for n in ("c", "b", "a"):
        gb.mark_ready(n)
    gb.finish()                                                  # no process group: a no-op, values unchanged
    assert float(gb.grad("a").mean()) == 2.0
    with pytest.raises(RuntimeError):
        gb.mark_ready("a")                                       # bucket already reduced this step
    gb.zero_()
    assert float(gb.grad("a").abs().max()) == 0.0
    with pytest.raises(ValueError):
        GradientBuckets([("a", (1,)), ("a", (2,))])
