"""N>1 host logic on CPU: clip sharding + max-over-ranks timing reduction with a world_size-2 gloo group."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lavt_rs_b200.sharding import all_slices, clip_slice, max_over_ranks  # noqa: E402


def test_slices_partition_the_clips():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            sl = all_slices(n, world)
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [e - s for s, e in sl]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = clip_slice(13, rank, world)
    # each rank "processes" its clips: result = clip index squared; the union must equal the single-process result
    mine = torch.tensor([i * i for i in range(s, e)], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()]))
    slowest = max_over_ranks(10.0 + rank)
    q.put((rank, mine.tolist(), [int(t) for t in sizes], slowest))
    dist.destroy_process_group()


def test_world_size_two_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29511 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = out[0][1] + out[1][1]
    assert merged == [i * i for i in range(13)]
    assert out[0][2] == [7, 6] and out[0][3] == out[1][3] == 11.0


def _grad_worker(rank, world, port, q):
    """training.allreduce_gradients: bucketed average of param.grad over the ranks (the one collective of the training path)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lavt_rs_b200.training import allreduce_gradients
    g = torch.Generator().manual_seed(5)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((3, 5), (7,), (2, 2, 2), (300000,), (11,))]
    base = [torch.randn(p.shape, generator=g) for p in params]
    for p, b in zip(params, base):
        p.grad = b * (rank + 1)            # rank r holds (r + 1) * base -> the average is 1.5 * base for two ranks
    params[2].grad = None                   # a parameter without gradient (frozen / dead) is skipped, not all-reduced
    allreduce_gradients(params, bucket_mb=1)        # 1 MiB buckets: the 300 000-element tensor forces several flushes
    ok = all(torch.allclose(p.grad, 1.5 * b, rtol=1e-6, atol=1e-6) for p, b in zip(params, base) if p.grad is not None)
    q.put((rank, ok, params[2].grad is None))
    dist.destroy_process_group()


def test_gradient_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31511 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and none_kept for _, ok, none_kept in out)


def _reducer_worker(rank, world, port, q):
    """training.GradReducer: per-group asynchronous all-reduce (what segment_backward's on_ready callback drives)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lavt_rs_b200.training import GradReducer
    from lavt_rs_b200.train_engine import GradStore
    g = torch.Generator().manual_seed(6)
    groups = [[torch.nn.Parameter(torch.zeros(s)) for s in shapes] for shapes in (((4, 3), (5,)), ((7, 2, 2),), ((1,), (9, 9)))]
    base = [[torch.randn(p.shape, generator=g) for p in grp] for grp in groups]
    ok = True
    # (slotted flat allocation | loose buffers) x (launch per group under the backward | deferred to wait()) x (fp32 | bf16 on the wire)
    for slotted, overlap, compress in ((True, True, None), (True, False, None), (False, True, None), (True, True, "bf16")):
        for grp in groups:
            for p in grp:
                p.grad = None
        # interleave the slot order of the groups so that a group maps to several non-adjacent slices of the flat buffer
        store = GradStore([groups[0][0], groups[2][0], groups[1][0], groups[0][1], groups[2][1]]) if slotted else GradStore()
        red = GradReducer(overlap=overlap, compress=compress)
        cb = red.ready(store)
        for grp, bs in zip(groups, base):
            for p, b in zip(grp, bs):
                store.of(p).add_(b * (rank + 1))
            cb(grp)                              # group finished: hand over + start its all-reduce
        red.wait()
        tol = 1e-2 if compress else 1e-6
        ok = ok and all(torch.allclose(p.grad, 1.5 * b, rtol=tol, atol=tol) for grp, bs in zip(groups, base) for p, b in zip(grp, bs))
        if slotted:      # param.grad aliases the flat allocation: the all-reduce ran in place, no copy-back exists
            flat = store._flat[groups[0][0].device]
            ok = ok and all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for grp in groups for p in grp)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_overlapped_gradient_reducer_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33511 + os.getpid() % 2000
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in out)
