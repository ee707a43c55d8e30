"""Data-parallel host logic on CPU with gloo, world_size 2 (the N>1 path of bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from glfusion_b200 import dp


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = dp.shard_range(n, r, world)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
            sizes = [dp.shard_range(n, r, world)[1] - dp.shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, lr, w = dp.init_process_group("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    bn = torch.nn.BatchNorm1d(3)
    # rank-dependent gradients and buffers
    for i, p in enumerate(lin.parameters()):
        p.grad = torch.full_like(p, float(rank + 1 + i))
    bn.running_mean.fill_(float(rank + 5))
    bucket = dp.GradBucket(lin.parameters())
    bucket.allreduce_mean()
    dp.broadcast_buffers(bn, src=0)
    t = dp.max_over_ranks(10.0 * (rank + 1))
    out = {"grads": [p.grad.clone() for p in lin.parameters()], "rm": bn.running_mean.clone(), "t": t}
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_and_buffer_broadcast_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in range(world):
        g0, g1 = res[r]["grads"]
        # mean over ranks of (rank+1+i): i=0 -> 1.5, i=1 -> 2.5
        assert torch.allclose(g0, torch.full_like(g0, 1.5)) and torch.allclose(g1, torch.full_like(g1, 2.5))
        assert torch.allclose(res[r]["rm"], torch.full_like(res[r]["rm"], 5.0))     # rank 0's buffer
        assert res[r]["t"] == 20.0


def test_single_process_is_a_noop():
    lin = torch.nn.Linear(2, 2)
    for p in lin.parameters():
        p.grad = torch.ones_like(p)
    b = dp.GradBucket(lin.parameters())
    b.allreduce_mean()
    assert all(torch.equal(p.grad, torch.ones_like(p)) for p in lin.parameters())
    assert dp.max_over_ranks(3.0) == 3.0


def test_grad_bucket_zero_copy_aliasing():
    """bind()-style zero-copy mode: when .grad aliases the bucket view no pack/unpack copy is made, and a gradient
    that lives elsewhere falls back to the copying path (single process: the all-reduce itself is a no-op)."""
    import torch
    from glfusion_b200.dp import GradBucket
    ps = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))]
    b = GradBucket(ps)
    for p, v in zip(ps, b.views):
        v.copy_(torch.arange(v.numel(), dtype=torch.float32).view_as(v))
        p.grad = v.detach()
    assert b.aliased()
    b.allreduce_mean()
    assert ps[0].grad.data_ptr() == b.views[0].data_ptr()
    assert float(ps[1].grad[4]) == 4.0
    ps[1].grad = torch.full((5,), 7.0)          # not an alias any more -> copying path
    assert not b.aliased()
    b.allreduce_mean()
    assert float(b.flat[-1]) == 7.0 and float(ps[1].grad[0]) == 7.0


def test_bucket_without_cuda_keeps_the_collective_path():
    """The peer-memory all-reduce (csrc/glf_p2p.cu) exists for CUDA buckets only: a CPU bucket reports it as
    unavailable and keeps using the process-group collective (exercised by the gloo tests above)."""
    import torch
    from glfusion_b200 import dp
    params = [torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(3, 2))]
    b = dp.GradBucket(params)
    assert b.enable_p2p() is False and not b.p2p_enabled
    assert b.flat.numel() == 11


def test_p2p_entry_points_reject_bad_arguments():
    """Argument validation of the C ABI happens before any device work (no GPU needed)."""
    import ctypes as C
    import glfusion_b200
    lib = glfusion_b200.load_library()
    assert lib.glf_p2p_signal_bytes(8) >= 2 * 148 * 8 * 4
    assert lib.glf_p2p_max_floats() >= 345_000
    off = C.c_uint64(0)
    assert lib.glf_p2p_export(None, C.create_string_buffer(64), C.byref(off)) < 0
    assert b"NULL" in lib.glf_last_error()
