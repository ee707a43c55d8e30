"""Worker for tests/test_gpu_p2p.py (launched under torchrun, one process per GPU): checks glf_p2p_allreduce against
an NCCL all-reduce of the same data, eagerly and replayed from a CUDA graph, and the bitwise agreement of all ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glfusion_b200 import dp  # noqa: E402


def main():
    rank, local_rank, world = dp.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n, device=dev)) for n in (128 * 264, 256 * 128, 256, 7, 345_001)]
    bucket = dp.GradBucket(params)
    assert bucket.enable_p2p(), "IPC exchange failed"
    views = bucket.views
    for it in range(6):
        g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
        local = [torch.randn(v.shape, generator=g, device=dev) * (1 + rank) for v in views]
        ref = [t.clone() for t in local]
        for t in ref:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        for p, v, t in zip(params, views, local):
            v.copy_(t)
            p.grad = v                      # aliased: the zero-copy path
        bucket.allreduce_mean()
        torch.cuda.synchronize()
        for v, t in zip(views, ref):
            assert torch.allclose(v, t / world, rtol=1e-6, atol=1e-6), (it, float((v - t / world).abs().max()))
        # every rank holds the same bits
        chk = bucket.flat.clone()
        dist.broadcast(chk, src=0)
        assert torch.equal(chk, bucket.flat)
    # replay from a CUDA graph, several times in a row (self-resetting signals)
    src = torch.randn(bucket.numel, device=dev) + rank
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        bucket.flat.copy_(src)
        bucket.allreduce_mean()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        bucket.flat.copy_(src)
        bucket.allreduce_mean()
    ref = src.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    for _ in range(20):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.allclose(bucket.flat, ref / world, rtol=1e-6, atol=1e-6)
    dist.barrier()
    if rank == 0:
        print("P2P_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
