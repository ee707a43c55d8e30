"""Worker for tests/test_gpu_p2p.py (launched under torchrun, one process per GPU): checks glf_p2p_allreduce against
an NCCL all-reduce of the same data, eagerly and replayed from a CUDA graph, and the bitwise agreement of all ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glfusion_b200 import dp  # noqa: E402


def fusion_overlap_check(rank, world, dev):
    """The all-reduce launched from inside the fused backward (GradBucket.overlap_with: side stream, before the gate
    backward) must give the same averaged gradients as the exchange issued after the backward pass."""
    from glfusion_b200 import GlobalLocalFusion
    B, C, V, h, w = 2, 256, 2, 14, 14
    torch.manual_seed(5)
    f = GlobalLocalFusion(in_channels=C)
    g = torch.Generator().manual_seed(6)
    with torch.no_grad():
        for m in (f.global_attn, f.local_attn):
            for p_, mean in ((m.W_z[1].weight, 1.0), (m.W_z[1].bias, 0.0)):
                p_.copy_(mean + 0.2 * torch.randn(p_.shape, generator=g))
    f = f.to(dev).train()
    gi = torch.Generator().manual_seed(100 + rank)
    f4 = [torch.randn(B, C, h, w, generator=gi).to(dev, torch.bfloat16).requires_grad_(True) for _ in range(V)]
    cl = [torch.randn(B, 5, h, w, generator=gi).to(dev) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=gi).to(dev) for _ in range(V)]
    dz = torch.randn(B, V, h, w, C, generator=gi).to(dev, torch.bfloat16).permute(0, 4, 1, 2, 3)
    params = list(f.global_attn._plist()) + list(f.local_attn._plist())
    bucket = dp.GradBucket(params)
    bucket.bind([f.global_attn, f.local_attn])
    assert bucket.enable_p2p(), "IPC exchange failed"

    def step():
        for p_ in params:
            p_.grad = None
        for t in f4:
            t.grad = None
        f.forward_stacked(f4, cl, ct).backward(dz)
        bucket.allreduce_mean()
        torch.cuda.synchronize()
        return [p_.grad.detach().clone() for p_ in params]
    ref = step()
    bucket.overlap_with(f)
    for _ in range(3):
        got = step()
        for a, b_ in zip(got, ref):
            # (not bitwise: at this small batch the token contractions are split along the tokens and accumulate
            # with unordered fp32 atomics, so two runs of the same step differ in the last bf16 bits already)
            err = float((a - b_).norm() / b_.norm().clamp_min(1e-20))
            assert err < 5e-3, err
    assert bucket._ar_stream is not None          # the hook really launched the exchange from the backward pass
    f.on_weight_grads_ready = None
    bucket.close_p2p()


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
    fusion_overlap_check(rank, world, dev)
    dist.barrier()
    if rank == 0:
        print("P2P_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
