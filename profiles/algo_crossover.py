#!/usr/bin/env python
"""Token-space vs Gram form of mode='dot' across channel counts and sequence lengths (one TPAVIModule, fwd+bwd, bf16
token-major input): where the library's rule (N >= 3 C at C = 256, 5 C below, 8 C above) puts the crossover.  python profiles/algo_crossover.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glfusion_b200 import TPAVIModule, tpavi  # noqa: E402

dev = "cuda:0"
out = {}
for C, B, T, H, W in ((256, 128, 4, 28, 28), (256, 256, 4, 20, 20), (256, 512, 4, 14, 14), (256, 1024, 2, 14, 14),
                      (512, 64, 4, 28, 28), (512, 128, 4, 20, 20), (512, 256, 4, 14, 14),
                      (1024, 32, 4, 28, 28), (1024, 16, 4, 40, 40), (128, 256, 4, 28, 28), (128, 1024, 2, 14, 14)):
    N = T * H * W
    torch.manual_seed(0)
    m = TPAVIModule(C).to(dev).train()
    with torch.no_grad():
        m.W_z[1].weight.normal_(1.0, 0.2)
        m.W_z[1].bias.normal_(0.0, 0.2)
    x = torch.randn(B, T, H, W, C, device=dev, dtype=torch.bfloat16).permute(0, 4, 1, 2, 3).requires_grad_(True)
    dz = torch.randn(B, T, H, W, C, device=dev, dtype=torch.bfloat16).permute(0, 4, 1, 2, 3)
    row = {}
    for name, algo in (("token", 1), ("gram", 2)):
        tpavi.DOT_ALGO = algo

        def step():
            x.grad = None
            for p in m.parameters():
                p.grad = None
            z, _ = m(x)
            z.backward(dz)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        row[name + "_ms"] = round(e0.elapsed_time(e1) / 10, 3)
    row["N_over_C"] = round(N / C, 2)
    row["auto_picks"] = "gram" if N >= (3 if C == 256 else (8 if C > 256 else 5)) * C else "token"
    out[f"C={C} N={N} B={B}"] = row
    print(f"C={C} N={N} B={B}", row, flush=True)
tpavi.DOT_ALGO = 0
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "crossover.json"), "w"), indent=1)
