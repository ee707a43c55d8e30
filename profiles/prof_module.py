#!/usr/bin/env python
"""One TPAVIModule fwd+bwd on cfg2-shaped token-major input, for ncu captures and per-mode timing.

    python profiles/prof_module.py --mode embedded --seqs 16 --reps 3            # prints ms per fwd+bwd
    ncu --set full -k regex:flash ... python profiles/prof_module.py --mode embedded --reps 1 --warm 1

Attention FLOPs (mode='embedded', per module): forward 4*B*N^2*Ci, backward (recompute, 7 products) 14*B*N^2*Ci.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glfusion_b200 import TPAVIModule  # noqa: E402
from oracle import tpavi_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="embedded")
    ap.add_argument("--seqs", type=int, default=16, help="sequences (B): 16 = one cfg2 clip in frames-as-batch layout")
    ap.add_argument("--T", type=int, default=4)
    ap.add_argument("--hw", type=int, default=28)
    ap.add_argument("--channels", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warm", type=int, default=3)
    a = ap.parse_args()
    dev = "cuda:0"
    C = a.channels
    m = TPAVIModule(C, mode=a.mode)
    m.load_state_dict(O.init_params(C, seed=0, randomize_affine=True), strict=True)
    m = m.to(dev).train()
    B, T, H, W = a.seqs, a.T, a.hw, a.hw
    x = torch.randn(B, T, H, W, C, device=dev).bfloat16().permute(0, 4, 1, 2, 3).requires_grad_(True)
    dz = torch.randn(B, T, H, W, C, device=dev).bfloat16().permute(0, 4, 1, 2, 3)

    def fwd():
        return m(x)[0]

    def step():
        x.grad = None
        z = fwd()
        z.backward(dz)

    for _ in range(a.warm):
        step()
    torch.cuda.synchronize()
    tf, tb = [], []
    for _ in range(a.reps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        x.grad = None
        e[0].record()
        z = fwd()
        e[1].record()
        z.backward(dz)
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]))
        tb.append(e[1].elapsed_time(e[2]))
    tf.sort()
    tb.sort()
    N = T * H * W
    Ci = C // 2
    out = {"mode": a.mode, "B": B, "N": N, "C": C, "fwd_ms": round(tf[len(tf) // 2], 4),
           "bwd_ms": round(tb[len(tb) // 2], 4)}
    if a.mode == "embedded":
        out["attn_fwd_gflop"] = round(4 * B * N * N * Ci / 1e9, 2)
        out["attn_bwd_gflop_executed"] = round(14 * B * N * N * Ci / 1e9, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
