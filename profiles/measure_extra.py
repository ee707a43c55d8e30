#!/usr/bin/env python
"""Extra measurements for SURVEY.md §8(d) (run on the B200 box; writes profiles-style JSON to stdout):

  1. GPU reference beside ours: the reference ALGORITHM (oracle port: literal N x N attention, autograd) in PyTorch eager
     on the same B200, fp32 and bf16 — "the bar to beat on the same box" (BASELINE.md §1).
  2. Our modules at the other layouts / widths: frames-in-sequence (Layout B, N = 50 176), C = 2048, mode='embedded',
     compute_precision='fp32'.
  3. BASELINE configs[3]: high-res stress, temporal-window sweep w in {1..32} at 64x64 tokens per view.
  4. BASELINE configs[4]: inference-only (eval, no grad) batch sweep.

Timing: CUDA events, 3 warm-ups, then `reps` iterations; inputs larger than L2 or an L2 flush between iterations.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import glfusion_b200  # noqa: E402
from glfusion_b200 import TPAVIModule  # noqa: E402
from oracle import tpavi_oracle as O  # noqa: E402

DEV = "cuda:0"
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    _flush.zero_()


def timed(fn, reps=5, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def module(C, mode="dot", precision="bf16", train=True):
    m = TPAVIModule(C, mode=mode)
    m.load_state_dict(O.init_params(C, seed=0, randomize_affine=True), strict=True)
    m.compute_precision = precision
    return m.to(DEV).train(train)


def ours_pair_fwd_bwd(B, C, T, H, W, mode="dot", precision="bf16", dtype=torch.bfloat16, reps=5):
    """MGFM + MLFM (two blocks) fwd+bwd on token-major inputs; returns ms."""
    mg, ml = module(C, mode, precision), module(C, mode, precision)
    x = torch.randn(B, T, H, W, C, device=DEV).to(dtype).permute(0, 4, 1, 2, 3).requires_grad_(True)
    dz = torch.randn(B, T, H, W, C, device=DEV).to(dtype).permute(0, 4, 1, 2, 3)

    def step():
        x.grad = None
        zg, _ = mg(x)
        zl, _ = ml(x)
        torch.autograd.backward([zg, zl], [dz, dz])
    return timed(step, reps=reps)


def ours_pair_infer(B, C, T, H, W, dtype=torch.bfloat16, reps=5):
    mg, ml = module(C, train=False), module(C, train=False)
    x = torch.randn(B, T, H, W, C, device=DEV).to(dtype).permute(0, 4, 1, 2, 3)

    def step():
        with torch.no_grad():
            mg(x)
            ml(x)
    return timed(step, reps=reps)


def eager_reference(B, C, T, H, W, dtype, reps=3):
    """Reference algorithm (N x N materialised) in PyTorch eager on the GPU: two blocks fwd+bwd."""
    ps = [{k: (v.to(DEV, dtype) if v.is_floating_point() else v.to(DEV)) for k, v in O.init_params(C, seed=s).items()}
          for s in (0, 1)]
    x = torch.randn(B, C, T, H, W, device=DEV, dtype=dtype)
    dz = torch.randn(B, C, T, H, W, device=DEV, dtype=dtype)

    def step():
        for p in ps:
            O.tpavi_fwd_bwd(x, dz, p, mode="dot")
    return timed(step, reps=reps, warm=2)


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    V, F, h, w = 4, 16, 28, 28
    # 1. reference algorithm, PyTorch eager, same GPU (1 clip = 16 sequences x 3136 tokens)
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        ms = eager_reference(F, 256, V, h, w, dt)
        out[f"eager_reference_{name}_cfg2_layoutA"] = {"ms_per_clip": round(ms, 3), "clips_per_s": round(1e3 / ms, 2)}
    # 2. ours, modules only (token-major input, no gate kernel), several regimes
    clips = 8
    ms = ours_pair_fwd_bwd(clips * F, 256, V, h, w)
    out["ours_bf16_cfg2_layoutA_modules_only"] = {"clips": clips, "ms": round(ms, 3), "clips_per_s": round(clips * 1e3 / ms, 1)}
    ms = ours_pair_fwd_bwd(clips, 256, V * F, h, w)
    out["ours_bf16_cfg2_layoutB_frames_in_sequence"] = {"clips": clips, "N": V * F * h * w, "ms": round(ms, 3),
                                                        "clips_per_s": round(clips * 1e3 / ms, 1)}
    ms = ours_pair_fwd_bwd(2 * F, 2048, V, h, w)
    out["ours_bf16_C2048_layoutA"] = {"clips": 2, "ms": round(ms, 3), "clips_per_s": round(2 * 1e3 / ms, 1),
                                      "tflops_algorithmic": round(2 * 13.5 * (2 * F * V * h * w) * 2048 * 2048 / (ms * 1e-3) / 1e12, 1)}
    ms = ours_pair_fwd_bwd(2 * F, 256, V, h, w, precision="fp32", dtype=torch.float32)
    out["ours_fp32x3_cfg2_layoutA"] = {"clips": 2, "ms": round(ms, 3), "clips_per_s": round(2 * 1e3 / ms, 1)}
    ms = ours_pair_fwd_bwd(F, 256, V, h, w, mode="embedded", reps=3)
    out["ours_bf16_embedded_cfg2_layoutA"] = {"clips": 1, "ms": round(ms, 3), "clips_per_s": round(1e3 / ms, 2)}
    # 3. cfg4: 4 views x 32 frames, 64x64 tokens per view, window of w frames per sequence
    sweep = {}
    for win in (1, 2, 4, 8, 16, 32):
        ms = ours_pair_fwd_bwd(32 // win, 256, 4 * win, 64, 64, reps=3)
        sweep[str(win)] = {"B": 32 // win, "N": 4 * win * 4096, "ms": round(ms, 3), "clips_per_s": round(1e3 / ms, 2)}
    out["cfg4_window_sweep_bf16_dot"] = sweep
    # 4. cfg5: inference-only sweep (eval mode, BN folded to running stats, no activations kept)
    inf = {}
    for c in (1, 4, 16, 64):
        ms = ours_pair_infer(c * F, 256, V, h, w, reps=3)
        inf[str(c)] = {"ms": round(ms, 3), "clips_per_s": round(c * 1e3 / ms, 1)}
    out["cfg5_inference_sweep_bf16"] = inf
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
