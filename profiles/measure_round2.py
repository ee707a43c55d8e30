#!/usr/bin/env python
"""Round-2 measurements beside the bench line (run on the B200 box; JSON on stdout, with the NVML clock record):

  parts     SURVEY 8(f)3: `forward_parts` (fused node + the MGFM part stored beside the sum) against `forward_stacked`,
            and the cycle-consistency pass alone (loss on the spatial sums of f4_global_fusion, R/main.py:229-237: the
            backward skips the MLFM block), 8 clips of cfg2, fwd + bwd, replayed from a CUDA graph.
  cfg5      BASELINE configs[4]: inference-only sweep, 1 ... 256 clips per call (eval BatchNorm, no grad), bf16 arm and
            the fp32-exact arm, with the relative error of each against the oracle's eval forward on 1 clip.
  width2048 the reference network's own width (R/models/ours.py:1746-1747: C = 2048, 3 views x 28 x 28 = 2352 tokens,
            8 frames as batch): fwd + bwd of the fused call site, clips/s and algorithmic TFLOP/s against the sustained
            tensor peak (token-space form: N < 5 C).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ClockSampler, peaks, randomize_affine_  # noqa: E402
from glfusion_b200 import GlobalLocalFusion  # noqa: E402
from oracle import tpavi_oracle as O  # noqa: E402

DEV = "cuda:0"
F = 16


def fusion(C, train=True, precision="bf16"):
    torch.manual_seed(0)
    f = GlobalLocalFusion(in_channels=C)
    randomize_affine_(f.global_attn, 10)
    randomize_affine_(f.local_attn, 11)
    f.global_attn.compute_precision = f.local_attn.compute_precision = precision
    return f.to(DEV).train(train)


def inputs(B, C, V, h, w, dtype=torch.bfloat16, seed=1):
    g = torch.Generator().manual_seed(seed)
    f4 = [torch.randn(B, C, h, w, generator=g).to(DEV, dtype) for _ in range(V)]
    cl = [torch.randn(B, 5, h, w, generator=g).to(DEV) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=g).to(DEV) for _ in range(V)]
    return f4, cl, ct


def graph_time(compute, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            compute()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        compute()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def eager_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def parts_section(out):
    clips, C, V, h, w = 8, 256, 4, 28, 28
    B = clips * F
    f = fusion(C)
    f4, cl, ct = inputs(B, C, V, h, w)
    for t in f4:
        t.requires_grad_(True)
    keys = [str(i) for i in range(V)]
    dz = torch.randn(B, V, h, w, C, device=DEV).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
    dsum = torch.randn(B, C, device=DEV)
    params = [p for p in f.parameters() if p.requires_grad]

    def zero():
        for t in f4:
            t.grad = None
        for p in params:
            p.grad = None

    def stacked():
        zero()
        f.forward_stacked(f4, cl, ct).backward(dz)

    def dict_api():           # the reference-shaped call: dicts keyed by view, one gradient per view
        zero()
        o = f(dict(zip(keys, f4)), dict(zip(keys, cl)), dict(zip(keys, ct)))
        torch.autograd.backward([o[k] for k in keys], [dz[:, :, i] for i in range(V)])

    def parts_seg():          # the supervised pass of the trainer: parts returned, only the fused sum carries a gradient
        zero()
        fus, glob, _ = f.forward_parts(dict(zip(keys, f4)), dict(zip(keys, cl)), dict(zip(keys, ct)), need_local=False)
        torch.autograd.backward([fus[k] for k in keys], [dz[:, :, i] for i in range(V)])

    def parts_cycle():        # the cycle pass: only the spatial sums of the MGFM part carry a gradient
        zero()
        fus, glob, _ = f.forward_parts(dict(zip(keys, f4)), dict(zip(keys, cl)), dict(zip(keys, ct)), need_local=False)
        sum((glob[k].float().sum(dim=(2, 3)) * dsum).sum() for k in keys).backward()

    from glfusion_b200 import cycle

    def parts_cycle_fused():  # the same pass through the library's cycle step: spatial_sum + dense_seg_cycle kernels
        zero()
        fus, glob, _ = f.forward_parts(dict(zip(keys, f4)), dict(zip(keys, cl)), dict(zip(keys, ct)), need_local=False)
        sum(cycle.dense_seg_cycle(cycle.spatial_sum(glob[k]), 16, 2, 3, 10.0) for k in keys).backward()
    f4cl = [t.detach().clone().contiguous(memory_format=torch.channels_last).requires_grad_(True) for t in f4]

    def stacked_cl():         # channels_last views -> the row kernels, no transposition either way
        for t in f4cl:
            t.grad = None
        for p in params:
            p.grad = None
        f.forward_stacked(f4cl, cl, ct).backward(dz)
    res = {}
    for name, fn in (("forward_stacked", stacked), ("forward_stacked_channels_last_inputs", stacked_cl), ("forward_dict_api", dict_api), ("forward_parts_supervised_pass", parts_seg),
                     ("forward_parts_cycle_pass", parts_cycle), ("forward_parts_cycle_pass_fused_loss", parts_cycle_fused)):
        ms = graph_time(fn)
        res[name] = {"ms_per_step": round(ms, 4), "clips_per_s": round(clips / (ms * 1e-3), 1)}
    # the same step WITHOUT a CUDA graph (nn.DataParallel callers of the reference are eager): device time per step and
    # the host time it takes to enqueue one (blob allocation, parameter tables, ~100 tensor-map encodes, 27 launches)
    import time
    for _ in range(5):
        stacked()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(20):
        stacked()
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    res["forward_stacked_eager"] = {"ms_per_step": round(e0.elapsed_time(e1) / 20, 4),
                                    "host_enqueue_ms_per_step": round((t1 - t0) * 1e3 / 20, 4)}
    feat = torch.randn(B, C, device=DEV).cumsum(0).requires_grad_(True)

    def loss_alone():         # dense_seg_cycle forward + backward on [128 frames, 256] features: 2 launches + 1 multiply
        feat.grad = None
        cycle.dense_seg_cycle(feat, 16, 2, 3, 10.0).backward()
    res["dense_seg_cycle_fwd_bwd_us"] = round(graph_time(loss_alone) * 1e3, 2)
    res["parts_over_stacked"] = round(res["forward_parts_supervised_pass"]["ms_per_step"] / res["forward_stacked"]["ms_per_step"], 4)
    res["parts_over_dict_api"] = round(res["forward_parts_supervised_pass"]["ms_per_step"] / res["forward_dict_api"]["ms_per_step"], 4)
    out["parts"] = res


def cfg5_section(out):
    C, V, h, w = 256, 4, 28, 28
    res = {}
    pg = O.init_params(C, seed=0, randomize_affine=True)
    pl = O.init_params(C, seed=1, randomize_affine=True)
    for prec, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        f = GlobalLocalFusion(in_channels=C)
        f.global_attn.load_state_dict(pg, strict=True)
        f.local_attn.load_state_dict(pl, strict=True)
        f.global_attn.compute_precision = f.local_attn.compute_precision = prec
        f = f.to(DEV).eval()
        # tolerance report on one clip against the oracle's eval forward
        g = torch.Generator().manual_seed(5)
        f4c = [torch.randn(F, C, h, w, generator=g) for _ in range(V)]
        clc = [torch.randn(F, 5, h, w, generator=g) for _ in range(V)]
        ctc = [torch.randn(F, 1, h, w, generator=g) for _ in range(V)]
        ref = O.global_local_fusion(f4c, clc, ctc, pg, pl, training=False)
        with torch.no_grad():
            o = f.forward_stacked([t.to(DEV, dt) for t in f4c], [t.to(DEV) for t in clc], [t.to(DEV) for t in ctc])
        err = max(O.rel_err(o[:, :, v].float().cpu(), ref[v]) for v in range(V))
        sweep = {}
        for clips in ((1, 4, 16, 64, 256) if prec == "bf16" else (1, 4, 16)):
            f4, cl, ct = inputs(clips * F, C, V, h, w, dtype=dt)

            def run():
                with torch.no_grad():
                    f.forward_stacked(f4, cl, ct)
            ms = eager_time(run)
            sweep[str(clips)] = {"ms": round(ms, 3), "clips_per_s": round(clips * 1e3 / ms, 1)}
            del f4, cl, ct
            torch.cuda.empty_cache()
        res[prec] = {"max_rel_err_vs_oracle_eval": float(f"{err:.3e}"), "sweep": sweep}
    out["cfg5_inference"] = res


def width_section(out):
    C, V, h, w, B = 2048, 3, 28, 28, 8
    f = fusion(C)
    f4, cl, ct = inputs(B, C, V, h, w)
    for t in f4:
        t.requires_grad_(True)
    dz = torch.randn(B, V, h, w, C, device=DEV).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
    params = [p for p in f.parameters() if p.requires_grad]

    def step():
        for t in f4:
            t.grad = None
        for p in params:
            p.grad = None
        f.forward_stacked(f4, cl, ct).backward(dz)
    ms = graph_time(step, reps=10)
    rows = B * V * h * w
    flops = 2 * 13.5 * rows * C * C           # token-space form, both blocks, fwd + bwd
    pk = peaks()
    out["width2048"] = {"shape": "C=2048, 3 views x 28x28 = 2352 tokens, 8 frames as batch (the reference network's fusion input)",
                        "ms_per_step": round(ms, 3), "frames_per_s": round(B * 1e3 / ms, 1),
                        "tflops_algorithmic": round(flops / (ms * 1e-3) / 1e12, 1),
                        "frac_of_sustained_tensor_peak": round(flops / (ms * 1e-3) / 1e12 / pk["bf16_tflops"], 3)}


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    sampler = ClockSampler(0)
    sampler.start()
    for sec in (parts_section, cfg5_section, width_section):
        try:
            sec(out)
        except Exception as exc:       # noqa: BLE001
            out[sec.__name__] = f"failed: {exc!r}"[:300]
        torch.cuda.empty_cache()
    out["clocks"] = sampler.stop()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
