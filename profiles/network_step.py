"""BASELINE configs[2] (full network train step, 4-view 16-frame 112x112 clips): step time of the reference network
``Global_and_Local`` built from the staged reference sources (oracle/_ref), unpatched vs with the fusion blocks replaced
by the B200 path (``glfusion_b200.install``), one micro-batch = ONE clip (16 frames x 4 views = 64 ResNet50 passes with
the stride-1 stem; 32 clips per GPU do not fit as one micro-batch).  fwd + bwd (BCE sum loss) + Adam
step, CUDA-event timed, with the NVML clock record.

    python profiles/network_step.py > profiles/r02_network_step.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import glfusion_b200  # noqa: E402
from bench import ClockSampler  # noqa: E402
from oracle import build_ref  # noqa: E402

VIEWS = ["1", "2", "3", "4"]
DEV = "cuda:0"


def build(ours, patched):
    orig = ours.TPAVIModule
    try:
        if patched:
            glfusion_b200.install(ours)
        torch.manual_seed(0)
        net = ours.Global_and_Local(VIEWS)
    finally:
        ours.TPAVIModule = orig
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for blk in (net.global_attn, net.local_attn):
            for p, mean in ((blk.W_z[1].weight, 1.0), (blk.W_z[1].bias, 0.0)):
                p.copy_(mean + 0.2 * torch.randn(p.shape, generator=g))
    return net.to(DEV).train()


def time_step(net, autocast, frames=16, reps=5):
    g = torch.Generator().manual_seed(2)
    imgs = {v: torch.rand(frames, 1, 112, 112, generator=g).to(DEV) for v in VIEWS}
    tgt = {v: (torch.rand(frames, 5, 112, 112, generator=g) > 0.7).float().to(DEV) for v in VIEWS}
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=3e-4, weight_decay=1e-5)   # main.py:163-165
    ts = []
    for i in range(reps + 2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            mask, _, _, _ = net(imgs)
        loss = sum(torch.nn.functional.binary_cross_entropy_with_logits(mask[v].float(), tgt[v], reduction="sum") for v in VIEWS)
        loss.backward()
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    ours = build_ref.import_reference_network()
    if ours is None:
        print(json.dumps({"unavailable": "oracle/_ref was not staged"}))
        return
    out = {"workload": "Global_and_Local, 4 views, micro-batch = 1 clip of 16 frames at 112x112 (64 backbone passes), "
                       "fwd + bwd + Adam step, PyTorch eager for everything outside the fusion blocks",
           "unit": "ms per micro-batch (1 clip)"}
    sampler = ClockSampler(0)
    sampler.start()
    for patched in (False, True):
        net = build(ours, patched)
        for ac in (False, True):
            key = ("b200_fusion" if patched else "reference") + ("_bf16_autocast" if ac else "_fp32")
            try:
                out[key] = round(time_step(net, ac), 2)
            except Exception as exc:      # noqa: BLE001
                out[key] = f"failed: {exc!r}"[:200]
        del net
        torch.cuda.empty_cache()
    out["clocks"] = sampler.stop()
    for a, b in (("reference_fp32", "b200_fusion_fp32"), ("reference_bf16_autocast", "b200_fusion_bf16_autocast")):
        if isinstance(out.get(a), float) and isinstance(out.get(b), float):
            out["speedup_" + a.split("_", 1)[1]] = round(out[a] / out[b], 3)
            out["clips_per_s_" + b] = round(1e3 / out[b], 2)
            out["clips_per_s_" + a] = round(1e3 / out[a], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
