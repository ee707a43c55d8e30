#!/usr/bin/env python
"""Relative errors of the CUDA path against the CPU oracle (fp32 restatement of the reference) at the cfg2 token
geometry, for both exact reassociations of mode='dot'.  python profiles/numerics_report.py > report.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glfusion_b200  # noqa: E402
from glfusion_b200 import tpavi  # noqa: E402
from oracle import tpavi_oracle as O  # noqa: E402

dev = "cuda:0"
B, C, V, h, w = 2, 256, 4, 28, 28
pg = O.init_params(C, seed=31, randomize_affine=True)
pl = O.init_params(C, seed=32, randomize_affine=True)
gen = torch.Generator().manual_seed(33)
f4 = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
cl = [torch.randn(B, 5, h, w, generator=gen) for _ in range(V)]
ct = [torch.randn(B, 1, h, w, generator=gen) for _ in range(V)]
do = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
outs, df4, dcls, dctr, gg, gl = O.fusion_fwd_bwd(f4, cl, ct, do, {k: v.clone() for k, v in pg.items()},
                                                 {k: v.clone() for k, v in pl.items()})
rep = {"shape": {"B": B, "C": C, "V": V, "h": h, "w": w}, "tolerance_bf16": 2e-2}
for name, algo in (("token", 1), ("gram", 2)):
    tpavi.DOT_ALGO = algo
    f = glfusion_b200.GlobalLocalFusion(in_channels=C)
    f.global_attn.load_state_dict({k: v.clone() for k, v in pg.items()}, strict=True)
    f.local_attn.load_state_dict({k: v.clone() for k, v in pl.items()}, strict=True)
    f = f.to(dev).train()
    f4d = [t.to(dev, torch.bfloat16).requires_grad_(True) for t in f4]
    cld = [t.to(dev).requires_grad_(True) for t in cl]
    ctd = [t.to(dev).requires_grad_(True) for t in ct]
    out = f.forward_stacked(f4d, cld, ctd)
    out.backward(torch.stack(do, dim=2).to(dev, torch.bfloat16))
    torch.cuda.synchronize()
    r = {"out": max(O.rel_err(out[:, :, v], outs[v]) for v in range(V)),
         "df4": max(O.rel_err(f4d[v].grad, df4[v]) for v in range(V)),
         "dctr": max(O.rel_err(ctd[v].grad, dctr[v]) for v in range(V))}
    for tag, mod, ref in (("g", f.global_attn, gg), ("l", f.local_attn, gl)):
        for k, p in mod.named_parameters():
            if k.startswith("align_channel") or k == "W_z.0.bias":
                continue
            r[f"grad_{tag}:{k}"] = O.rel_err(p.grad, ref[k])
    rep[name] = {k: float(f"{v:.3e}") for k, v in r.items()}
    rep[name]["max_param_grad"] = max(v for k, v in rep[name].items() if k.startswith("grad_"))
print(json.dumps(rep, indent=1))
