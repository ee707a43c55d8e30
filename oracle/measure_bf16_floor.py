"""The reference's OWN bf16 noise floor, per tensor (TEST INFRASTRUCTURE; run in the build container only).

north_star allows 2e-2 relative error for the bf16 arm.  Some gradients of the reference module are ill-conditioned in
bf16 no matter who computes them (theta.bias: the gradient of a bias that enters only through Theta M, a sum of
cancelling terms).  This script runs the UNMODIFIED reference ``R/models/TPAVI.py`` (mode='dot', train) twice on the same
seeded cfg2-geometry input — in fp32 and under ``torch.autocast(bfloat16)`` — and records the relative L2 error of every
output / gradient of the autocast run against the fp32 run.  ``tests/`` then bound our bf16 arm by
``max(2e-2, floor[tensor])`` instead of a hand-widened constant.

    python oracle/measure_bf16_floor.py        # writes tests/golden/bf16_floor.json
"""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from oracle import tpavi_oracle as O  # noqa: E402


def run(TPAVI, x, dz, params, autocast: bool, mode: str = "dot"):
    m = TPAVI(in_channels=x.shape[1], mode=mode)
    m.load_state_dict({k: v.clone() for k, v in params.items()}, strict=True)
    m.train()
    xx = x.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        z, _ = m(xx)
    z.float().backward(dz)
    out = {"z": z.detach().float(), "dx": xx.grad.detach().float()}
    for k, p in m.named_parameters():
        if p.grad is not None and not k.startswith("align_channel"):
            out[k] = p.grad.detach().float()
    return out


def main():
    assert build_ref.build_ref(), "needs /root/reference (build container)"
    TPAVI = build_ref.load_reference_tpavi()
    torch.manual_seed(0)
    res = {}
    for tag, (B, C, T, H, W, mode) in {"cfg2_B2": (2, 256, 4, 28, 28, "dot"), "c128_B2": (2, 128, 4, 14, 14, "dot"),
                                       "embedded_c128_B2": (2, 128, 4, 14, 14, "embedded"),
                                       "embedded_cfg2_B1": (1, 256, 4, 28, 28, "embedded")}.items():
        g = torch.Generator().manual_seed(7)
        x = torch.randn(B, C, T, H, W, generator=g)
        dz = torch.randn(B, C, T, H, W, generator=g)
        p = O.init_params(C, seed=5, randomize_affine=True)
        a = run(TPAVI, x, dz, p, False, mode)
        b = run(TPAVI, x, dz, p, True, mode)
        res[tag] = {k: float((b[k] - a[k]).norm() / a[k].norm().clamp_min(1e-30)) for k in a}
    res["_doc"] = ("relative L2 error of the unmodified reference module under torch.autocast(bfloat16) against its own "
                   "fp32 run (train, seeded input; tags starting with 'embedded' are mode='embedded', the others mode='dot'); W_z.0.bias is analytically zero (BatchNorm cancels it)")
    path = os.path.join(ROOT, "tests", "golden", "bf16_floor.json")
    with open(path, "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
