"""Golden vectors for the cycle-consistency loss from the UNMODIFIED reference methods.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden_cycle.py

``Trainer.seg_cycle`` / ``Trainer.dense_seg_cycle`` live in R/main.py:650-798, a script whose imports (nibabel, monai,
tensorboardX, the dataset loaders) are not installed here.  The two method definitions are therefore compiled from the
reference file where it lies (ast: the two FunctionDef nodes of ``class Trainer``, nothing is copied into the repo),
bound to a stand-in ``self`` that only carries ``device``, and run on CPU with autograd.  ``np.random`` is seeded before
``seg_cycle`` so that its ``np.random.choice`` draw is reproducible; the draw is recorded in the fixture.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_MAIN = "/root/reference/GLfusion/main.py"

# name: (T frames, C, target_region, cyc_off, chunk_size, temperature, kind, kwargs, seed)
CASES = {
    "cycle_single_t32_c256": (32, 256, 16, 2, 3, 10.0, "single", {}, 3),
    "cycle_single_t40_c64": (40, 64, 16, 2, 3, 10.0, "single", {}, 4),
    "cycle_dense_t32_c256": (32, 256, 16, 2, 3, 10.0, "dense", {"soft_label": False, "is_overlap": True}, 5),
    "cycle_dense_soft_t36_c128": (36, 128, 16, 2, 3, 10.0, "dense", {"soft_label": True, "is_overlap": True}, 6),
    "cycle_dense_nooverlap_t30_c96": (30, 96, 12, 1, 4, 5.0, "dense", {"soft_label": False, "is_overlap": False}, 7),
}


def load_reference_methods(path: str = REF_MAIN):
    """The two reference methods as plain functions f(self, feat_out, ...)."""
    tree = ast.parse(open(path).read(), filename=path)
    wanted = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "Trainer":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in ("seg_cycle", "dense_seg_cycle"):
                    wanted[item.name] = item
    if len(wanted) != 2:
        raise RuntimeError("reference main.py: Trainer.seg_cycle / dense_seg_cycle not found")
    mod = ast.Module(body=list(wanted.values()), type_ignores=[])
    ns = {"torch": torch, "np": np, "numpy": np}
    exec(compile(mod, path, "exec"), ns)
    return ns["seg_cycle"], ns["dense_seg_cycle"]


def features(T: int, C: int, seed: int) -> torch.Tensor:
    """Per-frame features with the scale spatial sums of LayerNorm outputs have (|x| ~ sqrt(h*w)), and a slow drift
    along the frames so that the softmax over the key positions is neither flat nor one-hot."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(1, C, generator=g)
    drift = torch.cumsum(torch.randn(T, C, generator=g) * 0.35, dim=0)
    return (base + drift) * 3.0


def run_case(seg_cycle, dense_seg_cycle, name, T, C, R, off, ch, temp, kind, kw, seed):
    me = types.SimpleNamespace(device=torch.device("cpu"))
    feat = features(T, C, seed).requires_grad_(True)
    start = -1
    if kind == "single":
        np.random.seed(seed)
        start = int(np.random.choice(R - (ch + off) + 1))        # the draw the method is about to make
        np.random.seed(seed)
        loss = seg_cycle(me, feat, target_region=R, cyc_off=off, chunk_size=ch, temperature=temp)
    else:
        loss = dense_seg_cycle(me, feat, target_region=R, cyc_off=off, chunk_size=ch, temperature=temp, **kw)
    loss.backward()
    return {"feat": feat.detach().numpy(), "loss": np.float64(loss.item()), "dfeat": feat.grad.numpy(),
            "target_region": np.int64(R), "cyc_off": np.int64(off), "chunk_size": np.int64(ch),
            "temperature": np.float64(temp), "target_strtpt": np.int64(start), "np_seed": np.int64(seed),
            "soft_label": np.int64(bool(kw.get("soft_label", False))), "is_overlap": np.int64(bool(kw.get("is_overlap", True))),
            "dense": np.int64(kind == "dense")}


def main():
    seg_cycle, dense_seg_cycle = load_reference_methods()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (T, C, R, off, ch, temp, kind, kw, seed) in CASES.items():
        rec = run_case(seg_cycle, dense_seg_cycle, name, T, C, R, off, ch, temp, kind, kw, seed)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(f"{name}: loss {float(rec['loss']):.6f}  |dfeat|max {np.abs(rec['dfeat']).max():.3e}  start {int(rec['target_strtpt'])}")


if __name__ == "__main__":
    main()
