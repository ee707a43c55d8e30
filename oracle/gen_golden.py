"""Generate golden input/output vectors from the UNMODIFIED reference module.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

It imports ``/root/reference/GLfusion/models/TPAVI.py`` (torch-only; semantically identical to the copy at
R/models/ours.py:770-917), loads seeded weights into it through ``load_state_dict(strict=True)``, re-randomises
the zero-initialised BatchNorm gamma/beta and the LayerNorm affine, runs forward + autograd
backward on CPU in fp32 and writes ``tests/golden/<case>.npz`` with every input, output, gradient and BN buffer.

The glue case additionally restates, literally, lines 1802-1834 of R/models/ours.py around two reference
modules (the enclosing class cannot be constructed offline: it downloads ImageNet weights and imports monai).
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference/GLfusion"

from oracle import tpavi_oracle as O  # noqa: E402

CASES = {
    # name: (B, C, T, H, W, mode, training, bn_layer)
    "dot_train_c128": (2, 128, 2, 8, 8, "dot", True, True),
    "dot_train_c256": (2, 256, 3, 4, 8, "dot", True, True),
    "dot_eval_c128": (2, 128, 2, 8, 8, "dot", False, True),
    "dot_nobn_c128": (2, 128, 2, 8, 8, "dot", True, False),
    "embedded_train_c128": (2, 128, 2, 8, 8, "embedded", True, True),
    "embedded_eval_c128": (1, 128, 3, 8, 8, "embedded", False, True),
    "dot_ragged_c128": (3, 128, 3, 5, 7, "dot", True, True),     # N = 105: not a multiple of any tile
}


def load_reference_module():
    sys.path.insert(0, REF)
    from models.TPAVI import TPAVIModule  # type: ignore
    return TPAVIModule


def make_module(TPAVIModule, C, mode, bn_layer, params):
    m = TPAVIModule(in_channels=C, mode=mode, bn_layer=bn_layer)
    sd = {k: v.clone() for k, v in params.items()}
    m.load_state_dict(sd, strict=True)
    return m


def run_case(TPAVIModule, name, B, C, T, H, W, mode, training, bn_layer, seed):
    torch.manual_seed(seed)
    params = O.init_params(C, seed=seed, randomize_affine=True, bn_layer=bn_layer)
    if bn_layer and not training:
        # eval mode reads running stats: make them non-trivial
        g = torch.Generator().manual_seed(seed + 1)
        params["W_z.1.running_mean"] = torch.randn(C, generator=g) * 0.1
        params["W_z.1.running_var"] = torch.rand(C, generator=g) * 0.5 + 0.5
    m = make_module(TPAVIModule, C, mode, bn_layer, params)
    m.train(training)
    g = torch.Generator().manual_seed(seed + 2)
    x = torch.randn(B, C, T, H, W, generator=g)
    dz = torch.randn(B, C, T, H, W, generator=g)
    x.requires_grad_(True)
    z, audio_temp = m(x)
    assert audio_temp == 0
    z.backward(dz)
    out = {"x": x.detach().numpy(), "dz": dz.numpy(), "z": z.detach().numpy(), "dx": x.grad.numpy()}
    for k, v in params.items():
        out["param:" + k] = v.numpy()
    for k, v in m.named_parameters():
        if k.startswith("align_channel"):
            continue
        out["grad:" + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    for k, v in m.named_buffers():
        out["buf_after:" + k] = v.numpy()
    out["meta"] = np.array([B, C, T, H, W, int(training), int(bn_layer)], dtype=np.int64)
    out["mode"] = np.array(mode)
    z_strides = np.array(z.stride(), dtype=np.int64)
    out["z_strides"] = z_strides
    return out


def run_glue_case(TPAVIModule, seed=7):
    """ours.py:1802-1834 restated literally around two reference TPAVIModule instances."""
    from torch import nn
    B, C, V, h, w = 2, 128, 3, 6, 6
    pg = O.init_params(C, seed=seed, randomize_affine=True)
    pl = O.init_params(C, seed=seed + 100, randomize_affine=True)
    global_attn = make_module(TPAVIModule, C, "dot", True, pg)
    local_attn = make_module(TPAVIModule, C, "dot", True, pl)
    gen = torch.Generator().manual_seed(seed + 2)
    f4 = [torch.randn(B, C, h, w, generator=gen, requires_grad=True) for _ in range(V)]
    cls_l = [torch.randn(B, 5, h, w, generator=gen, requires_grad=True) for _ in range(V)]
    ctr_l = [torch.randn(B, 1, h, w, generator=gen, requires_grad=True) for _ in range(V)]
    d_out = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    center_aware_weight = 20
    mask_bb, ctr, f4_local = [], [], []
    for v in range(V):
        mb = nn.Sigmoid()(cls_l[v])                                    # :1803
        n, c, hh, ww = mb.shape
        mb = nn.AdaptiveMaxPool3d((1, hh, ww))(mb)                     # :1805-1806
        mask_bb.append(mb)
        ctr.append(nn.Sigmoid()(ctr_l[v]))                             # :1810
    for v in range(V):
        atten = (center_aware_weight * mask_bb[v] * ctr[v]).sigmoid()  # :1815
        f4_local.append(f4[v].clone() * atten)                         # :1816
    gcat = torch.cat([f.unsqueeze(2) for f in f4], dim=2)              # :1819-1820
    gfeat, _ = global_attn(gcat)                                       # :1821
    lcat = torch.cat([f.unsqueeze(2) for f in f4_local], dim=2)        # :1826-1827
    lfeat, _ = local_attn(lcat)                                        # :1828
    outs = [gfeat[:, :, i] + lfeat[:, :, i] for i in range(V)]         # :1823,1830,1834
    torch.autograd.backward(outs, d_out)
    out = {"meta": np.array([B, C, V, h, w], dtype=np.int64)}
    for v in range(V):
        out[f"f4:{v}"] = f4[v].detach().numpy()
        out[f"cls:{v}"] = cls_l[v].detach().numpy()
        out[f"ctr:{v}"] = ctr_l[v].detach().numpy()
        out[f"d_out:{v}"] = d_out[v].numpy()
        out[f"out:{v}"] = outs[v].detach().numpy()
        out[f"df4:{v}"] = f4[v].grad.numpy()
        out[f"dcls:{v}"] = cls_l[v].grad.numpy()
        out[f"dctr:{v}"] = ctr_l[v].grad.numpy()
    for tag, p, m in (("g", pg, global_attn), ("l", pl, local_attn)):
        for k, t in p.items():
            out[f"param_{tag}:{k}"] = t.numpy()
        for k, t in m.named_parameters():
            if not k.startswith("align_channel"):
                out[f"grad_{tag}:{k}"] = t.grad.numpy()
        for k, t in m.named_buffers():
            out[f"buf_after_{tag}:{k}"] = t.numpy()
    return out


def main():
    TPAVIModule = load_reference_module()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    for i, (name, cfg) in enumerate(CASES.items()):
        out = run_case(TPAVIModule, name, *cfg, seed=100 + i)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print(name, "z", out["z"].shape, "strides", out["z_strides"])
    out = run_glue_case(TPAVIModule)
    np.savez_compressed(os.path.join(outdir, "glue_dot_c128.npz"), **out)
    print("glue_dot_c128 done")


if __name__ == "__main__":
    main()
