"""Segmentation Dice agreement (north_star: within 0.1 points) between the reference algorithm (fp32 CPU oracle) and the
CUDA path (bf16 and fp32 arms), with the Dice definition of R/main.py:800-815.

At random init the logits hover around 0 and `sigmoid > 0.5` flips on noise, so the fixture plants a
segmentation signal: the target blob is written into a group of feature channels and the head reads the same channels of
the fused feature map — both implementations then segment the blob and their Dice scores are comparable."""
import pytest
import torch

from gpu_util import DEV
from glfusion_b200 import GlobalLocalFusion
from oracle import tpavi_oracle as O

pytestmark = pytest.mark.gpu


def _fixture(seed=0):
    B, C, V, h, w = 4, 128, 2, 28, 28
    gen = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    targets, f4 = [], []
    for v in range(V):
        cy = torch.randint(8, 20, (B,), generator=gen)
        cx = torch.randint(8, 20, (B,), generator=gen)
        r = torch.randint(4, 8, (B,), generator=gen)
        mask = (((yy[None] - cy[:, None, None]) ** 2 + (xx[None] - cx[:, None, None]) ** 2) < (r[:, None, None] ** 2)).float()
        feat = torch.randn(B, C, h, w, generator=gen)
        feat[:, :16] += 4.0 * mask[:, None] - 2.0
        targets.append(mask)
        f4.append(feat)
    cls = [torch.randn(B, 5, h, w, generator=gen) for _ in range(V)]
    ctr = [torch.randn(B, 1, h, w, generator=gen) for _ in range(V)]
    return B, C, V, h, w, f4, cls, ctr, targets


def _head(fused):           # fixed linear head: mean of the signal channels
    return fused[:, :16].float().mean(dim=1)


@pytest.mark.parametrize("precision,dtype", [("bf16", torch.bfloat16), ("fp32", torch.float32)])
def test_dice_agreement_with_reference(precision, dtype):
    B, C, V, h, w, f4, cls, ctr, targets = _fixture()
    pg = O.init_params(C, seed=81, randomize_affine=True)
    pl = O.init_params(C, seed=82, randomize_affine=True)
    for p in (pg, pl):      # keep the LayerNorm affine neutral so the planted channels keep their sign
        p["norm_layer.weight"].fill_(1.0)
        p["norm_layer.bias"].fill_(0.0)
    ref = O.global_local_fusion(f4, cls, ctr, {k: v.clone() for k, v in pg.items()}, {k: v.clone() for k, v in pl.items()})
    f = GlobalLocalFusion(in_channels=C)
    f.global_attn.load_state_dict(pg, strict=True)
    f.local_attn.load_state_dict(pl, strict=True)
    f.global_attn.compute_precision = f.local_attn.compute_precision = precision
    f = f.to(DEV).train()
    with torch.no_grad():
        out = f.forward_stacked([t.to(DEV, dtype) for t in f4], [t.to(DEV) for t in cls], [t.to(DEV) for t in ctr])
    torch.cuda.synchronize()
    d_ref, d_our = [], []
    for v in range(V):
        d_ref.append(O.dice(_head(ref[v]), targets[v]))
        d_our.append(O.dice(_head(out[:, :, v].cpu()), targets[v]))
    mean_ref, mean_our = sum(d_ref) / V, sum(d_our) / V
    assert mean_ref > 0.6, f"fixture does not segment ({mean_ref})"
    # 0.1 Dice points = 0.001 in [0, 1]
    assert abs(mean_ref - mean_our) <= 1e-3, f"Dice ref {100 * mean_ref:.3f} vs ours {100 * mean_our:.3f}"
