"""The cycle-consistency step on the GPU (glfusion_b200.cycle, csrc/glf_cycle.cu) against the reference's own golden
vectors (R/main.py:650-798 via oracle/gen_golden_cycle.py) and the fp64 oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from gpu_util import DEV
from glfusion_b200 import cycle
from oracle import cycle_oracle as CO

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cycle_*.npz")))
TOL = 1e-4       # fp32 kernel vs the fp64 oracle / the reference's fp32 autograd


def run_gpu(d, feat):
    R, off, ch, temp = int(d["target_region"]), int(d["cyc_off"]), int(d["chunk_size"]), float(d["temperature"])
    if int(d["dense"]):
        return cycle.dense_seg_cycle(feat, R, off, ch, temp, soft_label=bool(d["soft_label"]), is_overlap=bool(d["is_overlap"]))
    np.random.seed(int(d["np_seed"]))                     # the reference's own draw (np.random.choice, main.py:655)
    return cycle.seg_cycle(feat, R, off, ch, temp)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cycle_loss_matches_reference_golden(path):
    d = np.load(path)
    feat = torch.from_numpy(d["feat"]).to(DEV).requires_grad_(True)
    loss = run_gpu(d, feat)
    (3.0 * loss).backward()
    assert abs(loss.item() - float(d["loss"])) <= TOL * abs(float(d["loss"]))
    err = np.abs(feat.grad.cpu().numpy() / 3.0 - d["dfeat"]).max() / np.abs(d["dfeat"]).max()
    assert err <= TOL, err
    # deterministic: fixed-order sums, no atomics
    feat2 = torch.from_numpy(d["feat"]).to(DEV).requires_grad_(True)
    loss2 = run_gpu(d, feat2)
    (3.0 * loss2).backward()
    assert torch.equal(loss, loss2) and torch.equal(feat.grad, feat2.grad)


@pytest.mark.parametrize("T,Cn,R,off,ch", [(48, 2048, 16, 2, 3), (21, 40, 8, 0, 1), (300, 64, 200, 3, 5)])
def test_cycle_loss_against_oracle_other_sizes(T, Cn, R, off, ch):
    """The network's own width (C = 2048, R/models/ours.py:1746), the smallest legal geometry, and many positions."""
    g = torch.Generator().manual_seed(T)
    feat = (torch.cumsum(torch.randn(T, Cn, generator=g) * 0.3, 0) * 2.0)
    for dense in (False, True):
        x = feat.to(DEV).requires_grad_(True)
        if dense:
            loss = cycle.dense_seg_cycle(x, R, off, ch, 10.0, soft_label=(R - off - ch + 1) > 1)
            lo, go = CO.dense_seg_cycle(feat.numpy(), R, off, ch, 10.0, soft_label=(R - off - ch + 1) > 1)
        else:
            s = (R - off - ch) // 2
            loss = cycle.seg_cycle(x, R, off, ch, 10.0, target_strtpt=s)
            lo, go = CO.seg_cycle(feat.numpy(), R, off, ch, 10.0, s)
        loss.backward()
        assert abs(loss.item() - lo) <= TOL * max(abs(lo), 1e-3)
        assert np.abs(x.grad.cpu().numpy() - go).max() <= TOL * max(np.abs(go).max(), 1e-12)


def test_cycle_loss_rejects_bad_geometry():
    from glfusion_b200._lib import GlfError
    x = torch.zeros(18, 64, device=DEV)
    with pytest.raises(GlfError):
        cycle.seg_cycle(x, 16, 2, 3, 10.0, target_strtpt=0)       # key region shorter than chunk + offset
    with pytest.raises(GlfError):
        cycle.seg_cycle(torch.zeros(40, 64, device=DEV), 16, 2, 3, 10.0, target_strtpt=12)    # start outside the positions
    with pytest.raises(GlfError):
        cycle.seg_cycle(torch.zeros(40, 64), 16, 2, 3, 10.0, target_strtpt=0)                 # no CPU path


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("layout", ["token_view", "nchw", "odd"])
def test_spatial_sum_matches_fp64(dtype, layout):
    B, Cn, h, w = 5, 256, 28, 28
    g = torch.Generator().manual_seed(2)
    if layout == "token_view":                     # what the fusion path returns: a view of the token-major stack
        big = torch.randn(B, 3, h, w, Cn, generator=g).to(DEV).to(dtype)
        x = big[:, 1].permute(0, 3, 1, 2)
    elif layout == "nchw":
        x = torch.randn(B, Cn, h, w, generator=g).to(DEV).to(dtype)
    else:                                          # ragged, unaligned channels-last
        Cn, h, w = 37, 5, 7
        x = torch.randn(B, h, w, Cn + 1, generator=g).to(DEV).to(dtype)[..., 1:].permute(0, 3, 1, 2)
    x = x.detach().requires_grad_(True)
    out = cycle.spatial_sum(x)
    want = CO.spatial_sum(x.detach().float().cpu().numpy())
    assert out.dtype == torch.float32 and out.shape == (B, Cn)
    assert np.abs(out.detach().cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max() + 1e-4
    wgt = torch.randn(B, Cn, generator=g).to(DEV)
    (out * wgt).sum().backward()
    assert torch.equal(x.grad.float(), wgt.to(dtype).float()[:, :, None, None].expand(B, Cn, h, w))


def test_cycle_pass_through_the_fusion_node():
    """forward_parts -> spatial_sum -> dense_seg_cycle -> backward, against the same chain written with the torch ops
    the reference trainer uses (sum(dim=(2, 3)), R/main.py:229) and the oracle's loss gradient."""
    from glfusion_b200 import GlobalLocalFusion
    from bench import randomize_affine_
    torch.manual_seed(0)
    B, Cn, V, h, w = 36, 128, 2, 6, 6
    fus = GlobalLocalFusion(Cn).to(DEV)
    randomize_affine_(fus.global_attn, 3)
    randomize_affine_(fus.local_attn, 4)
    keys = ["1", "3"]
    g = torch.Generator().manual_seed(9)
    drift = torch.cumsum(torch.randn(B, 1, 1, 1, generator=g) * 0.3, 0)
    f4 = {k: (torch.randn(B, Cn, h, w, generator=g) + drift).to(DEV).to(torch.bfloat16) for k in keys}
    cl = {k: torch.randn(B, 3, h, w, generator=g).to(DEV) for k in keys}
    ct = {k: torch.randn(B, 1, h, w, generator=g).to(DEV) for k in keys}
    grads = []
    for fused in (True, False):
        for p in fus.parameters():
            p.grad = None
        leaves = {k: f4[k].clone().requires_grad_(True) for k in keys}
        _, glob, _ = fus.forward_parts(leaves, cl, ct, need_local=False)
        loss = 0
        for k in keys:
            if fused:
                loss = loss + cycle.dense_seg_cycle(cycle.spatial_sum(glob[k]), 16, 2, 3, 10.0)
            else:
                feat = glob[k].float().sum(dim=(2, 3))
                lo, go = CO.dense_seg_cycle(feat.detach().cpu().numpy(), 16, 2, 3, 10.0)
                loss = loss + (feat * torch.from_numpy(go).to(DEV).float()).sum()      # same gradient, torch plumbing
        loss.backward()
        grads.append([leaves[k].grad.float() for k in keys] + [p.grad.clone() for p in fus.global_attn.parameters() if p.grad is not None])
    assert len(grads[0]) == len(grads[1]) and len(grads[0]) > len(keys)
    for a, b in zip(grads[0], grads[1]):
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-2 * scale + 1e-12


def test_install_trainer_replaces_the_two_methods_with_the_same_call_shape():
    """glfusion_b200.install_trainer(main.Trainer): the call sites of R/main.py:231 and :234 work unchanged."""
    import glfusion_b200

    class Trainer:                       # stands in for main.Trainer (the script cannot be imported off the build box)
        device = DEV

    glfusion_b200.install_trainer(Trainer)
    d = np.load(GOLDEN[0])               # dense fixture
    assert int(d["dense"]) == 1
    feat = torch.from_numpy(d["feat"]).to(DEV).requires_grad_(True)
    t = Trainer()
    loss = t.dense_seg_cycle(feat, target_region=int(d["target_region"]), cyc_off=int(d["cyc_off"]),
                             chunk_size=int(d["chunk_size"]), temperature=float(d["temperature"]),
                             soft_label=bool(d["soft_label"]), is_overlap=bool(d["is_overlap"]))
    assert abs(loss.item() - float(d["loss"])) <= TOL * abs(float(d["loss"]))
    s = np.load([g for g in GOLDEN if "single" in g][0])
    np.random.seed(int(s["np_seed"]))
    loss = t.seg_cycle(torch.from_numpy(s["feat"]).to(DEV), target_region=16, cyc_off=2, chunk_size=3, temperature=10)
    assert abs(loss.item() - float(s["loss"])) <= TOL * abs(float(s["loss"]))


def test_cycle_loss_odd_channel_count_and_bf16_features():
    """C = 100 (not a multiple of the warp width) and bf16 features: the loss is computed in fp32 from the rounded values
    and the gradient comes back in the features' dtype."""
    T, Cn, R, off, ch = 30, 100, 12, 1, 2
    g = torch.Generator().manual_seed(77)
    feat = torch.cumsum(torch.randn(T, Cn, generator=g) * 0.4, 0)
    x = feat.to(DEV).requires_grad_(True)
    loss = cycle.dense_seg_cycle(x, R, off, ch, 5.0)
    loss.backward()
    lo, go = CO.dense_seg_cycle(feat.numpy(), R, off, ch, 5.0)
    assert abs(loss.item() - lo) <= TOL * abs(lo)
    assert np.abs(x.grad.cpu().numpy() - go).max() <= TOL * np.abs(go).max()
    xb = feat.to(DEV).to(torch.bfloat16).requires_grad_(True)
    lb = cycle.dense_seg_cycle(xb, R, off, ch, 5.0)
    lb.backward()
    lo2, go2 = CO.dense_seg_cycle(xb.detach().float().cpu().numpy(), R, off, ch, 5.0)
    assert xb.grad.dtype == torch.bfloat16
    assert abs(lb.item() - lo2) <= TOL * abs(lo2)
    assert np.abs(xb.grad.float().cpu().numpy() - go2).max() <= 1e-2 * np.abs(go2).max()
