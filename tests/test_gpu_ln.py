"""BN-affine + residual + LayerNorm epilogues (R/models/ours.py:908-915 after the W_z GEMM) as standalone kernels:
single-block and fused MGFM+MLFM pair forms, forward and backward, against an fp32 PyTorch restatement of the same
arithmetic (floating-point kernel: tolerance = bf16 storage rounding, 1e-2 relative L2; statistics 1e-4)."""
import ctypes as C

import pytest
import torch

from gpu_util import DEV, stream
from glfusion_b200 import _lib as L
from oracle import tpavi_oracle as O

pytestmark = pytest.mark.gpu


def _mk(rows, Cc, seed, nmod):
    g = torch.Generator().manual_seed(seed)
    mods = []
    for _ in range(nmod):
        d = {
            "U": (torch.randn(rows, Cc, generator=g) * 1.5 + 0.3).to(torch.bfloat16),
            "X": torch.randn(rows, Cc, generator=g).to(torch.bfloat16),
            "a": torch.rand(Cc, generator=g) + 0.5, "b": torch.randn(Cc, generator=g) * 0.2,
            "lw": torch.rand(Cc, generator=g) + 0.5, "lb": torch.randn(Cc, generator=g) * 0.2,
            "mean": torch.randn(Cc, generator=g) * 0.3, "rstd": torch.rand(Cc, generator=g) + 0.5,
        }
        mods.append({k: v.to(DEV) for k, v in d.items()})
    dz = torch.randn(rows, Cc, generator=g).to(torch.bfloat16).to(DEV)
    z0 = torch.randn(rows, Cc, generator=g).to(torch.bfloat16).to(DEV)
    return mods, dz, z0


def _ref(mods, dz, eps=1e-5):
    """fp32 restatement: returns Z, per-module (mu, r, dV, d ln_w, d ln_b, d gamma, d beta)."""
    Z = 0
    outs = []
    for m in mods:
        u, x = m["U"].float(), m["X"].float()
        v = (m["a"] * u + m["b"] + x).requires_grad_(True)
        mu = v.mean(1, keepdim=True)
        var = ((v - mu) ** 2).mean(1, keepdim=True)
        r = torch.rsqrt(var + eps)
        xh = (v - mu) * r
        z = xh * m["lw"] + m["lb"]
        (dv,) = torch.autograd.grad(z, v, dz.float())
        uh = (u - m["mean"]) * m["rstd"]
        outs.append((mu.detach().squeeze(1), r.detach().squeeze(1), dv, (dz.float() * xh.detach()).sum(0),
                     dz.float().sum(0), (dv * uh).sum(0), dv.sum(0)))
        Z = Z + z.detach()
    return Z, outs


def _ptrs(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


@pytest.mark.parametrize("rows,Cc", [(1000, 256), (61, 128), (5, 256), (4099, 64)])
@pytest.mark.parametrize("accumulate", [0, 1])
def test_pair_forward(rows, Cc, accumulate):
    lib = L.load()
    mods, dz, z0 = _mk(rows, Cc, 7 + rows, 2)
    Zr, outs = _ref(mods, dz)
    Z = z0.clone()
    mu = [torch.empty(rows, device=DEV) for _ in range(2)]
    r = [torch.empty(rows, device=DEV) for _ in range(2)]
    L.check(lib.glf_bn_res_ln_pair_fwd(rows, Cc, _ptrs([m["U"] for m in mods]), _ptrs([m["X"] for m in mods]),
                                       _ptrs([m["a"] for m in mods]), _ptrs([m["b"] for m in mods]),
                                       _ptrs([m["lw"] for m in mods]), _ptrs([m["lb"] for m in mods]), L.ptr(Z),
                                       _ptrs(mu), _ptrs(r), 1e-5, accumulate, stream()))
    torch.cuda.synchronize()
    want = Zr + (z0.float() if accumulate else 0)
    assert O.rel_err(Z.float().cpu(), want.cpu()) < 1e-2
    for k in range(2):
        assert O.rel_err(mu[k].cpu(), outs[k][0].cpu()) < 1e-4
        assert O.rel_err(r[k].cpu(), outs[k][1].cpu()) < 1e-4


@pytest.mark.parametrize("rows,Cc", [(1000, 256), (61, 128), (5, 256), (4099, 64), (28 * 200, 256)])
@pytest.mark.parametrize("nmod", [1, 2])
def test_backward_single_and_pair(rows, Cc, nmod):
    lib = L.load()
    mods, dz, _ = _mk(rows, Cc, 11 + rows, nmod)
    _, outs = _ref(mods, dz)
    nbmax = lib.glf_bn_res_ln_bwd_max_blocks()
    dV = [torch.empty(rows, Cc, device=DEV, dtype=torch.bfloat16) for _ in range(nmod)]
    part = [torch.zeros(nbmax * 4 * Cc, device=DEV) for _ in range(nmod)]
    mu = [o[0].contiguous() for o in outs]
    r = [o[1].contiguous() for o in outs]
    nb = C.c_int(0)
    if nmod == 2:
        L.check(lib.glf_bn_res_ln_pair_bwd(rows, Cc, L.ptr(dz), _ptrs([m["U"] for m in mods]),
                                           _ptrs([m["X"] for m in mods]), _ptrs([m["a"] for m in mods]),
                                           _ptrs([m["b"] for m in mods]), _ptrs([m["mean"] for m in mods]),
                                           _ptrs([m["rstd"] for m in mods]), _ptrs([m["lw"] for m in mods]),
                                           _ptrs(mu), _ptrs(r), _ptrs(dV), _ptrs(part), C.byref(nb), stream()))
    else:
        m = mods[0]
        L.check(lib.glf_bn_res_ln_bwd(rows, Cc, L.ptr(dz), L.DTYPE_BF16, L.ptr(m["U"]), L.ptr(m["X"]), L.ptr(m["a"]),
                                      L.ptr(m["b"]), L.ptr(m["mean"]), L.ptr(m["rstd"]), L.ptr(m["lw"]), L.ptr(mu[0]),
                                      L.ptr(r[0]), L.ptr(dV[0]), L.ptr(part[0]), C.byref(nb), stream()))
    torch.cuda.synchronize()
    assert 0 < nb.value <= nbmax
    for k in range(nmod):
        assert O.rel_err(dV[k].float().cpu(), outs[k][2].cpu()) < 1e-2
        red = part[k][: nb.value * 4 * Cc].view(nb.value, 4, Cc).double().sum(0).float().cpu()
        for j, name in enumerate(("d_ln_w", "d_ln_b", "d_gamma", "d_beta")):
            assert O.rel_err(red[j], outs[k][3 + j].cpu()) < 2e-3, name


@pytest.mark.parametrize("rows,Cc", [(1000, 256), (333, 128)])
def test_forward_single_matches_pair_halves(rows, Cc):
    """Z_pair == Z_single(g) then accumulate Z_single(l): the fused pass changes bytes moved, not results."""
    lib = L.load()
    mods, dz, _ = _mk(rows, Cc, 23, 2)
    Zp = torch.empty(rows, Cc, device=DEV, dtype=torch.bfloat16)
    Zs = torch.empty_like(Zp)
    mu = [torch.empty(rows, device=DEV) for _ in range(4)]
    L.check(lib.glf_bn_res_ln_pair_fwd(rows, Cc, _ptrs([m["U"] for m in mods]), _ptrs([m["X"] for m in mods]),
                                       _ptrs([m["a"] for m in mods]), _ptrs([m["b"] for m in mods]),
                                       _ptrs([m["lw"] for m in mods]), _ptrs([m["lb"] for m in mods]), L.ptr(Zp),
                                       _ptrs(mu[:2]), _ptrs(mu[2:]), 1e-5, 0, stream()))
    for k, m in enumerate(mods):
        L.check(lib.glf_bn_res_ln_fwd(rows, Cc, L.ptr(m["U"]), L.ptr(m["X"]), L.ptr(m["a"]), L.ptr(m["b"]),
                                      L.ptr(m["lw"]), L.ptr(m["lb"]), L.ptr(Zs), L.DTYPE_BF16, L.ptr(mu[0]),
                                      L.ptr(mu[1]), 1e-5, k, stream()))
    torch.cuda.synchronize()
    # the two-step form rounds the first block's output to bf16 before adding the second: one extra rounding
    assert O.rel_err(Zs.float().cpu(), Zp.float().cpu()) < 6e-3
