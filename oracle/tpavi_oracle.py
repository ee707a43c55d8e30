"""CPU oracle for the GL-Fusion global/local cross-view fusion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``glfusion_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker / the CPU arm, never as the product path.

It restates, op for op, the reference algorithm (R = /root/reference/GLfusion):

* ``TPAVIModule.__init__``           R/models/ours.py:771-843   (dup R/models/TPAVI.py:7-83)
* ``TPAVIModule.forward``            R/models/ours.py:845-917   (dup R/models/TPAVI.py:86-155)
* call-site glue of ``Global_and_Local.forward``  R/models/ours.py:1802-1837

The arithmetic lives in PyTorch (a third-party library that IS installed here, reference pin
``torch==1.8.1+cu111`` R/requirements.txt:23, this image torch 2.11), so the restatement is written in
elementary tensor ops (matmul / mean / rsqrt ...) rather than by calling nn.Conv3d / nn.BatchNorm3d /
nn.LayerNorm, and is pinned against the *real* reference module:

* PARITY PIN: the reference ships no golden vectors or tests.  The pin is
  ``oracle/gen_golden.py``: it imports the unmodified reference ``models/TPAVI.py`` in the build container,
  runs it on seeded inputs and commits inputs+outputs under ``tests/golden/*.npz``.
  ``tests/test_oracle.py`` checks this file against those vectors (and against the live reference when
  ``/root/reference`` exists).  So parity is pinned on outputs of the reference itself run here.

Three restatements are provided:

1. ``tpavi_forward``                — literal order of operations (materialises the N x N matrix), autograd gives
                                     the backward exactly as the reference's autograd would.
2. ``tpavi_dot_closed_form``        — the reassociated ``mode='dot'`` algorithm  y = Theta (Phi^T G) / N with a
                                     hand-derived backward, chunkable, usable in fp64 at
                                     sequence lengths where N x N cannot be materialised.
3. ``tpavi_dot_gram_form``          — the second exact reassociation (Gram matrix S = X~^T X~ per sequence, every
                                     other product in channel space) that libglf_sm100a runs when N >= 5 C, with its
                                     hand-derived backward; pinned to the same golden vectors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

BN_EPS = 1e-5       # nn.BatchNorm3d default, R/models/ours.py:824
BN_MOMENTUM = 0.1   # nn.BatchNorm3d default
LN_EPS = 1e-5       # nn.LayerNorm default, R/models/ours.py:802

PARAM_KEYS = (
    "align_channel.weight", "align_channel.bias",
    "norm_layer.weight", "norm_layer.bias",
    "g.weight", "g.bias",
    "W_z.0.weight", "W_z.0.bias",
    "W_z.1.weight", "W_z.1.bias",
    "theta.weight", "theta.bias",
    "phi.weight", "phi.bias",
)
BUFFER_KEYS = ("W_z.1.running_mean", "W_z.1.running_var", "W_z.1.num_batches_tracked")


# --------------------------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------------------------
def _conv_init(out_c: int, in_c: int, gen: torch.Generator, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn.Conv3d default init (kaiming_uniform(a=sqrt(5)) -> U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both)."""
    bound = 1.0 / math.sqrt(in_c)
    w = (torch.rand(out_c, in_c, 1, 1, 1, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    b = (torch.rand(out_c, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    return w.to(dtype), b.to(dtype)


def init_params(in_channels: int, inter_channels: Optional[int] = None, seed: int = 0,
                randomize_affine: bool = True, bn_layer: bool = True,
                dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """State-dict-shaped parameters + buffers of one TPAVIModule (R/models/ours.py:771-843).

    ``randomize_affine`` re-draws BN gamma/beta and the LayerNorm affine: the reference zero-initialises BN
    gamma/beta (ours.py:826-827) which makes the whole attention branch output exactly 0 — parity on such
    weights is vacuous.
    """
    C = in_channels
    Ci = inter_channels if inter_channels is not None else max(C // 2, 1)
    gen = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    bound = 1.0 / math.sqrt(128)
    p["align_channel.weight"] = ((torch.rand(C, 128, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    p["align_channel.bias"] = ((torch.rand(C, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    p["norm_layer.weight"] = torch.ones(C, dtype=dtype)
    p["norm_layer.bias"] = torch.zeros(C, dtype=dtype)
    p["g.weight"], p["g.bias"] = _conv_init(Ci, C, gen, dtype)
    if bn_layer:
        p["W_z.0.weight"], p["W_z.0.bias"] = _conv_init(C, Ci, gen, dtype)
        p["W_z.1.weight"] = torch.zeros(C, dtype=dtype)
        p["W_z.1.bias"] = torch.zeros(C, dtype=dtype)
        p["W_z.1.running_mean"] = torch.zeros(C, dtype=dtype)
        p["W_z.1.running_var"] = torch.ones(C, dtype=dtype)
        p["W_z.1.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    else:
        p["W_z.weight"] = torch.zeros(C, Ci, 1, 1, 1, dtype=dtype)
        p["W_z.bias"] = torch.zeros(C, dtype=dtype)
    p["theta.weight"], p["theta.bias"] = _conv_init(Ci, C, gen, dtype)
    p["phi.weight"], p["phi.bias"] = _conv_init(Ci, C, gen, dtype)
    if randomize_affine:
        def nrm(mu, sd):
            return (torch.randn(C, generator=gen, dtype=torch.float64) * sd + mu).to(dtype)
        p["norm_layer.weight"] = nrm(1.0, 0.2)
        p["norm_layer.bias"] = nrm(0.0, 0.2)
        if bn_layer:
            p["W_z.1.weight"] = nrm(1.0, 0.2)
            p["W_z.1.bias"] = nrm(0.0, 0.2)
        else:
            p["W_z.weight"], p["W_z.bias"] = _conv_init(C, Ci, gen, dtype)
    return p


def _w2d(w: torch.Tensor) -> torch.Tensor:
    return w.reshape(w.shape[0], w.shape[1])


# --------------------------------------------------------------------------------------------------------------
# literal forward (R/models/ours.py:845-917)
# --------------------------------------------------------------------------------------------------------------
def tpavi_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], mode: str = "dot", training: bool = True,
                  bn_layer: bool = True, update_buffers: bool = True) -> torch.Tensor:
    """z = TPAVIModule(x)[0] for x [B,C,T,H,W], ``audio=None``.  Returns z as [B,C,T,H,W] (logical layout).

    Line references are to R/models/ours.py.  Buffers in ``p`` are updated in place when ``training``.
    """
    if mode not in ("dot", "embedded", "gaussian"):
        raise ValueError("oracle supports modes dot / embedded / gaussian (concatenate is O(N^2 C) memory "
                         "and unused by every constructor site in the reference)")
    B, C = x.shape[0], x.shape[1]
    xt = x.reshape(B, C, -1)                                   # [B,C,N]
    N = xt.shape[-1]
    # :866  g_x = self.g(x).view(B,C',-1).permute(0,2,1)          (1x1x1 conv == per-token matmul)
    g_x = (torch.matmul(_w2d(p["g.weight"]), xt) + p["g.bias"][None, :, None]).permute(0, 2, 1)   # [B,N,C']
    if mode == "gaussian":
        # :869-873
        theta_x = xt.permute(0, 2, 1)
        phi_x = xt
        # NOTE: with mode='gaussian' the reference applies W_z to y of width C (not C'); it only runs when
        # inter_channels == in_channels.  Kept for completeness of the restatement.
    else:
        # :878-880
        theta_x = (torch.matmul(_w2d(p["theta.weight"]), xt) + p["theta.bias"][None, :, None]).permute(0, 2, 1)
        phi_x = torch.matmul(_w2d(p["phi.weight"]), xt) + p["phi.bias"][None, :, None]             # [B,C',N]
    f = torch.matmul(theta_x, phi_x)                           # :881  [B,N,N]
    if mode in ("gaussian", "embedded"):
        f_div_C = torch.softmax(f, dim=-1)                     # :896-897
    else:
        f_div_C = f / N                                        # :899-900
    y = torch.matmul(f_div_C, g_x)                             # :902  [B,N,C']
    y = y.permute(0, 2, 1)                                     # :905  [B,C',N]
    # :908  W_z = Conv3d(1x1x1) (+ BatchNorm3d)
    if bn_layer:
        u = torch.matmul(_w2d(p["W_z.0.weight"]), y) + p["W_z.0.bias"][None, :, None]             # [B,C,N]
        if training:
            mean = u.mean(dim=(0, 2))
            var = u.var(dim=(0, 2), unbiased=False)
            if update_buffers:
                with torch.no_grad():
                    n = u.shape[0] * u.shape[2]
                    p["W_z.1.running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach().to(p["W_z.1.running_mean"].dtype))
                    p["W_z.1.running_var"].mul_(1 - BN_MOMENTUM).add_(
                        BN_MOMENTUM * (var.detach() * n / max(n - 1, 1)).to(p["W_z.1.running_var"].dtype))
                    p["W_z.1.num_batches_tracked"] += 1
        else:
            mean = p["W_z.1.running_mean"].to(u.dtype)
            var = p["W_z.1.running_var"].to(u.dtype)
        w_y = (u - mean[None, :, None]) * torch.rsqrt(var + BN_EPS)[None, :, None] \
            * p["W_z.1.weight"][None, :, None] + p["W_z.1.bias"][None, :, None]
    else:
        w_y = torch.matmul(_w2d(p["W_z.weight"]), y) + p["W_z.bias"][None, :, None]
    z = w_y + xt                                               # :910
    # :913-915  LayerNorm over C (biased variance)
    zt = z.permute(0, 2, 1)                                    # [B,N,C]
    mu = zt.mean(dim=-1, keepdim=True)
    va = zt.var(dim=-1, unbiased=False, keepdim=True)
    zt = (zt - mu) * torch.rsqrt(va + LN_EPS) * p["norm_layer.weight"] + p["norm_layer.bias"]
    return zt.permute(0, 2, 1).reshape(x.shape)


def tpavi_fwd_bwd(x: torch.Tensor, dz: torch.Tensor, p: Dict[str, torch.Tensor], mode: str = "dot",
                  training: bool = True, bn_layer: bool = True):
    """Forward + autograd backward.  Returns (z, dx, grads{name: tensor}).  Buffers in p are updated."""
    x = x.detach().clone().requires_grad_(True)
    names = [k for k in p if k not in BUFFER_KEYS and not k.startswith("align_channel")]
    if mode == "gaussian":
        names = [k for k in names if not (k.startswith("theta") or k.startswith("phi"))]
    leaves = {k: p[k].detach().clone().requires_grad_(True) for k in names}
    q = dict(p)
    q.update(leaves)
    z = tpavi_forward(x, q, mode=mode, training=training, bn_layer=bn_layer)
    for k in BUFFER_KEYS:
        if k in q:
            p[k] = q[k]
    z.backward(dz)
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return z.detach(), x.grad.detach(), grads


# --------------------------------------------------------------------------------------------------------------
# closed form for mode='dot' (reassociated, O(N C'^2)); SURVEY.md §8a row 10
# --------------------------------------------------------------------------------------------------------------
def tpavi_dot_closed_form(x: torch.Tensor, dz: Optional[torch.Tensor], p: Dict[str, torch.Tensor],
                          training: bool = True):
    """mode='dot', bn_layer=True, token-major math.  x, dz: [B,C,T,H,W].

    forward :  P=[Theta|Phi|G] = X W^T + b ;  M = Phi^T G / N ;  Y = Theta M ;  U = Y Wz^T + bz ;
               V = BN(U) ;  Z = LN(V + X)
    backward:  LN-bwd -> dV (=dX partial) -> BN-bwd -> dU -> dWz, dY -> dTheta, dM -> dPhi, dG -> dW*, dX.
    Returns (z, dx, grads, aux) ; dx/grads are None when dz is None.
    """
    B, C = x.shape[0], x.shape[1]
    X = x.reshape(B, C, -1).permute(0, 2, 1)                   # [B,N,C]
    N = X.shape[1]
    Wt, Wp, Wg = _w2d(p["theta.weight"]), _w2d(p["phi.weight"]), _w2d(p["g.weight"])
    Wz, bz = _w2d(p["W_z.0.weight"]), p["W_z.0.bias"]
    gam, bet = p["W_z.1.weight"], p["W_z.1.bias"]
    lw, lb = p["norm_layer.weight"], p["norm_layer.bias"]
    Th = X @ Wt.T + p["theta.bias"]
    Ph = X @ Wp.T + p["phi.bias"]
    G = X @ Wg.T + p["g.bias"]
    M = Ph.transpose(1, 2) @ G / N                             # [B,C',C']
    Y = Th @ M
    U = Y @ Wz.T + bz                                          # [B,N,C]
    if training:
        mean = U.mean(dim=(0, 1))
        var = U.var(dim=(0, 1), unbiased=False)
    else:
        mean, var = p["W_z.1.running_mean"].to(U.dtype), p["W_z.1.running_var"].to(U.dtype)
    rstd = torch.rsqrt(var + BN_EPS)
    Uh = (U - mean) * rstd
    V = Uh * gam + bet
    Zp = V + X
    mu = Zp.mean(-1, keepdim=True)
    r = torch.rsqrt(Zp.var(-1, unbiased=False, keepdim=True) + LN_EPS)
    Xh = (Zp - mu) * r
    Z = Xh * lw + lb
    z = Z.permute(0, 2, 1).reshape(x.shape)
    aux = {"mean": mean, "var": var, "M": M}
    if dz is None:
        return z, None, None, aux
    dZ = dz.reshape(B, C, -1).permute(0, 2, 1)
    grads: Dict[str, torch.Tensor] = {}
    grads["norm_layer.weight"] = (dZ * Xh).sum((0, 1))
    grads["norm_layer.bias"] = dZ.sum((0, 1))
    dXh = dZ * lw
    dZp = r * (dXh - dXh.mean(-1, keepdim=True) - Xh * (dXh * Xh).mean(-1, keepdim=True))
    dX = dZp.clone()
    dV = dZp
    grads["W_z.1.weight"] = (dV * Uh).sum((0, 1))
    grads["W_z.1.bias"] = dV.sum((0, 1))
    if training:
        n = B * N
        dU = gam * rstd * (dV - dV.sum((0, 1)) / n - Uh * (dV * Uh).sum((0, 1)) / n)
    else:
        dU = gam * rstd * dV
    grads["W_z.0.weight"] = torch.einsum("bnc,bnk->ck", dU, Y).reshape(p["W_z.0.weight"].shape)
    grads["W_z.0.bias"] = dU.sum((0, 1))
    dY = dU @ Wz
    dTh = dY @ M.transpose(1, 2)
    dM = Th.transpose(1, 2) @ dY                               # [B,C',C']
    dPh = G @ dM.transpose(1, 2) / N
    dG = Ph @ dM / N
    for nm, d in (("theta", dTh), ("phi", dPh), ("g", dG)):
        grads[nm + ".weight"] = torch.einsum("bnk,bnc->kc", d, X).reshape(p[nm + ".weight"].shape)
        grads[nm + ".bias"] = d.sum((0, 1))
    dX = dX + dTh @ Wt + dPh @ Wp + dG @ Wg
    dx = dX.permute(0, 2, 1).reshape(x.shape)
    return z, dx, grads, aux


# --------------------------------------------------------------------------------------------------------------
# Gram form for mode='dot' (second exact reassociation, O(N C^2) with every per-token product in channel space)
# --------------------------------------------------------------------------------------------------------------
def tpavi_dot_gram_form(x: torch.Tensor, dz: Optional[torch.Tensor], p: Dict[str, torch.Tensor],
                        training: bool = True, bn_layer: bool = True):
    """mode='dot' with theta / phi / g never materialised per token (what libglf_sm100a runs when N >> C).

    With the homogeneous token  x~ = [x, 1]  and  W~ = [W | b]  (R/models/ours.py:866,878-879 are affine maps):
        S~_b = X~_b^T X~_b                       Gram matrix of a sequence (the only token contraction, ours.py:881,902)
        M_b  = W~phi S~_b W~g^T / N              == Phi_b^T G_b / N
        W'_b = Wz M_b^T ;  Q~_b = W'_b W~theta   [C, C+1]
        U_b  = X~_b Q~_b^T                       == W_z(y) without its bias (ours.py:908)
    backward (dU = k1*dV + k2*U + k3, the BatchNorm backward, is never materialised either):
        dQ~_b = k1*(dV_b^T X~_b) + k2*(Q~_b S~_b) + k3 (1^T X~_b)
        ... small per-sequence products ...
        dX_b  = dV_b (k1*Q_b) + X_b (Q_b^T k2 Q_b + dS + dS^T) + 1 e_b^T + dV_b
    Returns (z, dx, grads, aux) like tpavi_dot_closed_form.
    """
    B, C = x.shape[0], x.shape[1]
    X = x.reshape(B, C, -1).permute(0, 2, 1)                   # [B,N,C]
    N = X.shape[1]
    cat = lambda w, b: torch.cat([_w2d(w), b[:, None]], dim=1)
    Wta = cat(p["theta.weight"], p["theta.bias"])              # [C',C+1]
    Wpa = cat(p["phi.weight"], p["phi.bias"])
    Wga = cat(p["g.weight"], p["g.bias"])
    if bn_layer:
        Wz, bz = _w2d(p["W_z.0.weight"]), p["W_z.0.bias"]
    else:
        Wz, bz = _w2d(p["W_z.weight"]), p["W_z.bias"]
    lw, lb = p["norm_layer.weight"], p["norm_layer.bias"]
    S = X.transpose(1, 2) @ X                                  # [B,C,C]
    s = X.sum(1)                                               # [B,C]
    Sa = torch.cat([torch.cat([S, s[:, :, None]], 2),
                    torch.cat([s, torch.full((B, 1), float(N), dtype=x.dtype)], 1)[:, None, :]], 1)   # [B,C+1,C+1]
    T = Wpa @ Sa                                               # [B,C',C+1]  (= Phi^T X~)
    M = T @ Wga.T / N                                          # [B,C',C']
    Wpr = Wz @ M.transpose(1, 2)                               # [B,C,C']
    Qa = Wpr @ Wta                                             # [B,C,C+1]
    Q, c = Qa[:, :, :C], Qa[:, :, C]
    U = X @ Q.transpose(1, 2) + c[:, None, :]                  # without bz
    if bn_layer:
        if training:
            mean = U.mean(dim=(0, 1))
            var = U.var(dim=(0, 1), unbiased=False)
            mean_true = mean + bz                              # what the reference's running_mean sees
        else:
            mean_true, var = p["W_z.1.running_mean"].to(U.dtype), p["W_z.1.running_var"].to(U.dtype)
            mean = mean_true - bz
        rstd = torch.rsqrt(var + BN_EPS)
        gam, bet = p["W_z.1.weight"], p["W_z.1.bias"]
        Uh = (U - mean) * rstd
        V = Uh * gam + bet
    else:
        V = U + bz
        mean_true = var = None
    Zp = V + X
    mu = Zp.mean(-1, keepdim=True)
    r = torch.rsqrt(Zp.var(-1, unbiased=False, keepdim=True) + LN_EPS)
    Xh = (Zp - mu) * r
    Z = Xh * lw + lb
    z = Z.permute(0, 2, 1).reshape(x.shape)
    aux = {"mean": mean_true, "var": var, "M": M}
    if dz is None:
        return z, None, None, aux
    dZ = dz.reshape(B, C, -1).permute(0, 2, 1)
    grads: Dict[str, torch.Tensor] = {}
    grads["norm_layer.weight"] = (dZ * Xh).sum((0, 1))
    grads["norm_layer.bias"] = dZ.sum((0, 1))
    dXh = dZ * lw
    dV = r * (dXh - dXh.mean(-1, keepdim=True) - Xh * (dXh * Xh).mean(-1, keepdim=True))
    n = B * N
    if bn_layer:
        grads["W_z.1.weight"] = (dV * Uh).sum((0, 1))
        grads["W_z.1.bias"] = dV.sum((0, 1))
        k1 = gam * rstd
        if training:
            m1 = dV.sum((0, 1)) / n
            m2 = (dV * Uh).sum((0, 1)) / n
            k2 = -k1 * m2 * rstd
            k3 = -k1 * m1 + k1 * m2 * rstd * mean
        else:
            k2 = torch.zeros_like(k1)
            k3 = torch.zeros_like(k1)
    else:
        k1 = torch.ones(C, dtype=x.dtype)
        k2 = torch.zeros(C, dtype=x.dtype)
        k3 = torch.zeros(C, dtype=x.dtype)
    # column sums of dU per channel (bias gradient of W_z): sum_b (k1 rv_b + k2 colsum(U_b) + N k3)
    R = dV.transpose(1, 2) @ X                                 # [B,C,C]
    rv = dV.sum(1)                                             # [B,C]
    Ra = torch.cat([R, rv[:, :, None]], 2)                     # [B,C,C+1] = dV^T X~
    QS = Qa @ Sa                                               # [B,C,C+1] = U^T X~
    sa = Sa[:, C, :]                                           # [B,C+1]   = 1^T X~
    dQa = k1[None, :, None] * Ra + k2[None, :, None] * QS + k3[None, :, None] * sa[:, None, :]
    bz_grad = dQa[:, :, C].sum(0)                              # = sum over all tokens of dU
    grads["W_z.0.bias" if bn_layer else "W_z.bias"] = bz_grad
    dWpr = dQa @ Wta.T                                         # [B,C,C']
    dWta = (Wpr.transpose(1, 2) @ dQa).sum(0)                  # [C',C+1]
    dWz = (dWpr @ M).sum(0)                                    # [C,C']
    dM = dWpr.transpose(1, 2) @ Wz                             # [B,C',C']
    D = dM / N
    dWga = (D.transpose(1, 2) @ T).sum(0)
    dT = D @ Wga                                               # [B,C',C+1]
    dWpa = (dT @ Sa).sum(0)
    dSa = Wpa.T @ dT                                           # [B,C+1,C+1]
    G0 = dSa + 0.5 * (Qa.transpose(1, 2) @ (k2[None, :, None] * Qa))
    Sym = G0 + G0.transpose(1, 2)
    E = k1[None, :, None] * Q                                  # [B,C,C]
    F = Sym[:, :C, :C]
    e = Sym[:, :C, C] + (Q.transpose(1, 2) @ k3[None, :, None].expand(B, C, 1)).squeeze(-1)
    dX = dV @ E + X @ F + e[:, None, :] + dV
    wkey = "W_z.0.weight" if bn_layer else "W_z.weight"
    grads[wkey] = dWz.reshape(p[wkey].shape)
    for nm, d in (("theta", dWta), ("phi", dWpa), ("g", dWga)):
        grads[nm + ".weight"] = d[:, :C].reshape(p[nm + ".weight"].shape)
        grads[nm + ".bias"] = d[:, C].clone()
    dx = dX.permute(0, 2, 1).reshape(x.shape)
    return z, dx, grads, aux


# --------------------------------------------------------------------------------------------------------------
# call-site glue (R/models/ours.py:1802-1837)
# --------------------------------------------------------------------------------------------------------------
def gate_from_logits(cls_logits: torch.Tensor, ctr_logits: torch.Tensor, weight: float = 20.0) -> torch.Tensor:
    """a = sigmoid(w * max_c sigmoid(cls) * sigmoid(ctr))      ours.py:1803-1815.

    cls_logits [B,5,h,w] (output of classifier[view]), ctr_logits [B,1,h,w] -> gate [B,1,h,w].
    AdaptiveMaxPool3d((1,h,w)) over a [B,5,h,w] tensor is a max over the 5 class channels (ours.py:1805-1806).
    """
    m = torch.sigmoid(cls_logits).amax(dim=1, keepdim=True)
    c = torch.sigmoid(ctr_logits)
    return torch.sigmoid(weight * m * c)


def gate_concat(f4: Sequence[torch.Tensor], cls_logits: Sequence[torch.Tensor],
                ctr_logits: Sequence[torch.Tensor], weight: float = 20.0):
    """Per-view f4 [B,C,h,w] -> (X_global, X_local) both [B,C,V,h,w]   ours.py:1814-1820,1826-1827."""
    xg = torch.cat([f.unsqueeze(2) for f in f4], dim=2)
    xl = torch.cat([(f * gate_from_logits(cl, ct, weight)).unsqueeze(2)
                    for f, cl, ct in zip(f4, cls_logits, ctr_logits)], dim=2)
    return xg, xl


def global_local_fusion(f4: Sequence[torch.Tensor], cls_logits: Sequence[torch.Tensor],
                        ctr_logits: Sequence[torch.Tensor], p_global: Dict[str, torch.Tensor],
                        p_local: Dict[str, torch.Tensor], mode: str = "dot", training: bool = True,
                        weight: float = 20.0) -> List[torch.Tensor]:
    """f4_fusion[view] = MGFM(cat f4)[:, :, v] + MLFM(cat f4*gate)[:, :, v]     ours.py:1819-1834."""
    xg, xl = gate_concat(f4, cls_logits, ctr_logits, weight)
    zg = tpavi_forward(xg, p_global, mode=mode, training=training)
    zl = tpavi_forward(xl, p_local, mode=mode, training=training)
    return [(zg[:, :, i] + zl[:, :, i]) for i in range(len(f4))]


def fusion_fwd_bwd(f4, cls_logits, ctr_logits, d_out, p_global, p_local, mode="dot", training=True, weight=20.0):
    """Forward+autograd backward of the whole MGFM+MLFM path.  Returns (outs, df4, dcls, dctr, grads_g, grads_l)."""
    f4 = [t.detach().clone().requires_grad_(True) for t in f4]
    cls_logits = [t.detach().clone().requires_grad_(True) for t in cls_logits]
    ctr_logits = [t.detach().clone().requires_grad_(True) for t in ctr_logits]

    def leafify(p):
        names = [k for k in p if k not in BUFFER_KEYS and not k.startswith("align_channel")]
        leaves = {k: p[k].detach().clone().requires_grad_(True) for k in names}
        q = dict(p)
        q.update(leaves)
        return q, leaves
    qg, lg = leafify(p_global)
    ql, ll = leafify(p_local)
    outs = global_local_fusion(f4, cls_logits, ctr_logits, qg, ql, mode=mode, training=training, weight=weight)
    for k in BUFFER_KEYS:
        if k in qg:
            p_global[k] = qg[k]
            p_local[k] = ql[k]
    torch.autograd.backward(outs, list(d_out))
    z = lambda t: t.grad if t.grad is not None else torch.zeros_like(t)
    return ([o.detach() for o in outs], [z(t) for t in f4], [z(t) for t in cls_logits], [z(t) for t in ctr_logits],
            {k: z(v) for k, v in lg.items()}, {k: z(v) for k, v in ll.items()})


# --------------------------------------------------------------------------------------------------------------
# metric helpers
# --------------------------------------------------------------------------------------------------------------
def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 in fp64 (the 'relative error' the north_star tolerances refer to)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def dice(pred_logits: torch.Tensor, target: torch.Tensor) -> float:
    """2TP/(2TP+FP+FN+1e-5) on sigmoid>0.5     R/main.py:800-815."""
    pr = (torch.sigmoid(pred_logits.float()) > 0.5)
    gt = target > 0.5
    tp = (pr & gt).sum().item()
    fp = (pr & ~gt).sum().item()
    fn = (~pr & gt).sum().item()
    return 2 * tp / (2 * tp + fp + fn + 1e-5)
