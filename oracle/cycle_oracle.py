"""CPU oracle for the cycle-consistency step that consumes the fusion path's MGFM output.

TEST INFRASTRUCTURE ONLY.  Nothing under ``glfusion_b200/`` may import this file.

Restates (R = /root/reference/GLfusion):

* the per-view spatial sum                 R/main.py:229      ``cyc_feat_out[view].sum(dim=(2, 3))``
* ``Trainer.seg_cycle``                    R/main.py:650-717
* ``Trainer.dense_seg_cycle``              R/main.py:719-798

The reference writes the loss with ``repeat`` / ``gather`` index gymnastics whose modulo wrap-around is cut off again
by the slices that follow (rows ``k < nk - chunk - off + 1`` never reach index ``nk``), so every gathered element is
a plain shifted read.  With ``feat`` the [T, C] per-frame features, R = target_region, off = cyc_off, ch =
chunk_size, K = feat[R:], nk = T - R, a = temperature / (C * ch) and a start position s:

    q_j      = feat[s + j]                                            j < ch            (main.py:660)
    sim_i    = -a * sum_j |K[i + j] - q_j|^2                          i < nk-ch-off+1   (main.py:666-679)
    beta     = softmax(sim)                                                              (main.py:680)
    w_j      = sum_i beta_i * K[off + i + j]                                             (main.py:685-693)
    z_m      = -a * sum_j |feat[off + m + j] - w_j|^2                 m < R-off-ch+1    (main.py:697-711)
    loss_s   = mean_m BCEWithLogits(z_m, [m == s])                                       (main.py:717)

``seg_cycle`` draws s with ``np.random.choice`` (main.py:655); ``dense_seg_cycle`` averages loss_s over every s
(step 1, or ``chunk_size`` without overlap) and divides by the number of positions R-ch-off+1 (main.py:798), with
optional soft labels 0.8 / 0.2/(L-1) (main.py:790).

PARITY PIN: ``oracle/gen_golden_cycle.py`` compiles the two reference methods from where they lie in R/main.py (the
enclosing script imports nibabel / monai / tensorboardX and cannot be imported here), runs them with autograd on seeded
inputs and commits ``tests/golden/cycle_*.npz``; ``tests/test_oracle_cycle.py`` pins this file to those vectors.
Everything here is numpy fp64 with a hand-derived backward.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def spatial_sum(f: np.ndarray) -> np.ndarray:
    """[B, C, h, w] -> [B, C] (R/main.py:229)."""
    return np.asarray(f, dtype=np.float64).sum(axis=(2, 3))


def positions(target_region: int, cyc_off: int, chunk_size: int) -> int:
    """Number of start positions / logits: R - (ch + off) + 1 (main.py:655)."""
    return target_region - (chunk_size + cyc_off) + 1


def _one_start(feat: np.ndarray, R: int, off: int, ch: int, temperature: float, s: int, target: np.ndarray,
               ) -> Tuple[float, np.ndarray]:
    """loss_s and d loss_s / d feat for one start position."""
    T, C = feat.shape
    K = feat[R:]
    nk = T - R
    Lk = nk - ch - off + 1
    Lq = R - off - ch + 1
    if Lk < 1 or Lq < 1:
        raise ValueError("cycle loss: not enough frames for target_region / cyc_off / chunk_size")
    a = temperature / (C * ch)
    q = feat[s:s + ch]                                                          # [ch, C]
    # sim_i = -a sum_j |K[i+j] - q_j|^2
    dK1 = np.stack([K[j:j + Lk] - q[j] for j in range(ch)], axis=1)             # [Lk, ch, C]
    sim = -a * (dK1 ** 2).sum(axis=(1, 2))
    e = np.exp(sim - sim.max())
    beta = e / e.sum()
    Kb = np.stack([K[off + j:off + j + Lk] for j in range(ch)], axis=1)         # [Lk, ch, C]
    w = (beta[:, None, None] * Kb).sum(axis=0)                                  # [ch, C]
    Qc = feat[off:R]
    dQ = np.stack([Qc[j:j + Lq] - w[j] for j in range(ch)], axis=1)             # [Lq, ch, C]
    z = -a * (dQ ** 2).sum(axis=(1, 2))
    # BCE with logits, mean over the Lq logits
    loss = float(np.mean(np.maximum(z, 0) - z * target + np.log1p(np.exp(-np.abs(z)))))
    # ---- backward
    sig = np.where(z >= 0, 1.0 / (1.0 + np.exp(-np.abs(z))), np.exp(-np.abs(z)) / (1.0 + np.exp(-np.abs(z))))
    dz = (sig - target) / Lq
    g = np.zeros_like(feat)
    dw = np.zeros_like(w)
    for j in range(ch):
        t = 2 * a * dz[:, None] * dQ[:, j]                                      # d z_m / d w_j = +2a (x - w_j)
        dw[j] = t.sum(axis=0)
        g[off + j:off + j + Lq] -= t
    dbeta = (Kb * dw[None]).sum(axis=(1, 2))
    for j in range(ch):
        g[R + off + j:R + off + j + Lk] += beta[:, None] * dw[j]
    dsim = beta * (dbeta - (beta * dbeta).sum())
    for j in range(ch):
        t = 2 * a * dsim[:, None] * dK1[:, j]
        g[R + j:R + j + Lk] -= t
        g[s + j] += t.sum(axis=0)
    return loss, g


def one_hot(L: int, s: int, soft_label: bool = False) -> np.ndarray:
    y = np.zeros(L)
    y[s] = 1.0
    if soft_label:
        y = np.where(y == 1, 0.8, 0.2 / (L - 1))                                # main.py:790
    return y


def seg_cycle(feat, target_region: int, cyc_off: int, chunk_size: int, temperature: float, target_strtpt: int,
              ) -> Tuple[float, np.ndarray]:
    """Trainer.seg_cycle (main.py:650-717) for a given draw of the start position; returns (loss, d loss / d feat)."""
    feat = np.asarray(feat, dtype=np.float64)
    L = positions(target_region, cyc_off, chunk_size)
    return _one_start(feat, target_region, cyc_off, chunk_size, temperature, int(target_strtpt),
                      one_hot(L, int(target_strtpt)))


def dense_starts(target_region: int, cyc_off: int, chunk_size: int, is_overlap: bool = True) -> Sequence[int]:
    return range(0, positions(target_region, cyc_off, chunk_size), 1 if is_overlap else chunk_size)


def dense_seg_cycle(feat, target_region: int, cyc_off: int, chunk_size: int, temperature: float,
                    soft_label: bool = False, is_overlap: bool = True) -> Tuple[float, np.ndarray]:
    """Trainer.dense_seg_cycle (main.py:719-798); returns (loss, d loss / d feat)."""
    feat = np.asarray(feat, dtype=np.float64)
    L = positions(target_region, cyc_off, chunk_size)
    loss, g = 0.0, np.zeros_like(feat)
    for s in dense_starts(target_region, cyc_off, chunk_size, is_overlap):
        l_s, g_s = _one_start(feat, target_region, cyc_off, chunk_size, temperature, s, one_hot(L, s, soft_label))
        loss += l_s
        g += g_s
    return loss / L, g / L
