import sys
sys.path.insert(0, '/root/repo/profiles'); sys.path.insert(0, '/root/repo')
import torch
import measure_extra as M
V, F, h, w = 4, 16, 28, 28
ms = M.ours_pair_fwd_bwd(2 * F, 2048, V, h, w)
print("C2048", round(ms, 3), "ms", round(2 * 13.5 * (2 * F * V * h * w) * 2048 * 2048 / (ms * 1e-3) / 1e12, 1), "TF/s")
ms = M.ours_pair_fwd_bwd(2 * F, 256, V, h, w, precision="fp32", dtype=torch.float32)
print("f32x3", round(ms, 3), "ms", round(2e3 / ms, 1), "clips/s")
