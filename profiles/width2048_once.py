"""One eager fwd+bwd step of the fusion node at the reference network's own width (C = 2048, 3 views, 28x28, 8 frames) after
warm-up — the command the C = 2048 launch list (ncu --metrics gpu__time_duration.sum) is taken from."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import measure_round2 as M  # noqa: E402

C, V, h, w, B = 2048, 3, 28, 28, 8
f = M.fusion(C)
f4, cl, ct = M.inputs(B, C, V, h, w)
for t in f4:
    t.requires_grad_(True)
dz = torch.randn(B, V, h, w, C, device=M.DEV).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
steps = int(os.environ.get("STEPS", "3"))
for _ in range(steps):
    for t in f4:
        t.grad = None
    f.forward_stacked(f4, cl, ct).backward(dz)
torch.cuda.synchronize()
print("ok")
