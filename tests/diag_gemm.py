"""GPU diagnostic (not a pytest file): runs each GEMM variant and prints its error; used when bringing up descriptors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from gpu_util import DEV, gemm
from oracle import tpavi_oracle as O

def mk(batch, rows, K, mn, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(batch, rows, K, generator=g).to(torch.bfloat16)
    d = t.to(DEV)
    return (d.transpose(1, 2).contiguous() if mn else d), t.float()

cases = [(128,128,64,1,0,0),(128,128,256,1,0,0),(128,64,64,1,0,0),(128,256,64,1,0,0),(256,384,256,1,0,0),
         (128,128,64,1,1,1),(128,128,256,2,1,1),(128,256,128,1,1,0),(200,128,128,1,0,1),(128,64,128,1,1,1)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for i, (M,N,K,b,am,bm) in enumerate(cases):
    if only is not None and i != only: continue
    A, Af = mk(b, M, K, am, 1); B, Bf = mk(b, N, K, bm, 2)
    try:
        D, _ = gemm(A, B, M, N, K, b, am, bm)
        ref = torch.matmul(Af, Bf.transpose(1,2))
        err = O.rel_err(D, ref)
        msg = ""
        if err > 1e-2:
            Dc = D.float().cpu()
            blk = [(O.rel_err(Dc[:, r:r+32, c:c+64], ref[:, r:r+32, c:c+64])) for r in range(0, min(M,128), 32) for c in range(0, min(N,128), 64)]
            msg = " blockerr(32x64)=" + ",".join(f"{e:.2f}" for e in blk) + f" D[0,0,:4]={Dc[0,0,:4].tolist()} ref={ref[0,0,:4].tolist()}"
        print(f"case {i} M{M} N{N} K{K} b{b} amn{am} bmn{bm}: rel_err {err:.3e}{msg}", flush=True)
    except Exception as e:
        print(f"case {i} M{M} N{N} K{K} b{b} amn{am} bmn{bm}: EXC {e}", flush=True)
        break
