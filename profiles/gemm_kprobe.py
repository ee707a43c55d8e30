import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/profiles')
import importlib.util
spec = importlib.util.spec_from_file_location("gp", "/root/repo/profiles/gemm_probe.py")
src = open("/root/repo/profiles/gemm_probe.py").read().split("rows, Ci, Cc, N, Bq =")[0]
exec(src)
M = 100352
for K in (64, 128, 256, 512, 1024):
    tiles = (M // 128)
    run(f"M={M} N=128 K={K} (tiles/SM {tiles/148:.1f})", M, 128, K, 1, alg_bytes=M * K * 2 + M * 128 * 2)
for K in (64, 128, 256, 512):
    run(f"M={M} N=256 K={K}", M, 256, K, 1, alg_bytes=M * K * 2 + M * 256 * 2)
