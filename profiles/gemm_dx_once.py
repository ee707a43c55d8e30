"""One launch of the dX-shaped GEMM (for ncu): M = 3136 x 128 sequences, N = 256, K = 512, MN-major B, bias + addend."""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.getcwd())
from glfusion_b200 import _lib as L
lib = L.load()
dev = "cuda:0"
B, N, Cc = 128, 3136, 256
A = torch.randn(B, N, 2 * Cc, device=dev).to(torch.bfloat16)
Bm = (torch.randn(B, 2 * Cc, Cc, device=dev) * 0.05).to(torch.bfloat16)
bias = torch.randn(Cc, device=dev)
D = torch.empty(B, N, Cc, device=dev, dtype=torch.bfloat16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
addend = "--no-addend" not in sys.argv
for _ in range(3):
    L.check(lib.glf_gemm_bf16(L.ptr(A), L.ptr(Bm), L.ptr(D), N, Cc, 2 * Cc, B, 0, 1, 2 * Cc, Cc, Cc, N * 2 * Cc, Cc * 2 * Cc, N * Cc,
                              L.ptr(bias), 1.0, L.ptr(A) if addend else None, 2 * Cc, N * 2 * Cc, 0, 1, None, st))
torch.cuda.synchronize()
print("ok")
