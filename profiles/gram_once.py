"""One S = X^T X launch of gram_kernel at the bench shape (for ncu)."""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.getcwd())
from glfusion_b200 import _lib as L
lib = L.load()
dev = "cuda:0"
B, N, Cc = 128, 3136, 256
X = torch.randn(B, N, Cc, device=dev).to(torch.bfloat16)
D = torch.empty(B, Cc, Cc, device=dev, dtype=torch.bfloat16)
rs = torch.empty(B, Cc, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    L.check(lib.glf_gram_contraction(L.ptr(X), L.ptr(X), L.ptr(D), L.ptr(rs), B, N, Cc, Cc, st))
torch.cuda.synchronize()
print("ok")
