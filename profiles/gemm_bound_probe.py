"""Which resource bounds the big K-major GEMM (dX shape: M = 3136 tokens x 128 sequences, N = 256, K = 512, MN-major B)?
GLF_GEMM_DBG knobs remove one kind of work at a time (results are wrong; timing only)."""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.getcwd())
from glfusion_b200 import _lib as L
lib = L.load()
dev = "cuda:0"
def stream(): return C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(M, N, K, batch, b_mn=1, reps=8, tag=""):
    A = torch.randn(batch, M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(batch, K, N, device=dev).to(torch.bfloat16) if b_mn else torch.randn(batch, N, K, device=dev).to(torch.bfloat16)
    D = torch.empty(batch, M, N, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def launch():
        L.check(lib.glf_gemm_bf16(L.ptr(A), L.ptr(B), L.ptr(D), M, N, K, batch, 0, b_mn, K, N if b_mn else K, N, M * K, N * K, M * N,
                                  None, 1.0, None, N, M * N, 0, 1, None, stream()))
    for _ in range(3): launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{tag:32s} M={M} N={N} K={K} batch={batch}: {ts[len(ts)//2]:.1f} us", flush=True)
for dbg, tag in ((0, "full"), (1, "half the MMAs"), (2, "no epilogue stores"), (3, "half MMAs + no epilogue"), (4, "second A tile = first (L2 hit)"),
                 (7, "all three")):
    os.environ["GLF_GEMM_DBG"] = str(dbg)
    run(3136, 256, 512, 128, tag=tag)
    run(3136, 256, 256, 128, b_mn=0, tag=tag + " (U shape)")
os.environ["GLF_GEMM_DBG"] = "0"
os.environ["GLF_GEMM_CL2"] = "1"
run(3136, 256, 512, 128, tag="2-CTA cluster, A multicast")
run(3136, 256, 256, 128, b_mn=0, tag="2-CTA cluster, A multicast (U shape)")
os.environ["GLF_GEMM_CL2"] = "0"
os.environ["GLF_GEMM_WIDE2"] = "1"
for dbg, tag in ((0, "256x256 CTA tiles"), (2, "256x256 CTA tiles, no epilogue stores")):
    os.environ["GLF_GEMM_DBG"] = str(dbg)
    run(3136, 256, 512, 128, tag=tag)
    run(3136, 256, 256, 128, b_mn=0, tag=tag + " (U shape)")
