import sys, os, ctypes as C, torch
sys.path.insert(0, os.getcwd())
from glfusion_b200 import _lib as L
lib = L.load()
dev = "cuda:0"
def stream(): return C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(M, N, K, batch, addend=False, colstats=False, bias=False, b_mn=0, reps=10):
    A = torch.randn(batch, M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(batch, N, K, device=dev).to(torch.bfloat16)
    D = torch.empty(batch, M, N, device=dev, dtype=torch.bfloat16)
    add = torch.randn(batch, M, N, device=dev).to(torch.bfloat16) if addend else None
    bi = torch.randn(N, device=dev) if bias else None
    cs = torch.zeros(batch * ((M + 127) // 128) * 4, 2, N, device=dev) if colstats else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def launch():
        L.check(lib.glf_gemm_bf16(L.ptr(A), L.ptr(B), L.ptr(D), M, N, K, batch, 0, b_mn, K, K if not b_mn else N, N, M * K, N * K, M * N,
                                  L.ptr(bi), 1.0, L.ptr(add), N, M * N, 0, 1, L.ptr(cs), stream()))
    for _ in range(3): launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    byts = (batch * M * K + batch * M * N + (batch * M * N if addend else 0)) * 2
    print(f"M={M} N={N} K={K} batch={batch} addend={addend} colstats={colstats} bias={bias}: {ts[len(ts)//2]:.1f} us  ({byts/ts[len(ts)//2]/1e3:.0f} GB/s algorithmic)", flush=True)
for env in ("2", "1"):
    os.environ["GLF_GEMM_MT"] = env
    print("GLF_GEMM_MT", env)
    run(3136, 256, 256, 128)
    run(3136, 256, 256, 128, colstats=True)
    run(3136, 256, 256, 128, colstats=True, bias=True)
    run(3136, 256, 512, 128)
    run(3136, 256, 512, 128, addend=True)
    run(3136, 256, 512, 128, addend=True, bias=True)
    run(3136, 256, 128, 128)
    run(3136, 256, 64, 128)
