"""Correctness + timing of the CTA-pair GEMM (glf_gemm2.cu) against torch on the U / dX shapes."""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from glfusion_b200 import _lib as L
lib = L.load()
dev = "cuda:0"
def stream(): return C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(M, N, K, batch, b_mn, bias, colstats, reps=8):
    torch.manual_seed(0)
    A = torch.randn(batch, M, K, device=dev).to(torch.bfloat16)
    Bk = (torch.randn(batch, N, K, device=dev) * 0.1).to(torch.bfloat16)
    B = Bk.transpose(1, 2).contiguous() if b_mn else Bk
    D = torch.zeros(batch, M, N, device=dev, dtype=torch.bfloat16)
    bi = torch.randn(N, device=dev) if bias else None
    cs = torch.zeros(batch * ((M + 127) // 128) * 4, 2, N, device=dev) if colstats else None
    def launch():
        L.check(lib.glf_gemm_bf16(L.ptr(A), L.ptr(B), L.ptr(D), M, N, K, batch, 0, b_mn, K, N if b_mn else K, N, M * K, N * K, M * N,
                                  L.ptr(bi), 1.0, None, N, M * N, 0, 1, L.ptr(cs), stream()))
    launch(); torch.cuda.synchronize()
    ref = torch.bmm(A.float(), Bk.float().transpose(1, 2)) + (bi if bias else 0)
    err = float((D.float() - ref).norm() / ref.norm())
    msg = f"M={M} N={N} K={K} batch={batch} b_mn={b_mn} bias={bias} colstats={colstats}: rel err {err:.2e}"
    if colstats:
        s1 = cs[:, 0].sum(0); s2 = cs[:, 1].sum(0)
        r1 = D.float().sum((0, 1)); r2 = (D.float() ** 2).sum((0, 1))
        msg += f" colsum err {float((s1 - r1).norm() / r1.norm()):.2e} colsq err {float((s2 - r2).norm() / r2.norm()):.2e}"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(msg, f"| {ts[len(ts)//2]:.1f} us", flush=True)
for pair, dbg in (("1", "0"), ("0", "0"), ("1", "2"), ("0", "2")):
    os.environ["GLF_GEMM_PAIR"] = pair
    os.environ["GLF_GEMM_DBG"] = dbg
    os.environ["GLF_GEMM_PAIR_STATS"] = "1"
    print("GLF_GEMM_PAIR", pair, "GLF_GEMM_DBG", dbg, "(2 = no epilogue: timing only)")
    run(3136, 256, 256, 4, 0, True, True)
    run(3136, 256, 256, 128, 0, True, True)
    run(3136, 256, 512, 128, 1, True, False)
    run(1100, 256, 128, 37, 1, False, True)
    if dbg == "0":
        run(18816, 3072, 2048, 1, 0, True, False, reps=4)      # the C = 2048 projection GEMM (tensor-bound)
        run(18816, 1024, 2048, 1, 1, False, False, reps=4)
