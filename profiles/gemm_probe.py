#!/usr/bin/env python
"""Standalone timings of the GEMM shapes of the cfg2 step (8 clips): python profiles/gemm_probe.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glfusion_b200 import _lib as L  # noqa: E402

dev = "cuda:0"
lib = L.load()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(name, M, N, K, batch, a_mn=0, b_mn=0, lda=None, ldb=None, ldd=None, sA=None, sB=None, sD=None, bias=False,
        addend=False, colstats=False, out_kind=0, split_k=1, alg_bytes=None, reps=20):
    lda = lda or (M if a_mn else K)
    ldb = ldb or (N if b_mn else K)
    ldd = ldd or N
    rowsA = K if a_mn else M
    rowsB = K if b_mn else N
    sA = (rowsA * lda) if sA is None else sA
    sB = (rowsB * ldb) if sB is None else sB
    sD = (M * ldd) if sD is None else sD
    A = torch.randn(max(batch * sA, rowsA * lda), device=dev).to(torch.bfloat16)
    B = torch.randn(max(batch * sB, rowsB * ldb), device=dev).to(torch.bfloat16)
    D = torch.zeros(max(batch * sD, M * ldd), device=dev, dtype=torch.bfloat16 if out_kind == 0 else torch.float32)
    bv = torch.randn(N, device=dev) if bias else None
    ad = torch.randn(batch * M * N, device=dev).to(torch.bfloat16) if addend else None
    cs = torch.zeros(batch * ((M + 127) // 128) * 4 * 2 * N, device=dev) if colstats else None

    def go():
        L.check(lib.glf_gemm_bf16(L.ptr(A), L.ptr(B), L.ptr(D), M, N, K, batch, a_mn, b_mn, lda, ldb, ldd, sA, sB, sD,
                                  L.ptr(bv), 1.0, L.ptr(ad), N, M * N, out_kind, split_k, L.ptr(cs), st()))
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    gb = (alg_bytes or 0) / 1e9
    print(f"{name:34s} {us:8.1f} us  {gb / (us * 1e-6) if gb else 0:7.0f} GB/s")


rows, Ci, Cc, N, Bq = 401408, 128, 256, 3136, 128
P = rows * Cc * 2
run("proj [rows,256]x[384,256]+bias", rows, 3 * Ci, Cc, 1, bias=True, alg_bytes=2.5 * P)
run("proj no bias", rows, 3 * Ci, Cc, 1, alg_bytes=2.5 * P)
run("U  batch128 [3136,128]x[256,128] cs+b", N, Cc, Ci, Bq, lda=3 * Ci, sA=N * 3 * Ci, bias=True, colstats=True, alg_bytes=1.5 * P)
run("U  no colstats", N, Cc, Ci, Bq, lda=3 * Ci, sA=N * 3 * Ci, bias=True, alg_bytes=1.5 * P)
run("U  no colstats no bias", N, Cc, Ci, Bq, lda=3 * Ci, sA=N * 3 * Ci, alg_bytes=1.5 * P)
run("U  dense A (lda=128)", N, Cc, Ci, Bq, bias=True, colstats=True, alg_bytes=1.5 * P)
run("U  unbatched rows x [256,128]", rows, Cc, Ci, 1, lda=3 * Ci, bias=True, colstats=True, alg_bytes=1.5 * P)
run("dPhi [3136,128]x[128,128] cs", N, Ci, Ci, Bq, lda=3 * Ci, sA=N * 3 * Ci, ldd=3 * Ci, sD=N * 3 * Ci, colstats=True, alg_bytes=1.0 * P)
run("dTheta [3136,256]x[128,256]mn cs", N, Ci, Cc, Bq, b_mn=1, ldb=Ci, sB=Cc * Ci, ldd=3 * Ci, sD=N * 3 * Ci, colstats=True, alg_bytes=1.5 * P)
run("dX [rows,384]x[256,384]+addend", rows, Cc, 3 * Ci, 1, addend=True, alg_bytes=3.5 * P)
run("dX no addend", rows, Cc, 3 * Ci, 1, alg_bytes=2.5 * P)
