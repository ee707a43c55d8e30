"""Out-of-bounds canaries (compute-sanitizer is closed on this GPU pool): outputs of the building-block kernels live in
the middle of sentinel-filled buffers; ragged shapes exercise the clipping paths (TMA bulk stores of the GEMM epilogue,
partial row tiles of the bulk-copy LayerNorm kernels).  Every byte outside the logical output must keep its sentinel."""
import ctypes as C

import pytest
import torch

from gpu_util import DEV, stream
from glfusion_b200 import _lib as L

pytestmark = pytest.mark.gpu
SENT = 12345.0


def _guarded(n, dtype, pad=4096):
    buf = torch.full((n + 2 * pad,), SENT, dtype=dtype, device=DEV)
    return buf, buf[pad:pad + n]


def _intact(buf, n, pad=4096):
    return bool((buf[:pad] == SENT).all()) and bool((buf[pad + n:] == SENT).all())


@pytest.mark.parametrize("M,N,K,batch,ldd", [(200, 72, 96, 3, 80), (129, 40, 64, 1, 40), (3136, 256, 128, 2, 256),
                                             (33, 32, 200, 2, 48), (257, 136, 72, 1, 136)])
@pytest.mark.parametrize("colstats", [False, True])
def test_gemm_bf16_store_clipping(M, N, K, batch, ldd, colstats):
    lib = L.load()
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(batch, M, K, generator=g).to(torch.bfloat16).to(DEV)
    B = torch.randn(batch, N, K, generator=g).to(torch.bfloat16).to(DEV)
    n_out = batch * M * ldd
    buf, D = _guarded(n_out, torch.bfloat16)
    cs = torch.zeros(batch * ((M + 127) // 128) * 4 * 2 * N, device=DEV) if colstats else None
    L.check(lib.glf_gemm_bf16(L.ptr(A), L.ptr(B), L.ptr(D), M, N, K, batch, 0, 0, K, K, ldd, M * K, N * K, M * ldd,
                              None, 1.0, None, 0, 0, 0, 1, L.ptr(cs), stream()))
    torch.cuda.synchronize()
    assert _intact(buf, n_out)
    Dv = D.view(batch, M, ldd)
    if ldd > N:
        assert bool((Dv[:, :, N:] == SENT).all()), "padding columns between rows were overwritten"
    ref = torch.einsum("bmk,bnk->bmn", A.float(), B.float())
    err = (Dv[:, :, :N].float() - ref).norm() / ref.norm()
    assert err < 1e-2


@pytest.mark.parametrize("rows,Cc", [(61, 256), (1, 128), (89, 64), (30 * 7 + 1, 256)])
def test_ln_pair_row_clipping(rows, Cc):
    lib = L.load()
    g = torch.Generator().manual_seed(rows)

    def bf(*s):
        return torch.randn(*s, generator=g).to(torch.bfloat16).to(DEV)
    U, X = [bf(rows, Cc), bf(rows, Cc)], [bf(rows, Cc), bf(rows, Cc)]
    vec = [[(torch.rand(Cc, generator=g) + 0.5).to(DEV) for _ in range(6)] for _ in range(2)]
    zbuf, Z = _guarded(rows * Cc, torch.bfloat16)
    stat = [_guarded(rows, torch.float32, pad=256) for _ in range(4)]

    def tab(ts):
        return (C.c_void_p * 2)(*[t.data_ptr() for t in ts])
    L.check(lib.glf_bn_res_ln_pair_fwd(rows, Cc, tab(U), tab(X), tab([v[0] for v in vec]), tab([v[1] for v in vec]),
                                       tab([v[2] for v in vec]), tab([v[3] for v in vec]), L.ptr(Z),
                                       tab([stat[0][1], stat[1][1]]), tab([stat[2][1], stat[3][1]]), 1e-5, 0, stream()))
    torch.cuda.synchronize()
    assert _intact(zbuf, rows * Cc)
    for b, _ in stat:
        assert _intact(b, rows, pad=256)
    assert torch.isfinite(Z.float()).all()
    dz = bf(rows, Cc)
    dvb = [_guarded(rows * Cc, torch.bfloat16) for _ in range(2)]
    nbmax = lib.glf_bn_res_ln_bwd_max_blocks()
    part = [torch.zeros(nbmax * 4 * Cc, device=DEV) for _ in range(2)]
    nb = C.c_int(0)
    L.check(lib.glf_bn_res_ln_pair_bwd(rows, Cc, L.ptr(dz), tab(U), tab(X), tab([v[0] for v in vec]),
                                       tab([v[1] for v in vec]), tab([v[4] for v in vec]), tab([v[5] for v in vec]),
                                       tab([v[2] for v in vec]), tab([stat[0][1], stat[1][1]]),
                                       tab([stat[2][1], stat[3][1]]), tab([d[1] for d in dvb]), tab(part), C.byref(nb),
                                       stream()))
    torch.cuda.synchronize()
    for b, v in dvb:
        assert _intact(b, rows * Cc)
        assert torch.isfinite(v.float()).all()
    assert 0 < nb.value <= nbmax
