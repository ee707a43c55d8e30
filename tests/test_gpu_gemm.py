"""tcgen05 GEMM building block (glf_gemm_bf16) vs torch fp32 matmul on bf16-rounded operands."""
import pytest
import torch

from glfusion_b200 import _lib as L
from gpu_util import DEV, gemm, stream
from oracle import tpavi_oracle as O

pytestmark = pytest.mark.gpu


def _mk(batch, rows, K, mn, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = torch.randn(batch, rows, K, generator=g).to(torch.bfloat16)
    dev = t.to(DEV)
    return (dev.transpose(1, 2).contiguous() if mn else dev), t.float()


@pytest.mark.parametrize("M,N,K,batch,a_mn,b_mn", [
    (128, 128, 64, 1, 0, 0),          # one tile, one k-block
    (256, 384, 256, 1, 0, 0),         # projection shape (C=256 -> 3C')
    (300, 256, 128, 2, 0, 0),         # ragged M, BN=256, batched (U = Theta W'^T)
    (128, 64, 192, 1, 0, 0),          # BN=64
    (128, 128, 200, 1, 0, 0),         # ragged K (TMA zero fill)
    (128, 128, 256, 3, 1, 1),         # token contraction, both MN-major (Phi^T G)
    (256, 128, 105, 2, 1, 1),         # ragged K with MN-major operands
    (384, 256, 512, 1, 1, 1),         # dWcat shape
    (200, 128, 128, 1, 0, 1),
    (128, 256, 128, 1, 1, 0),
    (3136, 384, 256, 1, 0, 0),        # cfg2 one frame
])
def test_gemm_bf16_out(M, N, K, batch, a_mn, b_mn):
    A, Af = _mk(batch, M, K, a_mn, 1)
    B, Bf = _mk(batch, N, K, b_mn, 2)
    D, _ = gemm(A, B, M, N, K, batch, a_mn, b_mn)
    ref = torch.matmul(Af, Bf.transpose(1, 2))
    assert O.rel_err(D, ref) < 6e-3


def test_gemm_bias_alpha_addend_colstats():
    M, N, K, batch = 333, 256, 128, 2
    A, Af = _mk(batch, M, K, 0, 3)
    B, Bf = _mk(batch, N, K, 0, 4)
    bias = torch.randn(N, device=DEV)
    add = torch.randn(batch, M, N, device=DEV).to(torch.bfloat16)
    D, cs = gemm(A, B, M, N, K, batch, bias=bias, alpha=0.5, colstats=True)
    ref = 0.5 * torch.matmul(Af, Bf.transpose(1, 2)) + bias.cpu()
    assert O.rel_err(D, ref) < 6e-3
    cs = cs.sum(0).cpu()                                   # the table's rows sum to the statistics over all rows
    Dr = D.float().cpu()
    assert O.rel_err(cs[0], Dr.sum((0, 1))) < 1e-4          # statistics are of the stored (bf16-rounded) values
    assert O.rel_err(cs[1], (Dr * Dr).sum((0, 1))) < 1e-4
    D2, _ = gemm(A, B, M, N, K, batch, bias=bias, alpha=0.5, addend=add)
    assert O.rel_err(D2, ref + add.float().cpu()) < 6e-3


def test_gemm_shared_b_and_strided_a():
    # A is a column slice of a wider buffer (Theta inside P=[Theta|Phi|G]): leading dimension != K
    import ctypes as C
    from glfusion_b200 import _lib as L
    from gpu_util import stream
    M, N, K = 256, 256, 128
    P = torch.randn(M, 3 * K, device=DEV).to(torch.bfloat16)
    W = torch.randn(N, K, device=DEV).to(torch.bfloat16)
    D = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    A = P[:, K:2 * K]
    L.check(L.load().glf_gemm_bf16(C.c_void_p(A.data_ptr()), L.ptr(W), L.ptr(D), M, N, K, 1, 0, 0, 3 * K, K, N, 0, 0, 0,
                                   None, 1.0, None, 0, 0, 0, 1, None, stream()))
    torch.cuda.synchronize()
    assert O.rel_err(D, A.float() @ W.float().T) < 6e-3


@pytest.mark.parametrize("split_k", [1, 4, 13])
def test_gemm_splitk_fp32_atomic(split_k):
    M, N, K, batch = 128, 128, 3136, 2
    A, Af = _mk(batch, M, K, 1, 5)
    B, Bf = _mk(batch, N, K, 1, 6)
    D, _ = gemm(A, B, M, N, K, batch, 1, 1, out_kind=2, split_k=split_k, alpha=1.0 / K)
    ref = torch.matmul(Af, Bf.transpose(1, 2)) / K
    assert O.rel_err(D, ref) < 1e-5
    D1, _ = gemm(A, B, M, N, K, batch, 1, 1, out_kind=1, split_k=1)
    assert O.rel_err(D1, ref * K) < 1e-5


@pytest.mark.parametrize("M,N,K,batch,a_mn,b_mn", [
    (2048, 1024, 200, 10, 1, 1),      # dW'_b at the network's width: token contraction, ragged K tail, batched
    (1024, 1024, 320, 40, 1, 1),      # Phi^T G / dM
    (4736, 1024, 256, 4, 0, 1),       # dTheta = dU W' (B read MN-major), ragged M (4736 = 18.5 x 256)
    (18944, 512, 128, 1, 0, 0),       # K-major, several column tiles
    (1152, 512, 128, 40, 1, 0),       # MN-major A with a K-major B
])
def test_gemm_cta_pair_operand_layouts(M, N, K, batch, a_mn, b_mn):
    """Shapes that route to the CTA-pair kernel (tcgen05 cta_group::2, csrc/glf_gemm2.cu): every operand-layout
    combination, ragged M and K, batched operands."""
    A, Af = _mk(batch, M, K, a_mn, 11)
    B, Bf = _mk(batch, N, K, b_mn, 12)
    D, _ = gemm(A, B, M, N, K, batch, a_mn, b_mn)
    ref = torch.matmul(Af, Bf.transpose(1, 2))
    assert O.rel_err(D, ref) < 6e-3


def test_gemm_cta_pair_epilogues():
    """alpha, bias, residual addend, BatchNorm column statistics over several column tiles, fp32 atomic output with K
    slices and a batch-reduced output — the epilogues the C = 2048 token-space path uses on the pair kernel."""
    M, N, K, batch = 2352, 1024, 1024, 8
    A, Af = _mk(batch, M, K, 0, 13)
    B, Bf = _mk(batch, N, K, 0, 14)
    ref = torch.matmul(Af, Bf.transpose(1, 2))
    bias = torch.randn(N, device=DEV)
    add = torch.randn(batch, M, N, device=DEV).to(torch.bfloat16)
    D, cs = gemm(A, B, M, N, K, batch, bias=bias, alpha=0.25, colstats=True)
    want = 0.25 * ref + bias.cpu()
    assert O.rel_err(D, want) < 6e-3
    cs = cs.sum(0).cpu()
    Dr = D.float().cpu()
    assert O.rel_err(cs[0], Dr.sum((0, 1))) < 1e-4
    assert O.rel_err(cs[1], (Dr * Dr).sum((0, 1))) < 1e-4
    D2, _ = gemm(A, B, M, N, K, batch, alpha=0.25, addend=add)
    assert O.rel_err(D2, 0.25 * ref + add.float().cpu()) < 6e-3
    # token contraction into fp32 with K slices (dWcat = dP^T X): A, B MN-major, K = rows
    M2, N2, K2 = 3072, 2048, 2368
    A2, A2f = _mk(1, M2, K2, 1, 15)
    B2, B2f = _mk(1, N2, K2, 1, 16)
    for split in (1, 3):
        D3, _ = gemm(A2, B2, M2, N2, K2, 1, 1, 1, out_kind=2, split_k=split, alpha=1.0 / K2)
        assert O.rel_err(D3, torch.matmul(A2f, B2f.transpose(1, 2)) / K2) < 1e-5


@pytest.mark.parametrize("M,K,batch,shared_b", [(3136, 256, 16, False), (1000, 128, 40, False), (50176, 128, 1, True),
                                                 (2049, 64, 20, False), (1536, 192, 30, False)])
def test_gemm_resident_b_operand(M, K, batch, shared_b):
    """N = 256, K <= 256, many row tiles per B operand: the kernel that keeps B in shared memory and streams only A
    (csrc/glf_gemm3.cu; U = X Q^T + c of the Gram form).  CTAs own contiguous tile ranges that straddle batch entries;
    ragged M; per-batch bias; BatchNorm column statistics as per-CTA running sums."""
    N = 256
    A, Af = _mk(batch, M, K, 0, 21)
    B, Bf = _mk(1 if shared_b else batch, N, K, 0, 22)
    ref = torch.matmul(Af, Bf.transpose(1, 2))
    D, _ = gemm(A, B, M, N, K, batch, shared_b=shared_b)
    assert O.rel_err(D, ref) < 6e-3
    bias = torch.randn(N, device=DEV)
    D, cs = gemm(A, B, M, N, K, batch, bias=bias, alpha=0.5, colstats=True, shared_b=shared_b)
    assert O.rel_err(D, 0.5 * ref + bias.cpu()) < 6e-3
    cs = cs.sum(0).cpu()
    Dr = D.float().cpu()
    assert O.rel_err(cs[0], Dr.sum((0, 1))) < 1e-4
    assert O.rel_err(cs[1], (Dr * Dr).sum((0, 1))) < 1e-4


def test_transpose_pack_roundtrip():
    import ctypes as C
    from glfusion_b200 import _lib as L
    from gpu_util import stream
    x = torch.randn(3, 40, 105, device=DEV)
    out = torch.empty(3, 105, 40, dtype=torch.bfloat16, device=DEV)
    L.check(L.load().glf_transpose(L.ptr(x), L.ptr(out), 3, 40, 105, L.DTYPE_F32, L.DTYPE_BF16, stream()))
    back = torch.empty(3, 40, 105, device=DEV)
    L.check(L.load().glf_transpose(L.ptr(out), L.ptr(back), 3, 105, 40, L.DTYPE_BF16, L.DTYPE_F32, stream()))
    torch.cuda.synchronize()
    assert torch.equal(out, x.transpose(1, 2).to(torch.bfloat16))
    assert torch.equal(back, x.to(torch.bfloat16).float())


@pytest.mark.parametrize("split_k", [1, 4])
@pytest.mark.parametrize("N", [256, 264, 64])
def test_gemm_rowsum_side_product(split_k, N):
    """Token contraction S_b = A_b^T X_b (both operands MN-major) with the row sums of A^T riding along on the tensor
    cores (the Gram form's S = X^T X, s = X^T 1)."""
    torch.manual_seed(5)
    batch, M, K = 3, 200, 1000
    A = torch.randn(batch, K, M, device=DEV).to(torch.bfloat16)       # [K, M]: MN-major A
    Bm = torch.randn(batch, K, N, device=DEV).to(torch.bfloat16)
    D = torch.zeros(batch, M, N, device=DEV)
    rs = torch.zeros(batch, M, device=DEV)
    lib = L.load()
    L.check(lib.glf_gemm_bf16_ex(L.ptr(A), L.ptr(Bm), L.ptr(D), M, N, K, batch, 1, 1, M, N, N, K * M, K * N, M * N,
                                 None, 0, 1.0, 2 if split_k > 1 else 1, split_k, L.ptr(rs), stream()))
    torch.cuda.synchronize()
    ref = torch.einsum("bkm,bkn->bmn", A.float(), Bm.float())
    assert O.rel_err(D, ref) < 1e-5
    assert O.rel_err(rs, A.float().sum(1)) < 1e-5


def test_gemm_per_batch_bias():
    torch.manual_seed(6)
    batch, M, N, K = 4, 300, 256, 256
    A = torch.randn(batch, M, K, device=DEV).to(torch.bfloat16)
    Bm = torch.randn(batch, N, K, device=DEV).to(torch.bfloat16)
    bias = torch.randn(batch, N, device=DEV)
    D = torch.zeros(batch, M, N, device=DEV, dtype=torch.bfloat16)
    lib = L.load()
    L.check(lib.glf_gemm_bf16_ex(L.ptr(A), L.ptr(Bm), L.ptr(D), M, N, K, batch, 0, 0, K, K, N, M * K, N * K, M * N,
                                 L.ptr(bias), N, 1.0, 0, 1, None, stream()))
    torch.cuda.synchronize()
    ref = torch.einsum("bmk,bnk->bmn", A.float(), Bm.float()) + bias[:, None, :]
    assert O.rel_err(D, ref) < 6e-3


@pytest.mark.parametrize("C,N,B,same", [(256, 3136, 5, True), (128, 200, 3, False), (256, 130, 150, False)])
def test_gram_contraction_kernel(C, N, B, same):
    """One CTA per sequence: D_b = A_b^T X_b and the column sums of A (ragged token counts, ring wrap, > 148 CTAs)."""
    torch.manual_seed(7)
    X = torch.randn(B, N, C, device=DEV).to(torch.bfloat16)
    A = X if same else torch.randn(B, N, C, device=DEV).to(torch.bfloat16)
    ldd = C + 8
    D = torch.zeros(B, ldd, ldd, device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(B, C, device=DEV)
    L.check(L.load().glf_gram_contraction(L.ptr(A), L.ptr(X), L.ptr(D), L.ptr(cs), B, N, C, ldd, stream()))
    torch.cuda.synchronize()
    ref = torch.einsum("bnm,bnk->bmk", A.float(), X.float())
    assert O.rel_err(D[:, :C, :C], ref) < 4e-3          # bf16 rounding of the stored result
    assert float(D[:, C:, :].abs().max()) == 0 and float(D[:, :, C:].abs().max()) == 0   # border untouched
    assert O.rel_err(cs, A.float().sum(1)) < 1e-5


@pytest.mark.parametrize("M,N,K,batch,b_mn", [
    (1100, 256, 256, 3, 0),          # ragged last super-tile (1100 = 4 x 256 + 76), two n-tiles
    (1153, 128, 512, 2, 1),          # the second 128-row half of the last super-tile is entirely out of range
    (3136, 256, 256, 5, 0),          # U = X Q^T of the Gram form, one cfg2 sequence per batch entry
])
def test_gemm_256_row_super_tiles(M, N, K, batch, b_mn):
    """M >= 1024 with a K-major A operand runs 256-row CTA tiles (two TMEM accumulators share each B k-block):
    plain output, bias + residual addend, and per-CTA column statistics."""
    A, Af = _mk(batch, M, K, 0, 11)
    B, Bf = _mk(batch, N, K, b_mn, 12)
    ref = torch.matmul(Af, Bf.transpose(1, 2))
    D, _ = gemm(A, B, M, N, K, batch, 0, b_mn)
    assert O.rel_err(D, ref) < 6e-3
    bias = torch.randn(N, device=DEV)
    add = torch.randn(batch, M, N, device=DEV).to(torch.bfloat16)
    D2, _ = gemm(A, B, M, N, K, batch, 0, b_mn, bias=bias, addend=add)
    assert O.rel_err(D2, ref + bias.cpu() + add.float().cpu()) < 6e-3
    D3, cs = gemm(A, B, M, N, K, batch, 0, b_mn, colstats=True)
    cs = cs.sum(0).cpu()
    Dr = D3.float().cpu()
    assert O.rel_err(cs[0], Dr.sum((0, 1))) < 1e-4
    assert O.rel_err(cs[1], (Dr * Dr).sum((0, 1))) < 1e-4
