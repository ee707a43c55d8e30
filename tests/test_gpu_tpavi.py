"""Parity of the CUDA TPAVIModule (through the nn.Module boundary -> C ABI) with the reference: golden vectors produced
by the reference module itself, and the CPU oracle on seeded inputs.  bf16 tensor-core operands: tolerance 2e-2."""
import pytest
import torch

from conftest import load_golden
from gpu_util import BF16_TOL, DEV, assert_close, grad_tol, dot_algo, golden_params, grad_scale, load_module_from_params
from glfusion_b200 import TPAVIModule
from oracle import tpavi_oracle as O

pytestmark = pytest.mark.gpu

DOT_CASES = ["dot_train_c128", "dot_train_c256", "dot_eval_c128", "dot_nobn_c128", "dot_ragged_c128"]


def _run(m, x, dz):
    x = x.clone().requires_grad_(True)
    z, aux = m(x)
    assert aux == 0
    z.backward(dz)
    torch.cuda.synchronize()
    return z.detach(), x.grad.detach()


@pytest.mark.parametrize("name", DOT_CASES)
@pytest.mark.parametrize("io", ["fp32", "bf16"])
@pytest.mark.parametrize("algo", ["token", "gram"])
def test_golden_dot(name, io, algo):
    """Both exact reassociations of mode='dot' (token-space and Gram form) against the reference's golden vectors."""
    g = load_golden(name)
    B, C, T, H, W, training, bn = [int(v) for v in g["meta"]]
    m = load_module_from_params(TPAVIModule, golden_params(g), C, "dot", bool(bn))
    m.train(bool(training))
    dt = torch.float32 if io == "fp32" else torch.bfloat16
    with dot_algo(algo):
        z, dx = _run(m, g["x"].to(DEV, dt), g["dz"].to(DEV, dt))
    assert tuple(z.shape) == (B, C, T, H, W) and z.dtype == dt
    assert list(z.stride()) == [int(s) for s in g["z_strides"]]     # same strided view as the reference returns
    assert_close("z", z, g["z"], BF16_TOL)
    assert_close("dx", dx, g["dx"], BF16_TOL)
    for k, p in m.named_parameters():
        if k.startswith("align_channel"):
            assert p.grad is None
            continue
        assert_close("grad:" + k, p.grad, g["grad:" + k], grad_tol(k), abs_floor=1e-3)
    if bn:
        sd = m.state_dict()
        for k in ("W_z.1.running_mean", "W_z.1.running_var"):
            assert_close(k, sd[k], g["buf_after:" + k], 1e-2)
        assert int(sd["W_z.1.num_batches_tracked"]) == int(g["buf_after:W_z.1.num_batches_tracked"])


@pytest.mark.parametrize("layout", ["ncthw", "token"])
@pytest.mark.parametrize("algo", ["token", "gram", "auto"])
def test_oracle_dot_layouts(layout, algo):
    """Seeded inputs, both accepted physical layouts of x (NCTHW, and channels-last 'token-major')."""
    B, C, T, H, W = 4, 256, 4, 14, 14
    p = O.init_params(C, seed=21, randomize_affine=True)
    gen = torch.Generator().manual_seed(22)
    x = torch.randn(B, C, T, H, W, generator=gen)
    dz = torch.randn(B, C, T, H, W, generator=gen)
    zo, dxo, go = O.tpavi_fwd_bwd(x, dz, {k: v.clone() for k, v in p.items()}, mode="dot")
    m = load_module_from_params(TPAVIModule, p, C, "dot", True).train()
    xd = x.to(DEV, torch.bfloat16)
    dzd = dz.to(DEV, torch.bfloat16)
    if layout == "token":
        xd = xd.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
        dzd = dzd.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    with dot_algo(algo):
        z, dx = _run(m, xd, dzd)
    if layout == "token":
        assert dx.permute(0, 2, 3, 4, 1).is_contiguous()
    assert_close("z", z, zo, BF16_TOL)
    assert_close("dx", dx, dxo, BF16_TOL)
    for k, pp in m.named_parameters():
        if not k.startswith("align_channel"):
            assert_close("grad:" + k, pp.grad, go[k], grad_tol(k), abs_floor=1e-3)


@pytest.mark.parametrize("C", [128, 256, 64])
def test_gram_many_sequences(C):
    """>= 96 sequences: the Gram form runs one CTA per sequence with no token split (bf16 results written straight
    into the augmented matrices); 4 k-blocks per sequence wrap the 3-stage ring.  C = 64 takes the generic tile GEMM."""
    B, T, H, W = 96, 4, 8, 8
    p = O.init_params(C, seed=51, randomize_affine=True)
    gen = torch.Generator().manual_seed(52)
    x = torch.randn(B, C, T, H, W, generator=gen)
    dz = torch.randn(B, C, T, H, W, generator=gen)
    zo, dxo, go = O.tpavi_fwd_bwd(x, dz, {k: v.clone() for k, v in p.items()}, mode="dot")
    m = load_module_from_params(TPAVIModule, p, C, "dot", True).train()
    with dot_algo("gram"):
        z, dx = _run(m, x.to(DEV, torch.bfloat16), dz.to(DEV, torch.bfloat16))
    assert_close("z", z, zo, BF16_TOL)
    assert_close("dx", dx, dxo, BF16_TOL)
    for k, pp in m.named_parameters():
        if not k.startswith("align_channel"):
            assert_close("grad:" + k, pp.grad, go[k], grad_tol(k), abs_floor=1e-3)


@pytest.mark.parametrize("C,T,H,W", [(96, 2, 16, 16), (512, 4, 24, 24), (192, 3, 20, 20)])
@pytest.mark.parametrize("algo", ["gram", "token"])
def test_dot_other_widths(C, T, H, W, algo):
    """Channel counts away from the tuned 128 / 256: C = 96 and 192 (the Gram contraction falls back to the tile GEMM
    with its row-sum side product, ragged 64-column boxes), C = 512 (two n-tiles per product, C > 256 LayerNorm path)."""
    B = 2
    p = O.init_params(C, seed=61, randomize_affine=True)
    gen = torch.Generator().manual_seed(62)
    x = torch.randn(B, C, T, H, W, generator=gen)
    dz = torch.randn(B, C, T, H, W, generator=gen)
    zo, dxo, go, _ = O.tpavi_dot_closed_form(x.double(), dz.double(), {k: (v.double() if v.is_floating_point() else v)
                                                                       for k, v in p.items()})
    m = load_module_from_params(TPAVIModule, p, C, "dot", True).train()
    with dot_algo(algo):
        z, dx = _run(m, x.to(DEV, torch.bfloat16), dz.to(DEV, torch.bfloat16))
    assert_close("z", z, zo, BF16_TOL)
    assert_close("dx", dx, dxo, BF16_TOL)
    for k, pp in m.named_parameters():
        if not k.startswith("align_channel"):
            assert_close("grad:" + k, pp.grad, go[k], grad_tol(k), abs_floor=1e-3)


def test_eval_no_grad_inference_and_zero_init_trap():
    C = 128
    m = TPAVIModule(C).to(DEV).eval()          # reference init: BN gamma=beta=0 -> z == LayerNorm(x) exactly (F3)
    x = torch.randn(2, C, 2, 8, 8, device=DEV)
    with torch.no_grad():
        z, _ = m(x)
    ln = torch.nn.functional.layer_norm(x.to(torch.bfloat16).float().permute(0, 2, 3, 4, 1), (C,)).permute(0, 4, 1, 2, 3)
    assert (z - ln).abs().max() < 1e-4
    assert int(m.W_z[1].num_batches_tracked) == 0


def test_linearity_property_full_size_cfg2():
    """BASELINE cfg2 frame-as-batch shape (B=16, C=256, T=4 views, 28x28): size-independent checks.
    (1) per-row LayerNorm statistics of z (affine removed) are mean 0 / var 1;
    (2) BN running stats moved toward batch stats; (3) sum of dx over... backward of a zero dz is zero."""
    B, C, T, H, W = 16, 256, 4, 28, 28
    p = O.init_params(C, seed=5, randomize_affine=True)
    p["norm_layer.weight"].fill_(1.0)
    p["norm_layer.bias"].fill_(0.0)
    m = load_module_from_params(TPAVIModule, p, C, "dot", True).train()
    x = torch.randn(B, C, T, H, W, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    z, _ = m(x)
    zt = z.float().permute(0, 2, 3, 4, 1)
    assert zt.mean(-1).abs().max() < 2e-2
    assert (zt.var(-1, unbiased=False) - 1).abs().max() < 5e-2
    z.backward(torch.zeros_like(z))
    torch.cuda.synchronize()
    assert x.grad.abs().max() == 0
    assert int(m.W_z[1].num_batches_tracked) == 1
    assert torch.isfinite(m.W_z[1].running_var).all()


@pytest.mark.parametrize("name", ["embedded_train_c128", "embedded_eval_c128"])
@pytest.mark.parametrize("io", ["fp32", "bf16"])
def test_golden_embedded(name, io):
    """mode='embedded' (softmax attention, ours.py:896-897) against the reference module's golden vectors."""
    g = load_golden(name)
    B, C, T, H, W, training, bn = [int(v) for v in g["meta"]]
    m = load_module_from_params(TPAVIModule, golden_params(g), C, "embedded", bool(bn))
    m.train(bool(training))
    dt = torch.float32 if io == "fp32" else torch.bfloat16
    z, dx = _run(m, g["x"].to(DEV, dt), g["dz"].to(DEV, dt))
    assert_close("z", z, g["z"], BF16_TOL)
    assert_close("dx", dx, g["dx"], BF16_TOL)
    scale = grad_scale([v for k, v in g.items() if k.startswith("grad:")])
    for k, p in m.named_parameters():
        if k.startswith("align_channel"):
            continue
        assert_close("grad:" + k, p.grad, g["grad:" + k], grad_tol(k, "embedded"), zero_scale=scale)


def test_oracle_embedded_cfg2_tokens():
    """Softmax mode at the cfg2 sequence length (4 views x 28 x 28 = 3136 tokens, C=256), ragged batch chunking."""
    B, C, T, H, W = 3, 256, 4, 28, 28
    p = O.init_params(C, seed=51, randomize_affine=True)
    # moderate logits: the reference applies no 1/sqrt(d) scale, random init gives |theta.phi| ~ 1
    gen = torch.Generator().manual_seed(52)
    x = torch.randn(B, C, T, H, W, generator=gen)
    dz = torch.randn(B, C, T, H, W, generator=gen)
    zo, dxo, go = O.tpavi_fwd_bwd(x, dz, {k: v.clone() for k, v in p.items()}, mode="embedded")
    m = load_module_from_params(TPAVIModule, p, C, "embedded", True).train()
    z, dx = _run(m, x.to(DEV, torch.bfloat16), dz.to(DEV, torch.bfloat16))
    assert_close("z", z, zo, BF16_TOL)
    assert_close("dx", dx, dxo, BF16_TOL)
    scale = grad_scale(go.values())
    for k, pp in m.named_parameters():
        if not k.startswith("align_channel"):
            assert_close("grad:" + k, pp.grad, go[k], grad_tol(k, "embedded"), zero_scale=scale)


FP32_TOL = 1e-4      # north_star: within 1e-4 relative error in fp32


@pytest.mark.parametrize("name", DOT_CASES)
def test_golden_dot_fp32_precision(name):
    """compute_precision='fp32' (3-limb bf16 split on tcgen05): the reference's fp32 tolerance, 1e-4."""
    g = load_golden(name)
    B, C, T, H, W, training, bn = [int(v) for v in g["meta"]]
    m = load_module_from_params(TPAVIModule, golden_params(g), C, "dot", bool(bn))
    m.compute_precision = "fp32"
    m.train(bool(training))
    z, dx = _run(m, g["x"].to(DEV), g["dz"].to(DEV))
    assert z.dtype == torch.float32
    assert_close("z", z, g["z"], FP32_TOL)
    assert_close("dx", dx, g["dx"], FP32_TOL)
    scale = grad_scale([v for k, v in g.items() if k.startswith("grad:")])
    for k, p in m.named_parameters():
        if k.startswith("align_channel"):
            continue
        assert_close("grad:" + k, p.grad, g["grad:" + k], FP32_TOL, zero_scale=scale)
    if bn:
        sd = m.state_dict()
        for k in ("W_z.1.running_mean", "W_z.1.running_var"):
            assert_close(k, sd[k], g["buf_after:" + k], FP32_TOL)


def test_fp32_precision_cfg2_tokens_vs_fp64_oracle():
    """C=256, N=3136 (cfg2 geometry), token-major fp32 input, against the fp64 closed-form oracle."""
    B, C, T, H, W = 2, 256, 4, 28, 28
    p = O.init_params(C, seed=61, randomize_affine=True)
    gen = torch.Generator().manual_seed(62)
    x = torch.randn(B, C, T, H, W, generator=gen)
    dz = torch.randn(B, C, T, H, W, generator=gen)
    p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
    zo, dxo, go, _ = O.tpavi_dot_closed_form(x.double(), dz.double(), p64, training=True)
    m = load_module_from_params(TPAVIModule, p, C, "dot", True).train()
    m.compute_precision = "fp32"
    xd = x.to(DEV).permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    dzd = dz.to(DEV).permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    z, dx = _run(m, xd, dzd)
    assert_close("z", z, zo, FP32_TOL)
    assert_close("dx", dx, dxo, FP32_TOL)
    scale = grad_scale(go.values())
    for k, pp in m.named_parameters():
        if not k.startswith("align_channel"):
            assert_close("grad:" + k, pp.grad, go[k], FP32_TOL, zero_scale=scale)


@pytest.mark.parametrize("C,B,T,H,W", [(2048, 2, 3, 8, 8), (512, 2, 2, 9, 7), (64, 3, 2, 6, 6)])
def test_oracle_dot_channel_widths(C, B, T, H, W):
    """The reference network's real width (in_channels=2048, ours.py:1746-1747: 128x256 tiles, K=2048, 8 LayerNorm
    slices per row), an odd one and a tiny one."""
    p = O.init_params(C, seed=71, randomize_affine=True)
    gen = torch.Generator().manual_seed(72)
    x = torch.randn(B, C, T, H, W, generator=gen)
    dz = torch.randn(B, C, T, H, W, generator=gen)
    zo, dxo, go = O.tpavi_fwd_bwd(x, dz, {k: v.clone() for k, v in p.items()}, mode="dot")
    m = load_module_from_params(TPAVIModule, p, C, "dot", True).train()
    z, dx = _run(m, x.to(DEV, torch.bfloat16), dz.to(DEV, torch.bfloat16))
    assert_close("z", z, zo, BF16_TOL)
    assert_close("dx", dx, dxo, BF16_TOL)
    scale = grad_scale(go.values())
    for k, pp in m.named_parameters():
        if not k.startswith("align_channel"):
            assert_close("grad:" + k, pp.grad, go[k], grad_tol(k), zero_scale=scale)


def test_network_width_at_the_network_sequence_length():
    """in_channels = 2048 with the reference network's own sequence (3 views x 28 x 28 = 2 352 tokens, ours.py:1746-1747,
    :1820): the token-space form with its big products on CTA pairs (projection, dTheta, dPhi / dG with alpha and column
    statistics, dX with the residual addend, dWcat with K slices, W_z with several column tiles of statistics) and the
    wide-row LayerNorm ring kernels, against the oracle."""
    C, B, T, H, W = 2048, 2, 3, 28, 28
    p = O.init_params(C, seed=81, randomize_affine=True)
    gen = torch.Generator().manual_seed(82)
    x = torch.randn(B, C, T, H, W, generator=gen)
    dz = torch.randn(B, C, T, H, W, generator=gen)
    zo, dxo, go = O.tpavi_fwd_bwd(x, dz, {k: v.clone() for k, v in p.items()}, mode="dot")
    m = load_module_from_params(TPAVIModule, p, C, "dot", True).train()
    z, dx = _run(m, x.to(DEV, torch.bfloat16), dz.to(DEV, torch.bfloat16))
    assert_close("z", z, zo, BF16_TOL)
    assert_close("dx", dx, dxo, BF16_TOL)
    scale = grad_scale(go.values())
    for k, pp in m.named_parameters():
        if not k.startswith("align_channel"):
            assert_close("grad:" + k, pp.grad, go[k], grad_tol(k), zero_scale=scale)


@pytest.mark.parametrize("C,B,T,H,W", [(128, 3, 3, 5, 7), (256, 2, 1, 9, 15), (256, 1, 2, 13, 20), (64, 2, 2, 6, 6)])
def test_oracle_embedded_ragged(C, B, T, H, W):
    """Softmax mode, token counts that are not multiples of the 128-wide query/key tiles (masking paths of the flash
    kernels: N = 105, 135, 520), both head widths (C'=64, 128), and a head width served by the chunked exact path (32)."""
    p = O.init_params(C, seed=91, randomize_affine=True)
    gen = torch.Generator().manual_seed(92)
    x = torch.randn(B, C, T, H, W, generator=gen)
    dz = torch.randn(B, C, T, H, W, generator=gen)
    zo, dxo, go = O.tpavi_fwd_bwd(x, dz, {k: v.clone() for k, v in p.items()}, mode="embedded")
    m = load_module_from_params(TPAVIModule, p, C, "embedded", True).train()
    z, dx = _run(m, x.to(DEV, torch.bfloat16), dz.to(DEV, torch.bfloat16))
    assert_close("z", z, zo, BF16_TOL)
    assert_close("dx", dx, dxo, BF16_TOL)
    scale = grad_scale(go.values())
    for k, pp in m.named_parameters():
        if not k.startswith("align_channel"):
            assert_close("grad:" + k, pp.grad, go[k], grad_tol(k, "embedded"), zero_scale=scale)


# ------------------------------------------------------------------------------------------------ dp.GradBucket.bind
def _bucket_case(seed=71):
    from glfusion_b200 import dp
    B, C, T, H, W = 2, 128, 2, 8, 8
    p = O.init_params(C, seed=seed, randomize_affine=True)
    gen = torch.Generator().manual_seed(seed + 1)
    xs = [torch.randn(B, C, T, H, W, generator=gen).to(DEV, torch.bfloat16) for _ in range(2)]
    dzs = [torch.randn(B, C, T, H, W, generator=gen).to(DEV, torch.bfloat16) for _ in range(2)]
    m = load_module_from_params(TPAVIModule, p, C, "dot", True).train()
    return dp, m, xs, dzs


def _grads(m):
    return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}


def test_bucket_bind_module_used_twice_in_one_backward():
    """The reference's training step runs the network twice before ONE backward (R/main.py:209 and :221, the cycle pass):
    with the zero-copy bucket bound, the second backward node must not overwrite the first one's gradients."""
    dp, m, xs, dzs = _bucket_case()

    def step():
        for p in m.parameters():
            p.grad = None
        z0, _ = m(xs[0])
        z1, _ = m(xs[1])
        torch.autograd.backward([z0, z1], dzs)
        torch.cuda.synchronize()
        return _grads(m)
    ref = step()                                     # unbound: autograd sums two fresh gradient tensors
    bucket = dp.GradBucket(m._plist())
    bucket.bind([m])
    got = step()
    for k in ref:
        assert_close("twice:" + k, got[k], ref[k], 1e-5, abs_floor=1e-6)
    bucket.allreduce_mean()                          # world size 1: pack / unpack must leave the values alone
    torch.cuda.synchronize()
    for k in ref:
        assert_close("twice+allreduce:" + k, dict(m.named_parameters())[k].grad, ref[k], 1e-5, abs_floor=1e-6)


def test_bucket_bind_gradient_accumulation():
    """Two forward/backward pairs without zeroing in between (gradient accumulation, zero_grad(set_to_none=False)):
    once .grad aliases the bucket view the kernels must not write into it again."""
    dp, m, xs, dzs = _bucket_case(seed=81)

    def two_steps():
        for p in m.parameters():
            p.grad = None
        for x, dz in zip(xs, dzs):
            z, _ = m(x)
            z.backward(dz)
        torch.cuda.synchronize()
        return _grads(m)
    ref = two_steps()
    bucket = dp.GradBucket(m._plist())
    bucket.bind([m])
    got = two_steps()
    for k in ref:
        assert_close("accum:" + k, got[k], ref[k], 1e-5, abs_floor=1e-6)
    # and the single-use step still takes the zero-copy path
    for p in m.parameters():
        p.grad = None
    z, _ = m(xs[0])
    z.backward(dzs[0])
    assert bucket.aliased()


def test_frozen_batchnorm_follows_the_holder_layer():
    """model.train() with the BatchNorm layer alone in eval mode (the usual 'freeze BN' pattern): the kernels must take
    train / eval, eps and momentum from the nn layers themselves."""
    B, C, T, H, W = 2, 128, 2, 8, 8
    p = O.init_params(C, seed=91, randomize_affine=True)
    p["W_z.1.running_mean"] = torch.randn(C) * 0.1
    p["W_z.1.running_var"] = torch.rand(C) + 0.5
    gen = torch.Generator().manual_seed(92)
    x = torch.randn(B, C, T, H, W, generator=gen)
    dz = torch.randn(B, C, T, H, W, generator=gen)
    zo, dxo, _ = O.tpavi_fwd_bwd(x, dz, {k: v.clone() for k, v in p.items()}, mode="dot", training=False)
    m = load_module_from_params(TPAVIModule, p, C, "dot", True).train()
    m.W_z[1].eval()
    before = m.W_z[1].running_mean.clone()
    z, dx = _run(m, x.to(DEV, torch.bfloat16), dz.to(DEV, torch.bfloat16))
    assert_close("z", z, zo, BF16_TOL)
    assert_close("dx", dx, dxo, BF16_TOL)
    assert torch.equal(m.W_z[1].running_mean, before)          # eval-mode BatchNorm does not touch its statistics
    m.W_z[1].momentum = None
    with pytest.raises(NotImplementedError):
        m(x.to(DEV, torch.bfloat16))
