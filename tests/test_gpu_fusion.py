"""Gate + concat + MGFM + MLFM + sum (GlobalLocalFusion) against the glue golden (reference lines ours.py:1802-1834
run around two reference modules) and the CPU oracle."""
import pytest
import torch

from conftest import load_golden
from gpu_util import BF16_TOL, DEV, assert_close, grad_tol, dot_algo, golden_params
from glfusion_b200 import GlobalLocalFusion
from oracle import tpavi_oracle as O

pytestmark = pytest.mark.gpu


def _build(C, pg, pl):
    f = GlobalLocalFusion(in_channels=C)
    f.global_attn.load_state_dict({k: v.clone() for k, v in pg.items()}, strict=True)
    f.local_attn.load_state_dict({k: v.clone() for k, v in pl.items()}, strict=True)
    return f.to(DEV).train()


@pytest.mark.parametrize("io", ["fp32", "bf16"])
@pytest.mark.parametrize("algo", ["token", "gram"])
def test_glue_golden(io, algo):
    g = load_golden("glue_dot_c128")
    B, C, V, h, w = [int(v) for v in g["meta"]]
    dt = torch.float32 if io == "fp32" else torch.bfloat16
    f = _build(C, golden_params(g, "param_g:"), golden_params(g, "param_l:"))
    f4 = [g[f"f4:{v}"].to(DEV, dt).requires_grad_(True) for v in range(V)]
    cl = [g[f"cls:{v}"].to(DEV).requires_grad_(True) for v in range(V)]
    ct = [g[f"ctr:{v}"].to(DEV).requires_grad_(True) for v in range(V)]
    with dot_algo(algo):
        out = f({str(v): f4[v] for v in range(V)}, {str(v): cl[v] for v in range(V)}, {str(v): ct[v] for v in range(V)})
        torch.autograd.backward([out[str(v)] for v in range(V)], [g[f"d_out:{v}"].to(DEV, dt) for v in range(V)])
    torch.cuda.synchronize()
    for v in range(V):
        assert_close(f"out:{v}", out[str(v)], g[f"out:{v}"], BF16_TOL)
        assert_close(f"df4:{v}", f4[v].grad, g[f"df4:{v}"], BF16_TOL)
        assert_close(f"dcls:{v}", cl[v].grad, g[f"dcls:{v}"], 4e-2)
        assert_close(f"dctr:{v}", ct[v].grad, g[f"dctr:{v}"], 4e-2)
    for tag, mod in (("g", f.global_attn), ("l", f.local_attn)):
        for k, p in mod.named_parameters():
            if not k.startswith("align_channel"):
                assert_close(f"grad_{tag}:{k}", p.grad, g[f"grad_{tag}:{k}"], grad_tol(k), abs_floor=1e-3)


@pytest.mark.parametrize("algo", ["token", "gram"])
def test_cfg2_shape_against_oracle(algo):
    """4 views x 28x28 tokens x C=256 (BASELINE cfg2 token geometry), 2 frames as batch, seeded oracle comparison."""
    B, C, V, h, w = 2, 256, 4, 28, 28
    pg = O.init_params(C, seed=31, randomize_affine=True)
    pl = O.init_params(C, seed=32, randomize_affine=True)
    gen = torch.Generator().manual_seed(33)
    f4 = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    cl = [torch.randn(B, 5, h, w, generator=gen) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=gen) for _ in range(V)]
    do = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    outs, df4, dcls, dctr, gg, gl = O.fusion_fwd_bwd(f4, cl, ct, do, {k: v.clone() for k, v in pg.items()},
                                                     {k: v.clone() for k, v in pl.items()})
    f = _build(C, pg, pl)
    f4d = [t.to(DEV, torch.bfloat16).requires_grad_(True) for t in f4]
    cld = [t.to(DEV).requires_grad_(True) for t in cl]
    ctd = [t.to(DEV).requires_grad_(True) for t in ct]
    with dot_algo(algo):
        out = f.forward_stacked(f4d, cld, ctd)
        out.backward(torch.stack(do, dim=2).to(DEV, torch.bfloat16))
    torch.cuda.synchronize()
    for v in range(V):
        assert_close(f"out:{v}", out[:, :, v], outs[v], BF16_TOL)
        assert_close(f"df4:{v}", f4d[v].grad, df4[v], BF16_TOL)
        assert_close(f"dctr:{v}", ctd[v].grad, dctr[v], 4e-2)
    for mod, ref in ((f.local_attn, gl), (f.global_attn, gg)):
        for k, p in mod.named_parameters():
            if not k.startswith("align_channel"):
                assert_close("grad:" + k, p.grad, ref[k], grad_tol(k), abs_floor=1e-3)


@pytest.mark.parametrize("h,w,V,io", [(5, 7, 3, "fp32"), (5, 7, 2, "bf16"), (9, 8, 1, "bf16")])
@pytest.mark.parametrize("algo", ["token", "gram"])
def test_ragged_spatial_sizes_against_oracle(h, w, V, io, algo):
    """Odd h*w (scalar NCHW path), tokens not a multiple of any tile, single view."""
    B, C = 3, 128
    dt = torch.float32 if io == "fp32" else torch.bfloat16
    pg = O.init_params(C, seed=41, randomize_affine=True)
    pl = O.init_params(C, seed=42, randomize_affine=True)
    gen = torch.Generator().manual_seed(43)
    f4 = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    cl = [torch.randn(B, 5, h, w, generator=gen) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=gen) for _ in range(V)]
    do = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    outs, df4, dcls, dctr, gg, gl = O.fusion_fwd_bwd(f4, cl, ct, do, {k: v.clone() for k, v in pg.items()},
                                                     {k: v.clone() for k, v in pl.items()})
    f = _build(C, pg, pl)
    f4d = [t.to(DEV, dt).requires_grad_(True) for t in f4]
    cld = [t.to(DEV).requires_grad_(True) for t in cl]
    ctd = [t.to(DEV).requires_grad_(True) for t in ct]
    with dot_algo(algo):
        out = f.forward_stacked(f4d, cld, ctd)
        out.backward(torch.stack(do, dim=2).to(DEV, dt))
    torch.cuda.synchronize()
    for v in range(V):
        assert_close(f"out:{v}", out[:, :, v], outs[v], BF16_TOL)
        assert_close(f"df4:{v}", f4d[v].grad, df4[v], BF16_TOL)
        assert_close(f"dcls:{v}", cld[v].grad, dcls[v], 4e-2)
        assert_close(f"dctr:{v}", ctd[v].grad, dctr[v], 4e-2)
    for k, p in f.global_attn.named_parameters():
        if not k.startswith("align_channel"):
            assert_close("grad_g:" + k, p.grad, gg[k], grad_tol(k), abs_floor=1e-3)


def test_glue_golden_fp32_precision():
    """Whole fused call site with compute_precision='fp32': 1e-4 against the reference golden."""
    g = load_golden("glue_dot_c128")
    B, C, V, h, w = [int(v) for v in g["meta"]]
    f = _build(C, golden_params(g, "param_g:"), golden_params(g, "param_l:"))
    f.global_attn.compute_precision = "fp32"
    f.local_attn.compute_precision = "fp32"
    f4 = [g[f"f4:{v}"].to(DEV).requires_grad_(True) for v in range(V)]
    cl = [g[f"cls:{v}"].to(DEV).requires_grad_(True) for v in range(V)]
    ct = [g[f"ctr:{v}"].to(DEV).requires_grad_(True) for v in range(V)]
    out = f({str(v): f4[v] for v in range(V)}, {str(v): cl[v] for v in range(V)}, {str(v): ct[v] for v in range(V)})
    torch.autograd.backward([out[str(v)] for v in range(V)], [g[f"d_out:{v}"].to(DEV) for v in range(V)])
    torch.cuda.synchronize()
    for v in range(V):
        assert_close(f"out:{v}", out[str(v)], g[f"out:{v}"], 1e-4)
        assert_close(f"df4:{v}", f4[v].grad, g[f"df4:{v}"], 1e-4)
        assert_close(f"dcls:{v}", cl[v].grad, g[f"dcls:{v}"], 5e-4)      # __expf sigmoids in the gate chain
        assert_close(f"dctr:{v}", ct[v].grad, g[f"dctr:{v}"], 5e-4)


def test_forward_parts_matches_reference_return_values():
    """ours.py:1843 returns f4_global_fusion / f4_local_fusion next to the fused sum, and the cycle-consistency pass
    (main.py:211-235) back-propagates through the spatial sums of the GLOBAL part alone."""
    B, C, V, h, w = 3, 128, 3, 12, 10
    pg = O.init_params(C, seed=71, randomize_affine=True)
    pl = O.init_params(C, seed=72, randomize_affine=True)
    gen = torch.Generator().manual_seed(73)
    f4 = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    cl = [torch.randn(B, 5, h, w, generator=gen) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=gen) for _ in range(V)]
    wsum = torch.randn(B, C, generator=gen)
    # oracle: the reference's lines on CPU
    f4o = [t.clone().requires_grad_(True) for t in f4]
    qg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k not in O.BUFFER_KEYS else v.clone()) for k, v in pg.items()}
    ql = {k: v.clone() for k, v in pl.items()}
    xg, xl = O.gate_concat(f4o, cl, ct)
    zg = O.tpavi_forward(xg, qg)
    zl = O.tpavi_forward(xl, ql)
    cyc = sum((zg[:, :, i].sum(dim=(2, 3)) * wsum).sum() for i in range(V))      # main.py:229 spatial sums
    cyc.backward()
    f = _build(C, pg, pl)
    f4d = [t.to(DEV, torch.bfloat16).requires_grad_(True) for t in f4]
    keys = [str(i) for i in range(V)]
    fus, glob, loc = f.forward_parts(dict(zip(keys, f4d)), dict(zip(keys, [t.to(DEV) for t in cl])),
                                     dict(zip(keys, [t.to(DEV) for t in ct])))
    for i, k in enumerate(keys):
        assert_close(f"global:{k}", glob[k], zg[:, :, i], BF16_TOL)
        assert_close(f"local:{k}", loc[k], zl[:, :, i], BF16_TOL)
        assert_close(f"fusion:{k}", fus[k], zg[:, :, i] + zl[:, :, i], BF16_TOL)
    cyc_d = sum((glob[k].float().sum(dim=(2, 3)) * wsum.to(DEV)).sum() for k in keys)
    cyc_d.backward()
    torch.cuda.synchronize()
    for i in range(V):
        assert_close(f"df4:{i}", f4d[i].grad, f4o[i].grad, 3e-2)
    assert f.local_attn.theta.weight.grad is None            # the local block is not on the cycle pass's path
    assert_close("grad_g:theta.weight", f.global_attn.theta.weight.grad, qg["theta.weight"].grad, 3e-2)


def test_bench_configuration_graph_replay():
    """The EXACT configuration bench.py times: 8 clips = 128 sequences x 3136 tokens (4 views x 28 x 28), C = 256, bf16,
    forward_stacked (fused two-stream node, Gram form with the per-sequence chain kernels), the step captured in a CUDA
    graph and REPLAYED; outputs and input gradients of the replay against the oracle's closed form evaluated in fp64
    (O(N C^2) per sequence: the N x N reference cannot be materialised at this size).  The oracle restatement itself
    runs on the GPU here only because fp64 matmuls over 1.6 M tokens take minutes on the host; it is the same
    oracle/tpavi_oracle.py code, pinned to the reference's golden vectors by tests/test_oracle.py."""
    clips, Fr, V, h, w, C = 8, 16, 4, 28, 28, 256
    B = clips * Fr
    pg = O.init_params(C, seed=61, randomize_affine=True)
    pl = O.init_params(C, seed=62, randomize_affine=True)
    gen = torch.Generator().manual_seed(63)
    f4 = [torch.randn(B, C, h, w, generator=gen).to(torch.bfloat16) for _ in range(V)]
    cl = [torch.randn(B, 5, h, w, generator=gen) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=gen) for _ in range(V)]
    dz = torch.randn(B, V, h, w, C, generator=gen).to(torch.bfloat16)
    f = _build(C, pg, pl)
    f4d = [t.to(DEV).requires_grad_(True) for t in f4]
    cld, ctd = [t.to(DEV) for t in cl], [t.to(DEV) for t in ct]
    dzd = dz.to(DEV).permute(0, 4, 1, 2, 3)
    params = [p for p in f.parameters() if p.requires_grad]
    holder = {}

    def compute():
        for t in f4d:
            t.grad = None
        for p in params:
            p.grad = None
        holder["out"] = f.forward_stacked(f4d, cld, ctd)
        holder["out"].backward(dzd)

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            compute()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        compute()
    # poison what the graph writes, then replay twice: the compared values are those of the LAST replay
    holder["out"].detach().zero_()
    for t in f4d:
        t.grad.detach().zero_()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    out = holder["out"]

    # ---- oracle: closed form in fp64
    with torch.no_grad():
        x = [t.to(DEV, torch.float64) for t in f4]
        xg, xl = O.gate_concat(x, [t.to(DEV, torch.float64) for t in cl], [t.to(DEV, torch.float64) for t in ct])
        gates = [O.gate_from_logits(a.to(DEV, torch.float64), b.to(DEV, torch.float64)) for a, b in zip(cl, ct)]
        dz64 = dzd.to(torch.float64)
        p64g = {k: (v.to(DEV, torch.float64) if v.is_floating_point() else v.to(DEV)) for k, v in pg.items()}
        p64l = {k: (v.to(DEV, torch.float64) if v.is_floating_point() else v.to(DEV)) for k, v in pl.items()}
        zg, dxg, gg, _ = O.tpavi_dot_closed_form(xg, dz64, p64g)
        del xg
        zl, dxl, gl, _ = O.tpavi_dot_closed_form(xl, dz64, p64l)
        del xl
        ref_out = zg + zl
        del zg, zl
        for v in range(V):
            assert_close(f"out:{v}", out[:, :, v], ref_out[:, :, v], BF16_TOL)
            assert_close(f"df4:{v}", f4d[v].grad, dxg[:, :, v] + gates[v] * dxl[:, :, v], BF16_TOL)
        for mod, ref in ((f.global_attn, gg), (f.local_attn, gl)):
            for k, p in mod.named_parameters():
                if not k.startswith("align_channel"):
                    assert_close("grad:" + k, p.grad, ref[k], grad_tol(k), abs_floor=1e-3)


def test_reference_checkpoint_inference(tmp_path):
    """a checkpoint in the reference trainer's wire format (R/main.py:857-872 writes
    {'network': model.module.state_dict()}; Trainer.test re-prefixes the keys with 'module.', main.py:454-457) is
    loaded from disk into the B200 path and run in inference mode (eval BatchNorm = running statistics, applied as a
    per-channel affine inside the fused LayerNorm pass: no statistics kernel, nothing to un-fold), against the oracle."""
    B, C, V, h, w = 3, 256, 4, 28, 28
    pg = O.init_params(C, seed=71, randomize_affine=True)
    pl = O.init_params(C, seed=72, randomize_affine=True)
    for p in (pg, pl):                       # a trained network has non-trivial running statistics
        p["W_z.1.running_mean"] = torch.randn(C) * 0.2
        p["W_z.1.running_var"] = torch.rand(C) + 0.5
        p["W_z.1.num_batches_tracked"] = torch.tensor(1234)
    net = {"module.global_attn." + k: v for k, v in pg.items()}
    net.update({"module.local_attn." + k: v for k, v in pl.items()})
    net["module.layer1.1.0.conv1.weight"] = torch.randn(8, 8, 1, 1)       # the rest of the network is ignored
    net["module.classifier.1.4.bias"] = torch.randn(5)
    path = tmp_path / "net_00099.pth"
    torch.save({"network": net}, path)

    f = GlobalLocalFusion(in_channels=C)
    f.load_reference_checkpoint(torch.load(path, map_location="cpu"), strict=True)
    f = f.to(DEV).eval()
    gen = torch.Generator().manual_seed(73)
    f4 = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    cl = [torch.randn(B, 5, h, w, generator=gen) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=gen) for _ in range(V)]
    ref = O.global_local_fusion(f4, cl, ct, pg, pl, training=False)
    before = f.global_attn.W_z[1].running_mean.clone()
    with torch.no_grad():
        out = f.forward_stacked([t.to(DEV, torch.bfloat16) for t in f4], [t.to(DEV) for t in cl], [t.to(DEV) for t in ct])
    torch.cuda.synchronize()
    for v in range(V):
        assert_close(f"out:{v}", out[:, :, v], ref[v], BF16_TOL)
    assert torch.equal(f.global_attn.W_z[1].running_mean, before)
    assert int(f.global_attn.W_z[1].num_batches_tracked) == 1234


def test_forward_parts_combined_segmentation_and_cycle_gradients():
    """One backward that carries BOTH a gradient for the fused sum (segmentation loss) and one for the MGFM part alone
    (cycle loss on its spatial sums, R/main.py:229-237): the fused node then runs the per-block LayerNorm backward."""
    B, C, V, h, w = 4, 256, 4, 14, 14
    pg = O.init_params(C, seed=81, randomize_affine=True)
    pl = O.init_params(C, seed=82, randomize_affine=True)
    gen = torch.Generator().manual_seed(83)
    f4 = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    cl = [torch.randn(B, 5, h, w, generator=gen) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=gen) for _ in range(V)]
    dseg = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    wsum = torch.randn(B, C, generator=gen)
    f4o = [t.clone().requires_grad_(True) for t in f4]

    def leaf(p):
        return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k not in O.BUFFER_KEYS else v.clone())
                for k, v in p.items()}
    qg, ql = leaf(pg), leaf(pl)
    xg, xl = O.gate_concat(f4o, cl, ct)
    zg, zl = O.tpavi_forward(xg, qg), O.tpavi_forward(xl, ql)
    loss = sum(((zg[:, :, i] + zl[:, :, i]) * dseg[i]).sum() + 0.5 * (zg[:, :, i].sum(dim=(2, 3)) * wsum).sum()
               for i in range(V))
    loss.backward()
    f = _build(C, pg, pl)
    f4d = [t.to(DEV, torch.bfloat16).requires_grad_(True) for t in f4]
    keys = [str(i) for i in range(V)]
    fus, glob, loc = f.forward_parts(dict(zip(keys, f4d)), dict(zip(keys, [t.to(DEV) for t in cl])),
                                     dict(zip(keys, [t.to(DEV) for t in ct])), need_local=False)
    assert loc is None
    loss_d = sum((fus[k].float() * dseg[i].to(DEV)).sum() + 0.5 * (glob[k].float().sum(dim=(2, 3)) * wsum.to(DEV)).sum()
                 for i, k in enumerate(keys))
    loss_d.backward()
    torch.cuda.synchronize()
    for i in range(V):
        assert_close(f"df4:{i}", f4d[i].grad, f4o[i].grad, BF16_TOL)
    for mod, ref in ((f.global_attn, qg), (f.local_attn, ql)):
        for k, p in mod.named_parameters():
            if not k.startswith("align_channel") and k != "W_z.0.bias":
                assert_close("grad:" + k, p.grad, ref[k].grad, grad_tol(k), abs_floor=1e-3)


@pytest.mark.parametrize("h,w,Cn", [(28, 28, 256), (5, 7, 64), (3, 33, 128)])
@pytest.mark.parametrize("src", [torch.bfloat16, torch.float32])
def test_views_to_tokens_is_the_per_view_transposition(h, w, Cn, src):
    """The dict-keyed backward gathers one gradient per view into the token-major stack (the transposition
    ours.py:1819-1820 builds forward with permute + cat).  Bit-exact: it is a copy with one round-to-nearest cast."""
    from glfusion_b200.fusion import views_to_tokens
    B, V = 3, 4
    gen = torch.Generator(device="cpu").manual_seed(11)
    views = [torch.randn(B, Cn, h, w, generator=gen).to(DEV).to(src) for _ in range(V)]
    views[1] = views[1].contiguous(memory_format=torch.channels_last)            # channels-last strides
    big = torch.randn(B, V, h, w, Cn, generator=gen).to(DEV).to(src)
    views[2] = big[:, 2].permute(0, 3, 1, 2)                                      # a view into a token-major stack
    views[3] = None                                                               # an output nobody differentiated
    if Cn == 128:                                                                 # the gradient of .sum((2, 3)): a broadcast
        views[0] = torch.randn(B, Cn, 1, 1, generator=gen).to(DEV).to(src).expand(B, Cn, h, w)
    if Cn == 64:                                                                  # rows that are not 16-byte aligned
        odd = torch.randn(B, h, w, Cn + 1, generator=gen).to(DEV).to(src)
        views[0] = odd[..., 1:].permute(0, 3, 1, 2)
    out = torch.full((B, V, h, w, Cn), 7.0, dtype=torch.bfloat16, device=DEV)
    assert views_to_tokens(views, out)
    for v, g in enumerate(views):
        want = torch.zeros(B, h, w, Cn, device=DEV) if g is None else g.permute(0, 2, 3, 1).float()
        assert torch.equal(out[:, v].float(), want.to(torch.bfloat16).float()), v
    # cases the kernel does not take are reported, not mangled
    assert not views_to_tokens([v.half() if v is not None else None for v in views], out)
    assert not views_to_tokens([None] * V, out)
    if w > 1:
        assert not views_to_tokens([views[0].transpose(2, 3).contiguous().transpose(2, 3)] + views[1:], out)   # h-major


@pytest.mark.parametrize("io", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("Cn,h,w", [(256, 28, 28), (128, 5, 7), (2048, 6, 6)])
def test_channels_last_hand_off_matches_nchw(io, Cn, h, w):
    """backbones / heads that run channels_last hand f4 over as rows of C channels.  The fused
    node then takes the row kernels (glf_gate_concat_cl_*, no transposition) and returns df4 channels_last; outputs and
    every gradient agree with the NCHW path on the same values."""
    torch.manual_seed(0)
    B, V = 3, 3
    fus = GlobalLocalFusion(Cn).to(DEV)
    from bench import randomize_affine_
    randomize_affine_(fus.global_attn, 5)
    randomize_affine_(fus.local_attn, 6)
    gen = torch.Generator().manual_seed(4)
    base = [torch.randn(B, Cn, h, w, generator=gen).to(DEV).to(io) for _ in range(V)]
    cl = [torch.randn(B, 4, h, w, generator=gen).to(DEV) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=gen).to(DEV) for _ in range(V)]
    dz = torch.randn(B, V, h, w, Cn, generator=gen).to(DEV).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
    res = []
    for fmt in ("nchw", "channels_last"):
        f4 = [t.clone() if fmt == "nchw" else t.clone().contiguous(memory_format=torch.channels_last) for t in base]
        lc = [t.clone().requires_grad_(True) for t in cl]
        lt = [t.clone().requires_grad_(True) for t in ct]
        for t in f4:
            t.requires_grad_(True)
        for p in fus.parameters():
            p.grad = None
        out = fus.forward_stacked(f4, lc, lt)
        out.backward(dz)
        if fmt == "channels_last":
            for t in f4:     # the gradient comes back in the view's own memory format
                assert t.grad.is_contiguous(memory_format=torch.channels_last) and t.grad.dtype == io
        res.append((out.detach().float(), [t.grad.float() for t in f4], [t.grad for t in lc], [t.grad for t in lt],
                    [p.grad.clone() for p in fus.parameters() if p.grad is not None]))
    (o0, g0, c0, t0, p0), (o1, g1, c1, t1, p1) = res
    # forward: the same products, only the read pattern differs; not bit-equal because fewer than 96 sequences split the
    # token contractions with fp32 atomics (run-to-run summation order)
    assert (o0 - o1).abs().max().item() <= 1e-2 * o0.abs().max().item()
    for a, b in zip(g0, g1):
        assert (a - b).abs().max().item() <= 8e-3 * a.abs().max().item()      # one bf16 rounding of dxg + a * dxl
    for a, b in zip(c0 + t0, c1 + t1):
        assert (a - b).abs().max().item() <= 2e-2 * max(a.abs().max().item(), 1e-6) + 1e-9   # sums of bf16 dxl products
    for a, b in zip(p0, p1):                         # the blocks see identical inputs
        assert (a - b).abs().max().item() <= 2e-2 * max(a.abs().max().item(), 1e-6) + 1e-7   # bf16 chain, run-to-run order


@pytest.mark.parametrize("h,w", [(4, 7), (28, 28), (5, 7)])
def test_dict_api_backward_reads_per_view_gradients_in_place(h, w):
    """The dict-keyed call hands the backward one gradient per view.  When they are rows of C channels (channels_last, or
    views of one token-major buffer) and h*w is a multiple of the LayerNorm pass's row tile, the fused LayerNorm backward
    reads them where they lie (glf_fusion_ln_bwd_views); NCHW gradients (and 5x7, which does not tile) are gathered
    first.  All three deliveries of the same gradient values give the same input and parameter gradients."""
    torch.manual_seed(0)
    B, V, Cn = 3, 2, 128
    fus = GlobalLocalFusion(Cn).to(DEV)
    from bench import randomize_affine_
    randomize_affine_(fus.global_attn, 7)
    randomize_affine_(fus.local_attn, 8)
    keys = ["1", "3"]
    gen = torch.Generator().manual_seed(6)
    f4v = [torch.randn(B, Cn, h, w, generator=gen).to(DEV).to(torch.bfloat16) for _ in keys]
    cl = {k: torch.randn(B, 3, h, w, generator=gen).to(DEV) for k in keys}
    ct = {k: torch.randn(B, 1, h, w, generator=gen).to(DEV) for k in keys}
    dzs = torch.randn(B, V, h, w, Cn, generator=gen).to(DEV).to(torch.bfloat16)      # token-major stack of gradients
    res = []
    for delivery in ("stack_views", "channels_last", "nchw"):
        f4 = {k: t.clone().requires_grad_(True) for k, t in zip(keys, f4v)}
        for p in fus.parameters():
            p.grad = None
        out = fus(f4, cl, ct)
        if delivery == "stack_views":
            gs = [dzs[:, i].permute(0, 3, 1, 2) for i in range(V)]
        elif delivery == "channels_last":
            gs = [dzs[:, i].permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last) for i in range(V)]
        else:
            gs = [dzs[:, i].permute(0, 3, 1, 2).contiguous() for i in range(V)]
        torch.autograd.backward([out[k] for k in keys], gs)
        res.append(([f4[k].grad.float() for k in keys], [p.grad.clone() for p in fus.parameters() if p.grad is not None]))
    for other in res[1:]:
        for a, b in zip(res[0][0] + res[0][1], other[0] + other[1]):
            assert (a - b).abs().max().item() <= 2e-2 * max(a.abs().max().item(), 1e-6) + 1e-7


@pytest.mark.parametrize("C,V,h,w", [(512, 2, 9, 8), (2048, 3, 6, 6), (768, 2, 5, 7), (1024, 2, 10, 8), (2048, 3, 8, 8)])
def test_wide_channel_fused_node_against_oracle(C, V, h, w):
    """Channel counts above 256: the fused node still runs both blocks' LayerNorms in one pass forward (sliced rows,
    ln_pair_fwd_ring_kernel) and returns the MGFM part beside the sum; the gate / concat kernels cut rows wider than 512
    channels into slabs (h*w a multiple of 8: TMA transposition kernels; otherwise the SIMT ones); seeded oracle
    comparison of outputs, parts and every gradient."""
    B = 2
    pg = O.init_params(C, seed=41, randomize_affine=True)
    pl = O.init_params(C, seed=42, randomize_affine=True)
    gen = torch.Generator().manual_seed(43)
    f4 = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    cl = [torch.randn(B, 5, h, w, generator=gen) for _ in range(V)]
    ct = [torch.randn(B, 1, h, w, generator=gen) for _ in range(V)]
    do = [torch.randn(B, C, h, w, generator=gen) for _ in range(V)]
    outs, df4, dcls, dctr, gg, gl = O.fusion_fwd_bwd(f4, cl, ct, do, {k: v.clone() for k, v in pg.items()},
                                                     {k: v.clone() for k, v in pl.items()})
    f = _build(C, pg, pl)
    keys = [str(i) for i in range(V)]
    f4d = {k: t.to(DEV, torch.bfloat16).requires_grad_(True) for k, t in zip(keys, f4)}
    cld = {k: t.to(DEV).requires_grad_(True) for k, t in zip(keys, cl)}
    ctd = {k: t.to(DEV).requires_grad_(True) for k, t in zip(keys, ct)}
    fus, glob, loc = f.forward_parts(f4d, cld, ctd)
    torch.autograd.backward([fus[k] for k in keys], [t.to(DEV, torch.bfloat16) for t in do])
    torch.cuda.synchronize()
    for v, k in enumerate(keys):
        assert_close(f"out:{v}", fus[k], outs[v], BF16_TOL)
        assert_close(f"global + local:{v}", glob[k].float() + loc[k].float(), outs[v], BF16_TOL)
        assert_close(f"df4:{v}", f4d[k].grad, df4[v], BF16_TOL)
        assert_close(f"dctr:{v}", ctd[k].grad, dctr[v], 4e-2)
    for mod, ref in ((f.local_attn, gl), (f.global_attn, gg)):
        for k, p in mod.named_parameters():
            if not k.startswith("align_channel"):
                assert_close("grad:" + k, p.grad, ref[k], grad_tol(k), abs_floor=1e-3)
