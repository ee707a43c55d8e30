"""The C-ABI library loads without a GPU and exports exactly what include/glfusion.h declares; host-only entry points
(sizes, error reporting) behave.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT

import glfusion_b200
from glfusion_b200 import _lib as L


def _declared():
    src = open(os.path.join(ROOT, "include", "glfusion.h")).read()
    return sorted(set(re.findall(r"GLF_API\s+[\w\s\*]+?\b(glf_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = glfusion_b200.load_library()
    names = _declared()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(L.EXPORTS) == names
    assert lib.glf_version() == 100


def _desc(**kw):
    d = L.GlfDesc()
    d.B, d.T, d.H, d.W, d.C, d.Ci = 2, 4, 28, 28, 256, 128
    d.mode, d.io_dtype, d.x_layout, d.dz_layout = L.MODE_DOT, L.DTYPE_BF16, L.LAYOUT_TOKEN, L.LAYOUT_TOKEN
    d.training, d.bn_layer = 1, 1
    d.eps_bn = d.eps_ln = 1e-5
    d.momentum = 0.1
    d.reserved[1] = kw.pop("algo", 1)      # mode='dot' algorithm: 1 = token-space form, 2 = Gram form, 0 = auto
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def test_sizes_are_host_only_and_scale_with_tokens():
    lib = glfusion_b200.load_library()
    s1, s2 = L.GlfSizes(), L.GlfSizes()
    assert lib.glf_tpavi_sizes(C.byref(_desc()), C.byref(s1)) == 0
    assert lib.glf_tpavi_sizes(C.byref(_desc(B=4)), C.byref(s2)) == 0
    rows = 2 * 4 * 28 * 28
    # saved holds at least P [rows, 3Ci] and U [rows, C] in bf16
    assert s1.saved_bytes >= rows * (3 * 128 + 256) * 2
    assert s2.saved_bytes > 1.9 * s1.saved_bytes * 0.98
    assert s1.ws_bwd_bytes >= rows * (256 * 2 + 3 * 128) * 2
    assert s1.saved_bytes % 256 == 0 and s1.ws_fwd_bytes % 256 == 0 and s1.ws_bwd_bytes % 256 == 0
    # the Gram form keeps U and per-sequence [C x C] matrices, never the per-token projections; auto picks it (N >= 5 C)
    sg, sa = L.GlfSizes(), L.GlfSizes()
    assert lib.glf_tpavi_sizes(C.byref(_desc(algo=2)), C.byref(sg)) == 0
    assert lib.glf_tpavi_sizes(C.byref(_desc(algo=0)), C.byref(sa)) == 0
    assert rows * 256 * 2 <= sg.saved_bytes < s1.saved_bytes
    assert sa.saved_bytes == sg.saved_bytes and sa.ws_bwd_bytes == sg.ws_bwd_bytes
    assert lib.glf_tpavi_sizes(C.byref(_desc(algo=2, mode=L.MODE_EMBEDDED)), C.byref(sg)) < 0
    # NCTHW / fp32 inputs need a packed copy of x as well
    s3 = L.GlfSizes()
    assert lib.glf_tpavi_sizes(C.byref(_desc(x_layout=L.LAYOUT_NCTHW, io_dtype=L.DTYPE_F32)), C.byref(s3)) == 0
    assert s3.saved_bytes >= s1.saved_bytes + rows * 256 * 2
    # the fp32-exact arm keeps fp32 activations plus three bf16 limb planes of the GEMM operands
    s4 = L.GlfSizes()
    assert lib.glf_tpavi_sizes(C.byref(_desc(precision=L.PRECISION_F32X3, io_dtype=L.DTYPE_F32)), C.byref(s4)) == 0
    assert s4.saved_bytes >= rows * (256 * 4 + 3 * 256 * 2 + 3 * 384 * 2)


@pytest.mark.parametrize("kw,frag", [
    (dict(B=0), "empty"),
    (dict(C=100), "multiples of 8"),
    (dict(C=4096, Ci=2048), "2048"),
    (dict(mode=3), "mode"),
    (dict(precision=5), "precision"),
    (dict(precision=1, mode=1), "F32X3"),
    (dict(io_dtype=7), "io_dtype"),
])
def test_invalid_descriptors_are_errors_with_messages(kw, frag):
    lib = glfusion_b200.load_library()
    s = L.GlfSizes()
    rc = lib.glf_tpavi_sizes(C.byref(_desc(**kw)), C.byref(s))
    assert rc < 0
    assert frag.lower() in lib.glf_last_error().decode().lower()
    with pytest.raises(L.GlfError):
        L.check(rc)


@pytest.mark.parametrize("C_,T,H,W", [(256, 4, 28, 28), (256, 4, 14, 14), (256, 4, 16, 20), (512, 4, 28, 28),
                                      (512, 8, 24, 24), (1024, 4, 40, 40), (128, 2, 16, 20), (2048, 3, 28, 28)])
def test_automatic_dot_algorithm_follows_the_documented_rule(C_, T, H, W):
    """reserved[1] = 0: Gram form iff N >= 3 C at C = 256 (the width with the per-sequence chain kernels), 5 C below,
    8 C above — observable through the blob sizes; bench.py mirrors the same rule for its FLOP / launch accounting."""
    import bench
    lib = glfusion_b200.load_library()
    N = T * H * W
    thr = 3 if C_ == 256 else (8 if C_ > 256 else 5)
    expect = 2 if N >= thr * C_ else 1
    kw = dict(C=C_, Ci=C_ // 2, T=T, H=H, W=W)
    sa, se, so = L.GlfSizes(), L.GlfSizes(), L.GlfSizes()
    assert lib.glf_tpavi_sizes(C.byref(_desc(algo=0, **kw)), C.byref(sa)) == 0
    assert lib.glf_tpavi_sizes(C.byref(_desc(algo=expect, **kw)), C.byref(se)) == 0
    assert lib.glf_tpavi_sizes(C.byref(_desc(algo=3 - expect, **kw)), C.byref(so)) == 0
    assert (sa.saved_bytes, sa.ws_bwd_bytes) == (se.saved_bytes, se.ws_bwd_bytes)
    assert (sa.saved_bytes, sa.ws_bwd_bytes) != (so.saved_bytes, so.ws_bwd_bytes)
    if (T, H, W) == (bench.V, bench.HH, bench.WW):
        assert bench.dot_algorithm(C_) == ("gram" if expect == 2 else "token")
