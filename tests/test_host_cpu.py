"""CPU-only checks: the C-ABI library loads and exports every symbol include/convgp_b200.h declares,
descriptor validation works without a GPU, the host-side classes keep the reference's surface, and
the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "convgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from convgp_b200 import _lib
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.EXPORTS) == names
    assert L.cgp_version() == 200


def test_ctypes_argtypes_match_header_arity():
    """Every prototype in include/convgp_b200.h has as many parameters as the ctypes binding declares (a mismatch
    only shows up as a TypeError at the first call -- on the GPU box)."""
    from convgp_b200 import _lib
    L = _lib.lib()
    src = open(os.path.join(ROOT, "include", "convgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(cgp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src)
    assert len(protos) >= 14
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        at = getattr(L, name).argtypes
        if n == 0 and at is None:
            continue
        assert at is not None, name + ": no argtypes declared"
        assert len(at) == n, (name, len(at), n)


def test_descriptor_validation_without_gpu():
    from convgp_b200 import _lib
    s = _lib.Spec(torch.float64, 28, 28, 1, 5, 5, _lib.LAYOUT_PLANAR, _lib.OUT_SUM, 1, False, False)
    assert (s.P, s.D) == (576, 25)
    s = _lib.Spec(torch.float32, 32, 32, 3, 5, 5, _lib.LAYOUT_COLOUR_PATCH, _lib.OUT_SUM, 3, False, True,
                  np.array([[(d % 3) == g for d in range(75)] for g in range(3)], dtype=np.uint8))
    assert (s.P, s.D) == (784, 75)
    s = _lib.Spec(torch.float64, 32, 32, 3, 5, 5, _lib.LAYOUT_PLANAR, _lib.OUT_PER_CHANNEL, 1, False, False)
    assert (s.P, s.D) == (2352, 25)
    with pytest.raises(_lib.CgpError) as e:
        _lib.Spec(torch.float64, 3, 3, 1, 5, 5, 0, 0, 1, False, False)
    assert e.value.status == -1
    with pytest.raises(_lib.CgpError):  # per-channel output with colour-patch layout
        _lib.Spec(torch.float64, 8, 8, 3, 3, 3, _lib.LAYOUT_COLOUR_PATCH, _lib.OUT_PER_CHANNEL, 1, False, False)
    with pytest.raises(TypeError):
        _lib.Spec(torch.float16, 8, 8, 1, 3, 3, 0, 0, 1, False, False)
    # specialised configs need no workspace, generic ones do
    L = _lib.lib()
    fast = _lib.Spec(torch.float64, 28, 28, 1, 5, 5, 0, 0, 1, False, False)
    gen = _lib.Spec(torch.float64, 28, 28, 1, 4, 4, 0, 0, 1, False, False)
    assert L.cgp_workspace_bytes(fast.ref(), 10, 5) == 0
    assert L.cgp_workspace_bytes(gen.ref(), 10, 5) >= 2 * 10 * 625 * 16 * 8
    assert L.cgp_status_string(-6).decode() == "matrix not positive definite"


def test_null_and_bad_arguments_return_status():
    from convgp_b200 import _lib
    L = _lib.lib()
    s = _lib.Spec(torch.float64, 28, 28, 1, 5, 5, 0, 0, 1, False, False)
    # NULL X with N > 0 -> CGP_ERR_BAD_POINTER before any CUDA call
    assert L.cgp_kdiag_fwd(s.ref(), 4, None, None, None, None, None, None, 0, None) == -3
    assert b"must not be NULL" in L.cgp_last_error()
    assert L.cgp_kuf_fwd(s.ref(), -1, 3, None, None, None, None, None, None, None, 0, None) == -1
    assert L.cgp_kuf_fwd(s.ref(), 0, 3, None, None, None, None, None, None, None, 0, None) == 0  # empty batch
    v = ctypes.c_double()
    assert L.cgp_peak_probe(99, ctypes.byref(v)) == -1


def test_kernel_surface_matches_reference():
    from convgp_b200 import convkernels as ck, kernels as bk, misvgp
    k = ck.WeightedConv(bk.RBF(25), [28, 28], [5, 5]) + bk.White(1, 1e-3)
    conv = k.kern_list[0]
    assert (conv.num_patches, conv.patch_len, conv.colour_channels) == (576, 25, 1)
    assert conv.W.shape == (576,) and np.all(conv.W.value == 1.0)
    for name in ("K", "Kdiag", "Kzx", "Kzz", "init_inducing", "compute_patches", "compute_Kzx", "compute_Kdiag",
                 "compute_K_symm", "compute_Kzz"):
        assert callable(getattr(conv, name)), name
    assert abs(k.white.variance.value - 1e-3) < 1e-12
    conv.basekern.variance = 0.25
    conv.basekern.lengthscales = 2.34
    assert abs(conv.basekern.variance.value - 0.25) < 1e-12 and abs(conv.basekern.lengthscales.value - 2.34) < 1e-12
    c3 = ck.Conv(bk.RBF(25), [32, 32], [5, 5], colour_channels=3)
    cp = ck.WeightedColourPatchConv(bk.RBF(75), [32, 32], [5, 5], colour_channels=3)
    mc = misvgp.WeightedMultiChannelConvGP(bk.RBF(25), [32, 32], [5, 5], 3)
    assert (c3.num_patches, c3.patch_len) == (2352, 25)
    assert (cp.num_patches, cp.patch_len, cp.W.shape) == (784, 75, (784,))
    assert (mc.num_patches, mc.W.shape, mc.inducing_outputs) == (2352, (3, 784), 3)
    # free-state round trip (opt_tools/helpers.py:108-117 drives models through this vector)
    x = k.get_free_state()
    k.set_state(x * 1.0)
    assert np.allclose(k.get_free_state(), x)
    k.fixed = True
    assert k.get_free_state().size == 0
    with pytest.raises(NotImplementedError):
        ck.Conv(bk.White(25), [28, 28], [5, 5])._structure()


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    from convgp_b200 import _lib, convkernels as ck, kernels as bk
    k = ck.Conv(bk.RBF(4), [3, 3], [2, 2])
    with pytest.raises(_lib.CgpError):
        k.compute_Kdiag(np.ones((2, 9)))
    with pytest.raises(_lib.CgpError):
        k.compute_patches(np.ones((2, 9)))


def test_shard_bounds_cover_all_rows():
    from convgp_b200.parallel import shard_bounds
    for n, w in ((200, 8), (30, 8), (7, 4), (3, 8)):
        got = [shard_bounds(n, w, r) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
        sizes = [hi - lo for lo, hi in got]
        assert max(sizes) - min(sizes) <= 1
    assert [hi - lo for lo, hi in (shard_bounds(30, 8, r) for r in range(8))] == [4, 4, 4, 4, 4, 4, 3, 3]


def test_argument_validation_raises_before_any_launch():
    """ops.validate: everything that crosses the C ABI as a bare pointer is shape- and dtype-checked on the host
    (the reference's TF graph raises a shape error in these cases); no GPU needed to see the error."""
    from convgp_b200 import _lib, ops
    s = _lib.Spec(torch.float64, 28, 28, 1, 5, 5, 0, 0, 1, False, False)
    X, Z, W = torch.zeros(3, 784, dtype=torch.float64), torch.zeros(4, 25, dtype=torch.float64), torch.ones(576, dtype=torch.float64)
    v, l = torch.ones(1, dtype=torch.float64), torch.ones(1, dtype=torch.float64)
    ops.validate(s, X=X, Z=Z, W=W, variance=v, lengthscales=l)
    with pytest.raises(ValueError, match="H\\*W\\*C"):   # CIFAR-sized rows into a 28x28x1 kernel
        ops.validate(s, X=torch.zeros(3, 3072, dtype=torch.float64))
    with pytest.raises(ValueError, match="patch_len"):
        ops.validate(s, Z=torch.zeros(4, 9, dtype=torch.float64))
    with pytest.raises(ValueError, match="num_patches"):
        ops.validate(s, W=torch.ones(100, dtype=torch.float64))
    with pytest.raises(TypeError, match="float_type"):
        ops.validate(s, X=X.float())
    with pytest.raises(ValueError, match="variance"):
        ops.validate(s, variance=torch.ones(3, dtype=torch.float64))
    with pytest.raises(ValueError, match="lengthscales"):
        ops.validate(s, lengthscales=torch.ones(25, dtype=torch.float64))
    # through the public class: a wrong-length W assigned to the Param is caught at the call, not in the kernel
    from convgp_b200 import convkernels as ck, kernels as bk
    k = ck.WeightedConv(bk.RBF(25), [28, 28], [5, 5])
    k.W = np.ones(7)
    with pytest.raises(ValueError, match="num_patches"):
        k.Kzx(Z.to(k.W().device), X.to(k.W().device))


def test_options_api_and_family_query():
    from convgp_b200 import _lib
    L = _lib.lib()
    f64 = _lib.Spec(torch.float64, 28, 28, 1, 5, 5, 0, 0, 1, False, False)
    f32 = _lib.Spec(torch.float32, 28, 28, 1, 5, 5, 0, 0, 1, False, False)
    gen = _lib.Spec(torch.float64, 28, 28, 1, 4, 4, 0, 0, 1, False, False)
    assert L.cgp_kernel_family(f64.ref(), 200, 750, 0) == b"dmma64"
    for which in (0, 1, 2):  # Kuf forward / backward, Kdiag forward: the two-fp16-term tcgen05 family by default
        assert L.cgp_kernel_family(f32.ref(), 200, 750, which) == b"umma16"
    # every float32 BASELINE configuration is on the tcgen05 kind::f16 family (cfg 1: 3x3 patches, M = 16; cfg 2: M = 50;
    # cfg 3 / 4: M = 750; cfg 5 below)
    cfg1 = _lib.Spec(torch.float32, 28, 28, 1, 3, 3, 0, 0, 1, False, False)
    for spec, n, m in ((cfg1, 100, 16), (f32, 100, 50), (f32, 200, 750)):
        assert [L.cgp_kernel_family(spec.ref(), n, m, w) for w in (0, 1, 2, 3)] == \
            [b"umma16", b"umma16", b"umma16", b"saved-row-sums"]
    # planar colour images (Conv with colour_channels = 3: one pairing domain) take the same family when the
    # shared-memory plan fits
    rgb = _lib.Spec(torch.float32, 16, 16, 3, 5, 5, _lib.LAYOUT_PLANAR, _lib.OUT_SUM, 1, False, False)
    assert all(L.cgp_kernel_family(rgb.ref(), 30, 100, w) == b"umma16" for w in (0, 1, 2))
    # config 5: colour patches with one RBF per channel (cifar.py:33-41) = one launch whose CTAs split over the groups
    mask = np.array([[(dd % 3) == gg for dd in range(75)] for gg in range(3)], dtype=np.uint8)
    cifar = _lib.Spec(torch.float32, 32, 32, 3, 5, 5, _lib.LAYOUT_COLOUR_PATCH, _lib.OUT_SUM, 3, False, True, mask)
    assert [L.cgp_kernel_family(cifar.ref(), 30, 1000, w) for w in (0, 1, 2, 3)] == \
        [b"umma16", b"umma16", b"umma16", b"saved-row-sums"]
    assert L.cgp_kdiag_saved_elems(cifar.ref(), 30) == 30 * (784 + 2 * 3)
    # per-channel outputs (WeightedMultiChannelConvGP): one group per channel in the same launches; its Kdiag backward
    # (no saved-row-sum variant for (N, C) outputs) recomputes on the mma.sync family
    multi = _lib.Spec(torch.float32, 16, 16, 3, 5, 5, _lib.LAYOUT_PLANAR, _lib.OUT_PER_CHANNEL, 1, False, False)
    assert [L.cgp_kernel_family(multi.ref(), 30, 100, w) for w in (0, 1, 2, 3)] == \
        [b"umma16", b"umma16", b"umma16", b"hmma32"]
    _lib.set_option("kuf_fwd_f32", 2)
    _lib.set_option("kuf_bwd_f32", 2)
    try:
        assert L.cgp_kernel_family(f32.ref(), 200, 750, 0) == b"umma32"
        assert L.cgp_kernel_family(f32.ref(), 200, 750, 1) == b"umma32"
    finally:
        _lib.set_option("kuf_fwd_f32", -1)
        _lib.set_option("kuf_bwd_f32", -1)
    assert L.cgp_kernel_family(gen.ref(), 200, 750, 2) == b"generic"
    _lib.set_option("kuf_fwd", 0)
    try:
        assert L.cgp_kernel_family(f64.ref(), 200, 750, 0) == b"scalar"
    finally:
        _lib.set_option("kuf_fwd", -1)
    assert L.cgp_kernel_family(f64.ref(), 200, 750, 0) == b"dmma64"
    assert L.cgp_set_option(b"no_such_option", 1) == -1


def test_library_reads_no_environment():
    src = open(os.path.join(ROOT, "convgp_b200", "csrc", "api.cu")).read()
    assert "getenv" not in src
