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
    assert L.cgp_version() == 100


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
