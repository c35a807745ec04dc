"""ctypes binding of libconvgp_b200.so (the C ABI declared in include/convgp_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libconvgp_b200.so")

CGP_F32, CGP_F64 = 0, 1
LAYOUT_PLANAR, LAYOUT_COLOUR_PATCH = 0, 1
OUT_SUM, OUT_PER_CHANNEL = 0, 1

EXPORTS = [
    "cgp_version", "cgp_last_error", "cgp_status_string", "cgp_geometry", "cgp_workspace_bytes", "cgp_patches",
    "cgp_kuf_fwd", "cgp_kuf_bwd", "cgp_kdiag_fwd", "cgp_kdiag_bwd", "cgp_kuu_fwd", "cgp_kuu_bwd",
    "cgp_kfull_fwd", "cgp_peak_probe", "cgp_kdiag_saved_elems", "cgp_kdiag_fwd_saved", "cgp_kdiag_bwd_saved",
    "cgp_kernel_family", "cgp_set_option", "cgp_get_option",
    "cgp_comm_create", "cgp_comm_handle", "cgp_comm_connect", "cgp_comm_local", "cgp_comm_result",
    "cgp_comm_allreduce", "cgp_comm_status", "cgp_comm_destroy",
    "cgp_gemm", "cgp_potrf_inv_workspace", "cgp_potrf_inv", "cgp_potrf_check", "cgp_project_fwd", "cgp_project_bwd", "cgp_prob_is_largest", "cgp_bernoulli_ve",
]


class CgpError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("convgp_b200: %s (status %d)" % (message, status))
        self.status = status


class KernelDesc(ctypes.Structure):
    _fields_ = [
        ("dtype", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("C", ctypes.c_int32),
        ("ph", ctypes.c_int32), ("pw", ctypes.c_int32), ("layout", ctypes.c_int32), ("out_mode", ctypes.c_int32),
        ("n_groups", ctypes.c_int32), ("ard", ctypes.c_int32), ("round_x_f32", ctypes.c_int32),
        ("reserved", ctypes.c_int32), ("active", ctypes.c_void_p),
    ]


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("dtype", ctypes.c_int32), ("m", ctypes.c_int32), ("n", ctypes.c_int32), ("k", ctypes.c_int32),
        ("trans_a", ctypes.c_int32), ("trans_b", ctypes.c_int32), ("a_tri", ctypes.c_int32), ("b_tri", ctypes.c_int32),
        ("c_mode", ctypes.c_int32), ("accumulate", ctypes.c_int32), ("batch", ctypes.c_int32),
        ("batch_reduce", ctypes.c_int32), ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64), ("ldc", ctypes.c_int64),
        ("stride_a", ctypes.c_int64), ("stride_b", ctypes.c_int64), ("stride_c", ctypes.c_int64),
        ("alpha", ctypes.c_double),
    ]


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "convgp_b200: %s is missing -- run `python -m convgp_b200.build` (needs nvcc). "
            "There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, sz, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_size_t, ctypes.c_double
    dp = ctypes.POINTER(KernelDesc)
    L.cgp_version.restype = ctypes.c_int
    L.cgp_last_error.restype = ctypes.c_char_p
    L.cgp_status_string.restype = ctypes.c_char_p
    L.cgp_status_string.argtypes = [ctypes.c_int]
    L.cgp_geometry.argtypes = [dp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.cgp_workspace_bytes.restype = sz
    L.cgp_workspace_bytes.argtypes = [dp, i32, i32]
    L.cgp_patches.argtypes = [dp, i32, vp, vp, vp]
    L.cgp_kuf_fwd.argtypes = [dp, i32, i32] + [vp] * 6 + [vp, sz, vp]
    L.cgp_kuf_bwd.argtypes = [dp, i32, i32] + [vp] * 10 + [vp, sz, vp]
    L.cgp_kdiag_fwd.argtypes = [dp, i32] + [vp] * 5 + [vp, sz, vp]
    L.cgp_kdiag_bwd.argtypes = [dp, i32] + [vp] * 8 + [vp, sz, vp]
    L.cgp_kdiag_saved_elems.restype = sz
    L.cgp_kdiag_saved_elems.argtypes = [dp, i32]
    L.cgp_kdiag_fwd_saved.argtypes = [dp, i32] + [vp] * 6 + [sz, vp]
    L.cgp_kdiag_bwd_saved.argtypes = [dp, i32] + [vp] * 4 + [sz] + [vp] * 4
    L.cgp_kuu_fwd.argtypes = [dp, i32, vp, vp, vp, dbl, vp, vp]
    L.cgp_kuu_bwd.argtypes = [dp, i32] + [vp] * 7 + [vp]
    L.cgp_kfull_fwd.argtypes = [dp, i32, i32] + [vp] * 6 + [vp, sz, vp]
    L.cgp_peak_probe.argtypes = [i32, ctypes.POINTER(dbl)]
    L.cgp_kernel_family.restype = ctypes.c_char_p
    L.cgp_kernel_family.argtypes = [dp, i32, i32, i32]
    L.cgp_set_option.argtypes = [ctypes.c_char_p, i32]
    L.cgp_get_option.argtypes = [ctypes.c_char_p, ctypes.POINTER(i32)]
    L.cgp_comm_create.argtypes = [i32, i32, sz, ctypes.POINTER(vp)]
    L.cgp_comm_handle.argtypes = [vp, vp]
    L.cgp_comm_connect.argtypes = [vp, vp]
    L.cgp_comm_local.restype = vp
    L.cgp_comm_local.argtypes = [vp]
    L.cgp_comm_result.restype = vp
    L.cgp_comm_result.argtypes = [vp]
    L.cgp_comm_allreduce.argtypes = [vp, i32, sz, vp]
    L.cgp_comm_status.argtypes = [vp]
    L.cgp_comm_destroy.argtypes = [vp]
    i64 = ctypes.c_int64
    L.cgp_gemm.argtypes = [ctypes.POINTER(GemmDesc), vp, vp, vp, vp]
    L.cgp_potrf_inv_workspace.restype = sz
    L.cgp_potrf_inv_workspace.argtypes = [i32, i32]
    L.cgp_potrf_inv.argtypes = [i32, i32, vp, i64, vp, i64, vp, sz, vp, vp]
    L.cgp_potrf_check.argtypes = [vp, vp]
    L.cgp_project_fwd.argtypes = [i32] * 6 + [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
    L.cgp_project_bwd.argtypes = [i32] * 5 + [vp, vp, vp, vp, i64] + [vp] * 8 + [vp]
    dp = ctypes.POINTER(ctypes.c_double)
    L.cgp_prob_is_largest.argtypes = [i32, i32, i32, vp, vp, vp, i32, dp, dp, vp, vp, vp, vp]
    L.cgp_bernoulli_ve.argtypes = [i32, i64, vp, vp, vp, i32, dp, dp, vp, vp, vp, vp]
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def set_option(name, value):
    """Explicit kernel-family selection (cgp_set_option): value < 0 restores the default."""
    check(lib().cgp_set_option(name.encode(), int(value)))


def check(status):
    if status != 0:
        L = lib()
        raise CgpError(status, (L.cgp_last_error() or L.cgp_status_string(status)).decode())


def dtype_code(dtype):
    if dtype == torch.float64:
        return CGP_F64
    if dtype == torch.float32:
        return CGP_F32
    raise TypeError("convgp_b200 kernels run in float32 or float64, got %s" % dtype)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CgpError(-3, "tensor is not on a CUDA device (no CPU path exists)")
    if not t.is_contiguous():
        raise CgpError(-3, "tensor is not contiguous")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Spec:
    """Host-side description of one patch kernel: geometry + base-kernel group structure.
    Owns the ctypes descriptor and the host mask it points to."""

    def __init__(self, dtype, H, W, C, ph, pw, layout, out_mode, n_groups, ard, round_x_f32, active=None):
        self.torch_dtype = dtype
        self._mask = None
        d = KernelDesc()
        d.dtype = dtype_code(dtype)
        d.H, d.W, d.C, d.ph, d.pw = int(H), int(W), int(C), int(ph), int(pw)
        d.layout, d.out_mode = int(layout), int(out_mode)
        d.n_groups, d.ard, d.round_x_f32 = int(n_groups), int(bool(ard)), int(bool(round_x_f32))
        if active is not None:
            self._mask = np.ascontiguousarray(np.asarray(active, dtype=np.uint8))
            d.active = self._mask.ctypes.data_as(ctypes.c_void_p)
        self.desc = d
        P, D = ctypes.c_int32(), ctypes.c_int32()
        check(lib().cgp_geometry(ctypes.byref(d), ctypes.byref(P), ctypes.byref(D)))
        self.P, self.D = P.value, D.value
        if self._mask is not None and self._mask.size != n_groups * self.D:
            raise CgpError(-1, "active mask must have n_groups*D = %d entries" % (n_groups * self.D))
        self.C = int(C)
        self.per_channel = out_mode == OUT_PER_CHANNEL
        self.n_groups = int(n_groups)
        self.ard = bool(ard)

    def ref(self):
        return ctypes.byref(self.desc)

    def workspace(self, N, M, device):
        nbytes = lib().cgp_workspace_bytes(self.ref(), int(N), int(M))
        if nbytes == 0:
            return None, 0
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return ws, nbytes
