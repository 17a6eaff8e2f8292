"""ctypes binding of the C ABI declared in include/hlvae_b200.h.

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a) into
`hl-vae_b200/lib/libhlvae_b200.so`.  There is no fallback: if the library is missing, or a
tensor is not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# HLVAE_B200_LIB: load another build of the same sources (A/B runs of kernel variants); default = the in-tree build
LIB_PATH = os.environ.get("HLVAE_B200_LIB") or os.path.join(_HERE, "lib", "libhlvae_b200.so")

MAX_COMPS, MAX_DISC, MAX_Q, TMAX, MAX_CLASS, MAX_Y = 8, 3, 8, 64, 16, 16
HEAD_AFFINE, HEAD_SIGMOID, HEAD_ZERO, HEAD_BIAS = 0, 1, 2, 3
F32, F64, U8 = 0, 1, 2
KIND_CAT, KIND_BIN = 1, 2
AUX_REAL, AUX_POS, AUX_BETA = 0, 1, 2
VAR_KINDS = {"real": 0, "pos": 1, "count": 2, "cat": 3, "ordinal": 4}
ACC_NAMES = ("S", "p", "gw", "scal", "gZ", "gos0", "gls0", "gos1", "gls1", "total")
NSCAL = 4
STATUS_NOT_PD, STATUS_T_TOO_LARGE = 1, 2


class Comp(C.Structure):
    _fields_ = [("se_col", C.c_int32), ("ndisc", C.c_int32),
                ("disc_kind", C.c_int32 * MAX_DISC), ("disc_col", C.c_int32 * MAX_DISC)]


class KSpec(C.Structure):
    _fields_ = [("ncomp", C.c_int32), ("reserved", C.c_int32), ("comp", Comp * MAX_COMPS)]


_lib = None

_P, _I, _L, _D = C.c_void_p, C.c_int, C.c_int64, C.c_double
_SIGS = {
    "hlvae_version": ([], _I),
    "hlvae_sizeof_kspec": ([], _I),
    "hlvae_kernel_eval_fwd": ([C.POINTER(KSpec), _P, _P, _I, _I, _P, _I, _L, _L, _P, _I, _L, _L, _P, _P], _I),
    "hlvae_kernel_eval_bwd": ([C.POINTER(KSpec), _P, _P, _I, _I, _P, _I, _L, _L, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P], _I),
    "hlvae_hyper_constrain": ([_I, _I, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_double), _P, _P, _P, _P], _I),
    "hlvae_subject_matvec": ([C.POINTER(KSpec), _P, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P], _I),
    "hlvae_kl_acc_layout": ([_I, _I, _I, C.POINTER(_L)], _I),
    "hlvae_kl_subject": ([C.POINTER(KSpec), _P, _P, C.POINTER(KSpec), _P, _P, _P, _I, _I, _P, _L, _P, _P, _P, _I, _I,
                          _P, _L, _I, _P, _L, _P, _I, _P, _D, _P, _P], _I),
    "hlvae_kl_panel": ([C.POINTER(KSpec), _P, _P, C.POINTER(KSpec), _P, _P, _I, _I, _I, _P, _L, _P, _P, _P, _P, _I, _I,
                        _P, _L, _I, _P, _P, _P, _L, _P, _P, _P, _D, _P, _I, _P], _I),
    "hlvae_mxm_workspace_doubles": ([_I, _I], _L),
    "hlvae_mxm_pre": ([C.POINTER(KSpec), _P, _P, _I, _I, _I, _P, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "hlvae_mxm_post": ([_I, _I, _D, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "hlvae_natgrad_update": ([_I, _I, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "hlvae_mxm_aux": ([C.POINTER(KSpec), _P, _P, _I, _I, _I, _P, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "hlvae_kernel_matvec": ([C.POINTER(KSpec), _P, _P, _I, _I, _P, _I, _P, _I, _P, _P, _P], _I),
    "hlvae_loglik_fwd": ([_L, _I, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "hlvae_loglik_bwd": ([_L, _I, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P], _I),
    "hlvae_statistics": ([_L, _I, _L, _P, _P, _P, _P, _P, _I, _P, _P, _P], _I),
    "hlvae_discrete_transform": ([_L, _I, _L, _P, _P, _P, _P, _I, _P, _P], _I),
    "hlvae_loglik_aux_fwd": ([_I, _L, _I, _P, _L, _I, _P, _L, _I, _P, _L, _L, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P], _I),
    "hlvae_loglik_aux_bwd": ([_I, _L, _I, _P, _L, _I, _P, _L, _I, _P, _L, _L, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "hlvae_contraction_probe": ([_I, _I, _I, _L, _I, _P, _P, _D, _D, _I, _P, _P, _P], _I),
    "hlvae_batch_norm_stats": ([_L, _I, _L, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P, _P], _I),
    "hlvae_batch_norm_apply": ([_L, _I, _L, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P], _I),
    "hlvae_theta_fwd": ([_L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _I, _P, _L, _P], _I),
    "hlvae_theta_bwd": ([_L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _I, _P, _I, _P, _L, _P, _P, _P, _P], _I),
}
EXPORTED = tuple(_SIGS)


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"hlvae_b200: CUDA library not built ({LIB_PATH} missing). Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. There is no CPU fallback.")
        h = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(h, name)
            fn.argtypes = args
            fn.restype = res
        if h.hlvae_sizeof_kspec() != C.sizeof(KSpec):
            raise RuntimeError("hlvae_b200: hlvae_kspec_t layout mismatch between binding and library")
        _lib = h
    return _lib


# Optional per-call device timing: when PROFILE is a list, every C-ABI call is bracketed by CUDA
# events on the launching stream (no sync); bench.py reads them after the timed region.
PROFILE = None
LAUNCHES = 0


def workspace(L, M, device):
    n = int(lib().hlvae_mxm_workspace_doubles(L, M))
    return torch.empty(n, dtype=torch.float64, device=device) if n else None


class _DevPtr(C.c_void_p):
    """Device pointer that remembers which GPU it belongs to (checked / used by `call`)."""
    dev = None


class _CurrentStream:
    """Placeholder returned by stream_ptr(): resolved by `call` to torch's current stream of the device the
    call's tensors live on."""


CURRENT_STREAM = _CurrentStream()


def call(name, *args):
    """Invoke C-ABI entry point `name` (each one enqueues exactly one kernel) and check its result.

    The kernel is launched on the device of the tensors passed through `ptr()` (they must share one device), on
    torch's current stream OF THAT DEVICE, with that device made current for the duration of the call: the C
    entry points never call cudaSetDevice themselves."""
    global LAUNCHES
    fn = getattr(lib(), name)
    dev = None
    for a in args:
        if isinstance(a, _DevPtr):
            if dev is None:
                dev = a.dev
            elif a.dev != dev:
                raise RuntimeError(f"hlvae_b200: {name}: tensors live on different CUDA devices ({dev} and {a.dev})")
    cur = torch.cuda.current_device()
    if dev is None:
        dev = cur
    if any(a is CURRENT_STREAM for a in args):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        args = tuple(st if a is CURRENT_STREAM else a for a in args)
    LAUNCHES += 1
    if dev != cur:
        with torch.cuda.device(dev):
            return _invoke(fn, name, args)
    return _invoke(fn, name, args)


def _invoke(fn, name, args):
    if PROFILE is None:
        return check(fn(*args), name)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    PROFILE.append((name, e0, e1))
    return check(rc, name)


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"hlvae_b200: {what} failed with code {rc}" +
                           (" (argument error)" if rc == -1 else " (unsupported size)" if rc == -2 else " (CUDA error)"))


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL), tagged with the tensor's device index."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("hlvae_b200: tensors must live on a CUDA device (no CPU fallback)")
    p = _DevPtr(t.data_ptr())
    p.dev = t.device.index
    return p


def stream_ptr():
    """The launching stream: torch's current stream of the call's device (resolved inside `call`)."""
    return CURRENT_STREAM


def dtype_code(t):
    if t.dtype == torch.float64:
        return F64
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"hlvae_b200: unsupported storage dtype {t.dtype}")


def acc_layout(L, M, Q):
    off = (_L * 10)()
    check(lib().hlvae_kl_acc_layout(L, M, Q, off), "hlvae_kl_acc_layout")
    return {n: int(off[i]) for i, n in enumerate(ACC_NAMES)}
