"""ctypes binding of ``libfloppity_b200.so`` (the C ABI declared in ``include/floppity_b200.h``).

The library is built in-tree by ``csrc/build.sh`` (nvcc, sm_100a only).  There is NO fallback: if the
shared object is missing, or a kernel is requested without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfloppity_b200.so")
BUILD_SCRIPT = os.path.join(_HERE, "csrc", "build.sh")

FP_E = {-1: "FP_E_BADARG", -2: "FP_E_TOOLARGE", -3: "FP_E_NOTREADY"}

# fp_param_kind
P_CTX_W, P_CTX_B, P_W0X, P_W1, P_B1, P_W2, P_B2, P_WF, P_BF, P_LU_LOWER, P_LU_UPPER, P_LU_DIAG, P_LU_BIAS = range(13)
# fp_gemm_flags
G_RELU_A, G_RELU_B, G_BIAS, G_ADD_CTX, G_GLU_RES, G_RELUMASK, G_ACCUM, G_ATOMIC, G_TF32X3 = 1, 2, 4, 8, 16, 32, 64, 128, 256
G_COLSUM_A, G_DROP_OUT = 4096, 8192


class NsfConfig(Structure):
    """Mirror of ``fp_nsf_config``."""
    _fields_ = [("D", c_int32), ("C", c_int32), ("T", c_int32), ("K", c_int32), ("H", c_int32), ("Bk", c_int32),
                ("tail_bound", c_float), ("min_bin_width", c_float), ("min_bin_height", c_float),
                ("min_derivative", c_float), ("lu_eps", c_float), ("dropout_p", c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}

    @classmethod
    def from_dict(cls, d):
        return cls(**d)


class GemmArgs(Structure):
    """Mirror of ``fp_gemm_args``."""
    _fields_ = [("A", c_void_p), ("lda", c_int64), ("a_kmajor", c_int32),
                ("B", c_void_p), ("ldb", c_int64), ("b_kmajor", c_int32),
                ("C", c_void_p), ("ldc", c_int64),
                ("M", c_int64), ("N", c_int32), ("K", c_int64),
                ("flags", c_int32),
                ("bias", c_void_p),
                ("G", c_void_p), ("ldg", c_int64), ("ctx_div", c_int32),
                ("Rsd", c_void_p), ("ldr", c_int64),
                ("U", c_void_p), ("ldu", c_int64),
                ("Msk", c_void_p), ("ldm", c_int64),
                ("split_k", c_int32), ("mask_scale", c_float), ("drop_p", c_float), ("drop_seed", c_uint64),
                ("a_scale", c_void_p), ("status", c_void_p),
                ("pre_hi16", c_void_p), ("pre_lo16", c_void_p), ("pre_ld", c_int64)]


_CFG = POINTER(NsfConfig)
_P = c_void_p

# name -> (restype, argtypes); every symbol include/floppity_b200.h declares
SIGNATURES = {
    "fp_nsf_param_count": (c_int64, [_CFG]),
    "fp_nsf_grad_count": (c_int64, [_CFG]),
    "fp_nsf_param_offset": (c_int64, [_CFG, c_int32, c_int32, c_int32]),
    "fp_nsf_param_size": (c_int64, [_CFG, c_int32, c_int32]),
    "fp_nsf_dtr": (c_int32, [_CFG, c_int32]),
    "fp_nsf_prepared_count": (c_int64, [_CFG]),
    "fp_nsf_workspace_count": (c_int64, [_CFG, c_int64, c_int64, c_int32]),
    "fp_rqs_forward": (c_int32, [_CFG, c_int32, _P, _P, _P, _P, _P, c_int64, _P, _P]),
    "fp_rqs_inverse": (c_int32, [_CFG, c_int32, _P, _P, _P, _P, _P, c_int64, _P, _P]),
    "fp_rqs_backward": (c_int32, [_CFG, c_int32, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "fp_gemm": (c_int32, [POINTER(GemmArgs), _P]),
    "fp_gemm_tc": (c_int32, [POINTER(GemmArgs), _P, _P, _P]),
    "fp_split_tf32": (c_int32, [_P, _P, _P, c_int64, _P]),
    "fp_gemm_tc16": (c_int32, [POINTER(GemmArgs), _P, _P, _P]),
    "fp_split_f16": (c_int32, [_P, _P, _P, c_int64, _P]),
    "fp_gemm_tc16_tn": (c_int32, [POINTER(GemmArgs), _P]),
    "fp_gemm_tc_tn": (c_int32, [POINTER(GemmArgs), _P]),
    "fp_tc_set_trace": (c_int32, [_P]),
    "fp_dropout_mask": (c_int32, [_P, c_int64, c_int32, c_float, c_uint64, _P]),
    "fp_dropout_layer_seed": (c_uint64, [c_uint64, c_int32, c_int32]),
    "fp_set_mode": (c_int32, [c_int32, c_int32, c_int32, c_int32]),
    "fp_set_precision": (c_int32, [c_int32]),
    "fp_set_gemm_format": (c_int32, [c_int32]),
    "fp_nsf_prepare": (c_int32, [_CFG, _P, _P, _P]),
    "fp_nsf_logprob_forward": (c_int32, [_CFG, _P, _P, _P, _P, c_int64, c_int64, _P, _P, c_float, _P, _P, c_int32,
                                         c_uint64, _P, _P]),
    "fp_nsf_logprob_backward": (c_int32, [_CFG, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, c_uint64, _P, _P]),
    "fp_nsf_finish_grads": (c_int32, [_CFG, _P, _P, _P, _P]),
    "fp_nsf_sample": (c_int32, [_CFG, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, _P]),
    "fp_atom_choices": (c_int32, [_P, c_int64, c_int32, c_uint64, _P]),
    "fp_atom_choices_step": (c_int32, [_P, c_int64, c_int32, c_uint64, _P, _P]),
    "fp_atom_gather": (c_int32, [_P, _P, _P, c_int64, c_int32, c_int32, _P]),
    "fp_atomic_loss": (c_int32, [_P, _P, _P, _P, _P, c_int32, _P, _P, c_int64, c_int32, c_int32, _P, _P]),
    "fp_grad_sumsq": (c_int32, [_P, c_int64, _P, _P]),
    "fp_adam_step": (c_int32, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, _P, c_float, c_float, _P, _P]),
    "fp_gather_rows": (c_int32, [_P, _P, _P, c_int64, c_int32, _P]),
    "fp_standardize": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, _P]),
    "fp_sum": (c_int32, [_P, c_int64, c_float, _P, _P]),
    "fp_within_box": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, _P]),
    "fp_compact_inside": (c_int32, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int64, _P]),
    "fp_preprocess_rows": (c_int32, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, c_int64, _P]),
    "fp_gauss_loglik": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, _P]),
    "fp_relu": (c_int32, [_P, _P, c_int64, _P]),
    "fp_relu_bwd": (c_int32, [_P, _P, _P, c_int64, _P]),
    "fp_conv1d_relu_pool_fwd": (c_int32, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P]),
    "fp_conv1d_relu_pool_bwd": (c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32,
                                          c_int32, _P]),
    "fp_nsf_ctx_backward": (c_int32, [_CFG, _P, c_int64, c_int64, _P, _P, _P]),
    "fp_launch_count": (c_int64, []),
    "fp_prof_enable": (c_int32, [c_int32]),
    "fp_prof_ntags": (c_int32, []),
    "fp_prof_tag_name": (c_char_p, [c_int32]),
    "fp_prof_collect": (c_int32, [_P, _P, _P, _P]),
    "fp_tc_set_debug": (c_int32, [c_int32]),
    "fp_cf_set_trace": (c_int32, [_P]),
    "fp_version": (c_char_p, []),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into ``libfloppity_b200.so`` (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["bash", BUILD_SCRIPT], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libfloppity_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    global _lib
    _lib = None
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the shared object (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or floppity_deprecated_b200/csrc/build.sh). This package has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int, what: str) -> None:
    if code == 0:
        return
    if code < 0:
        raise RuntimeError(f"{what}: {FP_E.get(code, code)}")
    raise RuntimeError(f"{what}: CUDA error {code} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no device'})")


def ptr(t) -> int:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("floppity_deprecated_b200 kernels need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise RuntimeError("tensor must be contiguous")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("floppity_deprecated_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
