"""ctypes binding of the C ABI in include/mimrl_b200.h.

The shared library is the product: there is no CPU or PyTorch fallback.  If it
has not been built, importing this module raises; if a tensor is not a CUDA
tensor, the wrappers raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmimrl_b200.so")

BOUND_IDS = {"dv": 0, "mine": 1, "tuba": 2, "nwj": 3, "infonce": 4, "js_fgan": 5, "js": 6, "smile": 7,
             "interpolate": 8}
STAT_CLAMP, STAT_SOFTPLUS, STAT_MAXONLY = 1, 2, 4
WEIGHT_EXP, WEIGHT_SIGMOID, WEIGHT_INTERP = 0, 1, 2
IMPL_AUTO, IMPL_FFMA, IMPL_TCGEN05 = 0, 1, 2

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). mimrl_b200 has no fallback path.")

lib = ctypes.CDLL(LIB_PATH)

_P = c_void_p
_SIGS = {
    "mimrl_version": (c_int, []),
    "mimrl_last_error": (c_char_p, []),
    "mimrl_launch_count": (c_uint64, []),
    "mimrl_sep_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mimrl_sep_selected_impl": (c_int, [c_int, c_int, c_int, c_int]),
    "mimrl_sep_row_stats": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "mimrl_sep_fused_forward": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "mimrl_sep_online_forward": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "mimrl_sep_interp_stats": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "mimrl_sep_weighted_sum": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_int, _P,
                                       _P, c_size_t, _P]),
    "mimrl_bound_finalize": (c_int, [c_int, _P, _P, _P, _P, _P, c_int, _P, _P]),
    "mimrl_bound_backward_coef": (c_int, [c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P]),
    "mimrl_bound_weight_family": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "mimrl_scores_row_stats": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "mimrl_scores_grad": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "mimrl_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mimrl_gemm_f32x3": (c_int, [c_int, _P, _P, _P, c_int, c_int, c_int, _P, c_int, _P, _P, c_size_t, _P]),
    "mimrl_split_bytes": (c_size_t, [c_int, c_int]),
    "mimrl_split_f32": (c_int, [_P, _P, c_int, c_int, _P, _P, _P]),
    "mimrl_gemm_split_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mimrl_gemm_split": (c_int, [c_int, _P, _P, c_int, c_int, c_int, _P, c_int, _P, _P, c_size_t, _P]),
    "mimrl_split_f32_hmask": (c_int, [_P, _P, c_int, c_int, _P, _P, _P]),
    "mimrl_mlp4_supported": (c_int, [c_int, c_int, c_int]),
    "mimrl_mlp4_fwd": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               _P, _P]),
    "mimrl_mlp4_bwd": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               _P, _P, _P, _P]),
    "mimrl_gemm_split_blocked": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "mimrl_gemm_split_blocked_acc": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    "mimrl_knn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mimrl_knn_search": (c_int, [_P, c_int, c_int, _P, c_int, c_int, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "mimrl_legacy_permutation_head": (c_int, [_P, _P, c_int64, c_int64, _P]),
    "mimrl_knn_fit_bytes": (c_size_t, [c_int, c_int]),
    "mimrl_knn_fit": (c_int, [_P, c_int, c_int, _P, c_size_t, _P]),
    "mimrl_knn_search_fitted": (c_int, [_P, c_int, c_int, _P, c_size_t, _P, c_int, c_int, c_float, c_int, _P, _P, _P, _P,
                                        c_size_t, _P]),
    "mimrl_knn_search_rows": (c_int, [_P, c_int, c_int, c_int64, _P, c_int, _P, c_int, c_int, c_int, _P, _P, _P,
                                      c_size_t, _P]),
    "mimrl_gather_rows": (c_int, [_P, c_int, c_int, _P, c_int, c_int, c_int, _P, _P]),
    "mimrl_vcmi_head_fwd": (c_int, [_P, c_int, c_int, _P, _P]),
    "mimrl_vcmi_head_bwd": (c_int, [_P, c_int, c_int, _P, _P, _P]),
    "mimrl_cubemlp_saved_floats": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mimrl_cubemlp_mix_fwd": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, _P, c_int, c_int,
                                      _P, _P, _P]),
    "mimrl_cubemlp_tc_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "mimrl_cubemlp_tc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mimrl_cubemlp_mix_fwd_tc": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, _P, c_int, _P, _P,
                                         _P, c_size_t, _P, _P, c_int, c_int, _P]),
    "mimrl_cubemlp_prep_many": (c_int, [c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mimrl_cubemlp_mix_bwd": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, _P, c_int,
                                      c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mimrl_cubemlp_small_supported": (c_int, [c_int, c_int, c_int]),
    "mimrl_cubemlp_small_bwd": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, _P, c_int, c_int,
                                        _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mimrl_cubemlp_tc_fibre_rows": (c_int64, [c_int, c_int]),
    "mimrl_cubemlp_mix_bwd_tc": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, _P, c_int, _P,
                                         _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, c_int, _P]),
    "mimrl_cubemlp_tc_op_bytes": (c_size_t, [c_int, c_int64]),
    "mimrl_concat_tc_supported": (c_int, [c_int, c_int]),
    "mimrl_concat_workspace_bytes": (c_size_t, [c_int]),
    "mimrl_concat_scores": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "mimrl_concat_pair_rows": (c_int64, [c_int, c_int]),
    "mimrl_concat_grad": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _P, _P, c_size_t, _P]),
    "mimrl_concat_stats_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mimrl_concat_row_stats": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                       _P, c_size_t, _P]),
    "mimrl_concat_grad_fused": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _P,
                                        _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "mimrl_linear_small_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mimrl_linear_small": (c_int, [c_int, _P, _P, _P, c_int, c_int, c_int, _P, c_int, _P, _P, _P, c_size_t, _P]),
    "mimrl_mlp4_small_supported": (c_int, [c_int, c_int, c_int]),
    "mimrl_mlp4_small_fwd": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P]),
    "mimrl_mlp4_small_bwd": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                     _P, _P, _P, _P, _P, _P]),
    "mimrl_feature_stack_fwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "mimrl_feature_stack_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
    "mimrl_feature_reduce_fwd": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P]),
    "mimrl_feature_reduce_bwd": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P]),
}
EXPORTS = tuple(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


class MimrlError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise MimrlError(f"mimrl_b200 error {rc}: {lib.mimrl_last_error().decode()}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or NULL for None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise MimrlError("mimrl_b200 kernels need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise MimrlError("mimrl_b200 kernels need contiguous tensors")
    return t.data_ptr()


def f32(t):
    """Borrow ``t`` as a contiguous fp32 CUDA tensor (copy only if needed)."""
    if not t.is_cuda:
        raise MimrlError("mimrl_b200 kernels need CUDA tensors; there is no CPU fallback")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib.mimrl_launch_count())
