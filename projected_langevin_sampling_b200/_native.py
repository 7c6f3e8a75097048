"""ctypes binding of libpls_b200.so (C ABI in include/pls_b200.h).

There is no CPU or PyTorch fallback: if the library is missing, or no sm_100 device is present when a compute call
is made, this module raises.  PyTorch is used only for device memory (caller-owned tensors) and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpls_b200.so")

# enums of include/pls_b200.h
KERNEL_RBF, KERNEL_LINEAR = 0, 1
COST_GAUSSIAN, COST_BERNOULLI, COST_POISSON, COST_MULTIMODAL, COST_STUDENT_T = range(5)
LINK_IDENTITY, LINK_SIGMOID, LINK_PROBIT, LINK_SQUARE = range(4)
EPI_PREDICTION, EPI_COST_DERIVATIVE, EPI_COST, EPI_COST_DERIVATIVE_AND_COST = range(4)
NOISE_NONE, NOISE_GIVEN, NOISE_PHILOX = range(3)
CV_TIES_HIGHEST_INDEX, CV_TIES_HOST = 0, 1
CV_HEADER_DOUBLES = 16
ABI_VERSION = 2
COST_VALUE_TILE_ROWS = 128  # rows per partial sum of pls_cost_value_f64 (dense F)


class PlsCost(C.Structure):
    """struct pls_cost"""

    _fields_ = [
        ("cost_id", C.c_int32),
        ("link_id", C.c_int32),
        ("closed_form", C.c_int32),
        ("reserved", C.c_int32),
        ("observation_noise", C.c_double),
        ("shift", C.c_double),
        ("bernoulli_noise", C.c_double),
        ("degrees_of_freedom", C.c_double),
        ("scale", C.c_double),
        ("link_jitter", C.c_double),
        ("probit_divisor", C.c_double),
        ("log_weight_1", C.c_double),
        ("log_weight_2", C.c_double),
        ("log_normaliser", C.c_double),
    ]


class StepPlan(C.Structure):
    """struct pls_step_plan"""

    _fields_ = [(k, C.c_int64) for k in ("n", "m", "m_k", "j", "ldj", "chunk_rows", "cost_tiles")] + \
               [(k, C.c_int32) for k in ("n_chunks", "splits", "tile_rows", "gram_mode", "with_cost", "reserved")] + \
               [(k, C.c_int64) for k in ("off_w", "off_gm", "off_dc", "off_gp", "off_cost_partial", "off_kstage", "workspace_bytes")]


GRAM_GENERATED, GRAM_STAGED, GRAM_CACHED = range(3)


class NativeLibraryError(RuntimeError):
    pass


_i64, _int, _dbl, _vp, _u64 = C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_uint64
_costp = C.POINTER(PlsCost)
# int64_t (*pls_cv_tie_fn)(void* user, const double* d_host, int64_t n, const int64_t* chosen, int n_chosen)
CV_TIE_FN = C.CFUNCTYPE(C.c_int64, C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int64), C.c_int)

# name -> (restype, argtypes); kept in the order of include/pls_b200.h
SIGNATURES = {
    "pls_abi_version": (_int, []),
    "pls_ctx_create": (_int, [_int, C.POINTER(_vp)]),
    "pls_ctx_destroy": (None, [_vp]),
    "pls_last_error": (C.c_char_p, [_vp]),
    "pls_sm_count": (_int, [_vp]),
    "pls_point_stride": (_int, [_int]),
    "pls_forward_tile_rows": (_int, [_vp, _i64]),
    "pls_set_tile_shape": (None, [_vp, _int]),
    "pls_set_tile_sets": (None, [_vp, _int]),
    "pls_set_tile_cluster": (None, [_vp, _int]),
    "pls_backward_splits": (_int, [_vp, _i64, _i64, _i64]),
    "pls_prepare_points_f64": (_int, [_vp, _int, _vp, _i64, _int, _i64, C.POINTER(_dbl), C.POINTER(_dbl), _dbl, _vp, _vp]),
    "pls_gram_f64": (_int, [_vp, _int, _vp, _i64, _vp, _i64, _int, _vp, _i64, _vp]),
    "pls_gemm_f64": (_int, [_vp, _int, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp]),
    "pls_forward_f64": (_int, [_vp, _int, _vp, _i64, _vp, _i64, _int, _vp, _i64, _i64, _int, _costp, _vp, _vp, _i64, _vp]),
    "pls_forward_step_f64": (_int, [_vp, _int, _vp, _i64, _vp, _i64, _int, _vp, _i64, _i64, _costp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "pls_backward_f64": (_int, [_vp, _int, _vp, _i64, _vp, _i64, _int, _vp, _i64, _i64, _vp, _i64, _int, _int, _vp]),
    "pls_gram_fill_f64": (_int, [_vp, _int, _vp, _i64, _vp, _i64, _int, _vp, _i64, _vp]),
    "pls_gram_cache_ld": (_i64, [_i64]),
    "pls_gram_cache_rows": (_i64, [_i64]),
    "pls_forward_cached_f64": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _int, _costp, _vp, _vp, _i64, _vp]),
    "pls_forward_step_cached_f64": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _costp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "pls_backward_cached_f64": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _int, _int, _vp]),
    "pls_reduce_splits_f64": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _vp, _i64, _vp]),
    "pls_project_update_f64": (
        _int,
        [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _i64, _vp, _dbl, _int, _vp, _i64, _u64, _u64, _i64, _int, _vp, _i64, _vp],
    ),
    "pls_set_step_counter": (None, [_vp, _vp]),
    "pls_advance_step_counter": (_int, [_vp, _vp, _u64, _vp]),
    "pls_cost_derivative_f64": (_int, [_vp, _costp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "pls_cost_value_f64": (_int, [_vp, _costp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "pls_energy_terms_f64": (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "pls_philox_normal_f64": (_int, [_vp, _u64, _u64, _i64, _i64, _i64, _vp, _i64, _vp]),
    "pls_gram_exp_f64": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "pls_lincomb3_f64": (_int, [_vp, _i64, _i64, _dbl, _vp, _i64, _dbl, _vp, _i64, _dbl, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    "pls_flat_math_f64": (_int, [_vp, _int, _vp, _vp, _i64, _vp, _vp]),
    "pls_step_plan_f64": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _int, _int, C.POINTER(StepPlan)]),
    "pls_grad_f64": (_int, [_vp, C.POINTER(StepPlan), _int, _int, _vp, _vp, _vp, _i64, _vp, _i64, _costp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "pls_step_f64": (_int, [_vp, C.POINTER(StepPlan), _int, _int, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _costp, _vp, _vp, _i64, _dbl, _int,
                            _vp, _i64, _u64, _u64, _i64, _int, _vp, _i64, _vp, _vp, _vp]),
    "pls_profile_begin": (_int, [_vp]),
    "pls_profile_end": (_int, [_vp, C.POINTER(_dbl)]),
    "pls_cv_scratch_doubles": (_i64, [_i64, _int, _int]),
    "pls_cv_shard_scratch_doubles": (_i64, [_i64, _int, _int]),
    "pls_cv_candidate_doubles": (_i64, [_int, _int]),
    "pls_cv_shard_begin_f64": (_int, [_vp, _int, _vp, _i64, _i64, _int, _dbl, _int, _dbl, _vp, _vp, _vp, _vp]),
    "pls_cv_shard_pick_f64": (_int, [_vp, _vp, _int, _int, _int, _int, _dbl, _int, _int, _int, _i64, _i64, _vp, _vp, _vp]),
    "pls_cv_shard_update_f64": (_int, [_vp, _int, _vp, _i64, _i64, _int, _int, _int, _dbl, _vp, _vp, _vp, _vp, _vp]),
    "pls_cv_shard_force_f64": (_int, [_vp, _vp, _i64, _i64, _int, _int, _int, _i64, _vp, _vp, _vp, _vp, _vp]),
    "pls_cv_shard_status": (_int, [_vp, _vp, C.POINTER(_i64), _vp]),
    "pls_cv_shard_finish": (_int, [_vp, _vp, C.POINTER(_int), _vp]),
    "pls_cv_select_f64": (_int, [_vp, _int, _vp, _i64, _int, _dbl, _int, _dbl, _dbl, _int, _int, CV_TIE_FN, _vp, _vp, _vp, _vp, _vp,
                                 C.POINTER(_int), _vp]),
}

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()
_ctxs: Dict[int, "Context"] = {}


def load_library() -> C.CDLL:
    """dlopen libpls_b200.so and bind every symbol of the header.  Raises if the library was not built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  projected_langevin_sampling_b200 has no CPU / PyTorch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.pls_abi_version() != ABI_VERSION:
            raise NativeLibraryError(f"ABI mismatch: library {lib.pls_abi_version()} != binding {ABI_VERSION}")
        _lib = lib
        return lib


def point_stride(d: int) -> int:
    sp = load_library().pls_point_stride(int(d))
    if sp < 0:
        raise ValueError(f"input dimension D={d} is not supported by the CUDA path (1 <= D <= 26)")
    return sp


def require_cuda_tensor(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float64:
        raise TypeError(f"{name} must be a float64 CUDA tensor")
    return t


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Context:
    """One pls_ctx per CUDA device.  All calls are enqueued on torch's current stream of that device."""

    def __init__(self, device_index: int):
        self.lib = load_library()
        self.device_index = device_index
        handle = _vp()
        rc = self.lib.pls_ctx_create(device_index, C.byref(handle))
        if rc != 0:
            raise NativeLibraryError(self.lib.pls_last_error(None).decode())
        self.handle = handle
        self.sm_count = self.lib.pls_sm_count(handle)
        self.launches = 0  # kernels launched through this context (bench.py's gpu_launches)

    def stream(self) -> int:
        return torch.cuda.current_stream(self.device_index).cuda_stream

    def check(self, rc: int):
        if rc != 0:
            raise NativeLibraryError(self.lib.pls_last_error(self.handle).decode())

    def __del__(self):  # pragma: no cover
        try:
            self.lib.pls_ctx_destroy(self.handle)
        except Exception:
            pass


def context(device: Optional[torch.device] = None) -> Context:
    """The Context of `device` (default: the current CUDA device).  Fails loudly without a GPU."""
    if not torch.cuda.is_available():
        raise NativeLibraryError(
            "no CUDA device: projected_langevin_sampling_b200 runs its hot path only on a B200 (sm_100a); "
            "there is no CPU fallback"
        )
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    with _lock:
        ctx = _ctxs.get(idx)
    if ctx is None:
        ctx = Context(idx)
        with _lock:
            _ctxs[idx] = ctx
    return ctx
