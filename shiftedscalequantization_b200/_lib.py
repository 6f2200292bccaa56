"""ctypes binding of libssq_b200.so (declared in include/ssq_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails the caller gets an
exception, never a silently slower or CPU result.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libssq_b200.so"

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int
_f = C.c_float
_d = C.c_double
_sz = C.c_size_t

MT_MAX = 16
MT_TILE = 4096
MAX_SHIFTS = 4
N_CANDIDATES = 80
SHIFT_DEQUANT = 0
SHIFT_ADASHIFT = 1


class AdaRoundDesc(C.Structure):
    """mirror of ssq_adaround_desc"""
    _fields_ = [
        ("w", _p), ("alpha", _p), ("delta", _p), ("zero_point", _p),
        ("wq", _p), ("gwq", _p), ("galpha", _p),
        ("n", _i64), ("inner", _i64), ("nchan", _i64), ("tile_begin", _i64),
        ("qmin", _f), ("qmax", _f),
    ]


class AffineDesc(C.Structure):
    """mirror of ssq_affine_desc"""
    _fields_ = [
        ("gamma", _p), ("phi", _p), ("bias", _p), ("beff", _p), ("gbeff", _p), ("ggamma", _p), ("gphi", _p),
        ("row_begin", _i64),
    ]


class IterState(C.Structure):
    """mirror of ssq_iter_state"""
    _fields_ = [
        ("step", _p), ("idx_table", _p), ("idx_live", _p), ("b_table", _p), ("b_live", _p),
        ("lr_table", _p), ("lr_live", _p), ("n_steps", _i64), ("batch", _i32),
    ]


# name -> (restype, argtypes); must list every function of include/ssq_b200.h
PROTOTYPES = {
    "ssq_abi_version": (_i32, []),
    "ssq_status_string": (C.c_char_p, [_i32]),
    "ssq_ws_bytes": (_sz, [_i64]),
    "ssq_fq_affine_fwd": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _f, _p]),
    "ssq_fq_affine_bwd": (_i32, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _f, _p, _sz, _p]),
    "ssq_fq_adaround_fwd": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _f, _i32, _p, _f, _p, _p, _sz, _p]),
    "ssq_fq_adaround_bwd": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _f, _p, _f, _p, _i32, _p]),
    "ssq_adaround_init_alpha": (_i32, [_p, _p, _p, _i64, _i64, _i64, _p]),
    "ssq_round_reg_fwd": (_i32, [_p, _i64, _p, _f, _p, _p, _sz, _p]),
    "ssq_round_reg_bwd": (_i32, [_p, _i64, _p, _f, _p, _p, _i32, _p]),
    "ssq_fq_adaround_fwd_mt": (_i32, [C.POINTER(AdaRoundDesc), _i32, _i64, _i32, _p, _f, _p, _p, _sz, _p]),
    "ssq_fq_adaround_bwd_mt": (_i32, [C.POINTER(AdaRoundDesc), _i32, _i64, _p, _f, _p]),
    "ssq_shift_probs_fwd": (_i32, [_p, _p, _i64, _i32, _i32, _p, _f, _p, _p, _sz, _p]),
    "ssq_shift_probs_bwd": (_i32, [_p, _p, _p, _i64, _i32, _i32, _p, _f, _p, _p]),
    "ssq_fq_shift_fwd": (_i32, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _f, _f, _p]),
    "ssq_fq_shift_bwd": (_i32, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _f, _f, _p, _sz, _p]),
    "ssq_shift_bwd_ws_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32]),
    "ssq_mse_scale_search": (_i32, [_p, _i64, _i64, _i32, _i32, _f, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ssq_mse_scale_search_ws_bytes": (_sz, [_i64, _i64]),
    "ssq_row_minmax": (_i32, [_p, _i64, _i64, _p, _p, _p, _sz, _p]),
    "ssq_inp_scale_search": (_i32, [_p, _p, _p, _p, _i32, _f, _f, _f, _p, _i64, _i64, _p, _sz, _p]),
    "ssq_inp_scale_search_ex": (_i32, [_p, _p, _p, _p, _i32, _f, _f, _f, _p, _i64, _i64, _i32, _p, _sz, _p]),
    "ssq_inp_scale_search_ws_bytes": (_sz, [_i64]),
    "ssq_inp_scale_search_ws_bytes2": (_sz, [_i64, _i64]),
    "ssq_recon_loss": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i64, _d, _i32, _f, _p, _p, _sz, _p]),
    "ssq_recon_loss_bwd": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i64, _d, _i32, _f, _p, _sz, _p]),
    "ssq_chan_affine_fwd": (_i32, [_p, _p, _p, _p, _i64, _i64, _i64, _p]),
    "ssq_chan_affine_bwd": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p, _sz, _p]),
    "ssq_adam_step": (_i32, [_p, _p, _p, _p, _i64, _p, _d, _d, _d, _p, _p]),
    "ssq_adam_step_end_iteration": (_i32, [_p, _p, _p, _p, _i64, _p, _d, _d, _d, _p, _p, _sz, _p]),
    "ssq_adam_step_pending": (_i32, [_p, _p, _p, _p, _i64, _p, _d, _d, _d, _p, _p]),
    "ssq_iter_prologue": (_i32, [C.POINTER(IterState), _p, _p, _i64, C.POINTER(AdaRoundDesc), C.POINTER(AffineDesc), _i32, _i64, _f, _p, _p, _sz, _p]),
    "ssq_fq_adaround_bwd_adam_mt": (_i32, [C.POINTER(AdaRoundDesc), C.POINTER(AffineDesc), _i32, _i64, _p, _f, _p, _p, _p, _p, _p, _d, _d, _d, _i32, _i32, _p, _sz, _p]),
    "ssq_affine_grad_mt": (_i32, [C.POINTER(AdaRoundDesc), C.POINTER(AffineDesc), _i32, _p]),
    "ssq_exchange_pad_bytes": (_sz, []),
    "ssq_exchange_shard_elems": (_i64, [_i64, _i32]),
    "ssq_grad_exchange_adam": (_i32, [_p, _p, _p, _i32, _i32, _i64, _p, _p, _p, _p, _d, _d, _d, _p, _p, _p, _p, _sz, _p]),
    "ssq_gather_rows": (_i32, [_p, _p, _p, _i64, _i64, _p]),
    "ssq_packed_row_bytes": (_i64, [_i64, _i32]),
    "ssq_export_codes": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _f, _f, _i32, _p]),
    "ssq_import_codes": (_i32, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _f, _i32, _p]),
    "ssq_stage_rows_h2d": (_i32, [_p, _p, _p, _i64, _i64, _p]),
    "ssq_pull_rows_host": (_i32, [_p, _p, _p, _i64, _i64, _p, _i64, _i64, _i32, _p]),
    "ssq_pull_rows_host_packed": (_i32, [_p, _p, _p, _p, _p, _i64, _i64, _p, _i64, _i64, _i32, _p]),
    "ssq_loop_advance": (_i32, [_p, _p, _p, _i32, _p, _p, _p, _p, _i64, _p]),
}

_lib = None


class SsqError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library once; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SsqError(
            f"{LIB_PATH} is missing: build it with `python -m shiftedscalequantization_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the quantiser kernels."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().ssq_status_string(status).decode()
        raise SsqError(f"{what} failed with status {status}: {msg}")
