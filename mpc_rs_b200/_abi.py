"""ctypes declarations for libmpc_b200.so (include/mpc_b200.h).  No compute happens in Python.

The library is CUDA-only: loading works without a GPU (symbol checks, argument validation), but every
create() fails with MPCB_CUDA_ERROR / MPCB_BAD_ARG when no device is present — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPCB_LIB_PATH") or os.path.join(_HERE, "libmpc_b200.so")  # override: developer A/B builds

# ---- enums (include/mpc_b200.h) ----
OK, NO_FINITE_COST, SUM_ZERO, U_INVALID, INVERSE_FAIL, CHOLESKY_FAIL, BAD_ARG, CUDA_ERROR, NCCL_ERROR, NOT_PREDICTED, \
    PEER_TIMEOUT, RTC_ERROR = range(12)
MODEL_L, MODEL_NL, MODEL_NL6, MODEL_USER = 0, 1, 2, 3
USER_PARAMS = 24
MODEL_PEN_LIN, MODEL_PEN_NL, MODEL_PEN6, MODEL_NL6_UKF, MODEL_USER_UKF = 16, 17, 18, 19, 20
F32, F64, F64_FAST = 0, 1, 2
PRECISIONS = {"f32": F32, "f64": F64, "f64fast": F64_FAST}
DT_F32, DT_F64 = 0, 1
SQRT_CHOLESKY, SQRT_EIG = 0, 1
ORDER_LIBRARY, ORDER_INTERLEAVED = 0, 1


class ModelParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("m1", "r_w", "m2", "l", "j1", "j2", "g", "kt", "dt")] + [
        ("cost", C.c_double * 12)
    ]


class MppiCfg(C.Structure):
    _fields_ = [
        ("model_id", C.c_int32),
        ("precision", C.c_int32),
        ("horizon", C.c_int32),
        ("state_dim", C.c_int32),
        ("samples", C.c_int64),
        ("controllers", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
        ("lambda_", C.c_double),
        ("std_dev", C.c_double),
        ("limit_lo", C.c_double),
        ("limit_hi", C.c_double),
        ("seed", C.c_uint64),
        ("keep_costs", C.c_int32),
        ("reserved", C.c_int32),
        ("model", ModelParams),
    ]


class MppiInfo(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("reserved", C.c_int32),
        ("argmax", C.c_int64),
        ("max", C.c_double),
        ("sum", C.c_double),
        ("n_finite", C.c_int64),
    ]


class UkfCfg(C.Structure):
    _fields_ = [
        ("model_id", C.c_int32),
        ("n", C.c_int32),
        ("o", C.c_int32),
        ("sqrt_mode", C.c_int32),
        ("sigma_order", C.c_int32),
        ("device", C.c_int32),
        ("exact", C.c_int32),
        ("reserved", C.c_int32),
        ("batch", C.c_int64),
        ("model", ModelParams),
    ]


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_H = C.c_void_p  # opaque handles

# name -> (restype, argtypes); every symbol include/mpc_b200.h declares
class ClosedLoopCfg(C.Structure):
    """mpcb_closed_loop_cfg (include/mpc_b200.h)."""
    _fields_ = [("controllers", C.c_int64), ("samples", C.c_int64), ("controller_offset", C.c_int64), ("tick_dt", C.c_double),
                ("seed", C.c_uint64), ("use_estimate", C.c_int32), ("precision", C.c_int32), ("exact_ukf", C.c_int32),
                ("device", C.c_int32)]


SYMBOLS = {
    "mpcb_status_string": (C.c_char_p, [C.c_int]),
    "mpcb_last_error_string": (C.c_char_p, []),
    "mpcb_abi_version": (C.c_int, []),
    "mpcb_device_count": (C.c_int, []),
    "mpcb_model_defaults": (C.c_int, [C.c_int32, C.POINTER(ModelParams)]),
    "mpcb_mppi_default_cfg": (C.c_int, [C.c_int32, C.POINTER(MppiCfg)]),
    "mpcb_mppi_create": (C.c_int, [C.POINTER(_H), C.POINTER(MppiCfg)]),
    "mpcb_mppi_create_user": (C.c_int, [C.POINTER(_H), C.POINTER(MppiCfg), C.c_char_p, C.POINTER(C.c_double), C.c_int32]),
    "mpcb_mppi_check_user_source": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32]),
    "mpcb_rtc_log": (C.c_char_p, []),
    "mpcb_mppi_destroy": (None, [_H]),
    "mpcb_mppi_compute": (C.c_int, [_H, _dp, _dp, _dp, C.POINTER(MppiInfo)]),
    "mpcb_mppi_compute_replay": (C.c_int, [_H, _dp, _dp, _vp, C.c_int32, C.c_int32, _dp, C.POINTER(MppiInfo)]),
    "mpcb_mppi_compute_dump": (C.c_int, [_H, _dp, _dp, _vp, _dp, C.POINTER(MppiInfo)]),
    "mpcb_mppi_get_costs": (C.c_int, [_H, _dp]),
    "mpcb_mppi_compute_device": (C.c_int, [_H, _vp, _vp, _vp, C.c_int32, _vp]),
    "mpcb_mppi_sync": (C.c_int, [_H]),
    "mpcb_mppi_last_info": (C.c_int, [_H, C.POINTER(MppiInfo)]),
    "mpcb_mppi_stream": (C.c_void_p, [_H]),
    "mpcb_mppi_launches": (C.c_int64, [_H]),
    "mpcb_mppi_local_samples": (C.c_int64, [_H]),
    "mpcb_mppi_partial_len": (C.c_int32, [_H]),
    "mpcb_mppi_compute_partial": (C.c_int, [_H, _dp, _dp, _vp, C.c_int32, _vp]),
    "mpcb_mppi_combine": (C.c_int, [_H, _vp, C.c_int32, _dp, C.POINTER(MppiInfo)]),
    "mpcb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "mpcb_mppi_attach_comm": (C.c_int, [_H, C.c_char_p]),
    "mpcb_mppi_first_control_device": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "mpcb_ukf_set_enable": (C.c_int, [_H, C.c_uint32]),
    "mpcb_ukf_gen_r": (C.c_int, [_H, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mpcb_ukf_gather_state_device": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_int32), C.c_void_p]),
    "mpcb_mppi_peer_handle": (C.c_int, [_H, C.c_char_p]),
    "mpcb_mppi_attach_peers": (C.c_int, [_H, C.c_char_p]),
    "mpcb_ukf_default_cfg": (C.c_int, [C.c_int32, C.POINTER(UkfCfg)]),
    "mpcb_ukf_default_noise": (C.c_int, [C.c_int32, C.c_double, _dp, _dp, _dp]),
    "mpcb_ukf_create": (C.c_int, [C.POINTER(_H), C.POINTER(UkfCfg)]),
    "mpcb_ukf_create_user": (C.c_int, [C.POINTER(_H), C.POINTER(UkfCfg), C.c_char_p, C.POINTER(C.c_double), C.c_int32]),
    "mpcb_ukf_check_user_source": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32]),
    "mpcb_ukf_destroy": (None, [_H]),
    "mpcb_ukf_init": (C.c_int, [_H, _dp, _dp, _dp, _dp]),
    "mpcb_ukf_set_state": (C.c_int, [_H, _dp, _dp]),
    "mpcb_ukf_get_state": (C.c_int, [_H, _dp, _dp]),
    "mpcb_ukf_get_state_range": (C.c_int, [_H, C.c_int64, C.c_int64, _dp, _dp]),
    "mpcb_ukf_set_q": (C.c_int, [_H, _dp]),
    "mpcb_ukf_set_r": (C.c_int, [_H, _dp]),
    "mpcb_ukf_predict": (C.c_int, [_H, _dp, C.c_double, C.c_double]),
    "mpcb_ukf_update": (C.c_int, [_H, _dp]),
    "mpcb_ukf_step": (C.c_int, [_H, _dp, C.c_double, C.c_double, _dp]),
    "mpcb_ukf_run_device": (C.c_int, [_H, C.c_int32, _vp, C.c_double, C.c_double, _vp]),
    "mpcb_ukf_sync": (C.c_int, [_H]),
    "mpcb_ukf_get_status": (C.c_int, [_H, C.POINTER(C.c_int32)]),
    "mpcb_ukf_stream": (C.c_void_p, [_H]),
    "mpcb_ukf_launches": (C.c_int64, [_H]),
    "mpcb_ukf_device_x": (C.c_void_p, [_H]),
    "mpcb_ukf_device_p": (C.c_void_p, [_H]),
    "mpcb_mppi_set_controller_offset": (C.c_int, [_H, C.c_int64]),
    "mpcb_closed_loop_default_cfg": (C.c_int, [C.POINTER(ClosedLoopCfg)]),
    "mpcb_closed_loop_create": (C.c_int, [C.POINTER(_H), C.POINTER(ClosedLoopCfg)]),
    "mpcb_closed_loop_destroy": (None, [_H]),
    "mpcb_closed_loop_set_state": (C.c_int, [_H, _dp, _dp]),
    "mpcb_closed_loop_set_truth": (C.c_int, [_H, _dp]),
    "mpcb_closed_loop_set_controls": (C.c_int, [_H, _dp]),
    "mpcb_closed_loop_tick": (C.c_int, [_H, C.c_int32]),
    "mpcb_closed_loop_tick_replay": (C.c_int, [_H, _dp, _vp, C.c_int32]),
    "mpcb_closed_loop_sync": (C.c_int, [_H]),
    "mpcb_closed_loop_get": (C.c_int, [_H, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int32)]),
    "mpcb_closed_loop_mppi": (C.c_void_p, [_H]),
    "mpcb_closed_loop_ukf": (C.c_void_p, [_H]),
    "mpcb_closed_loop_ticks": (C.c_int64, [_H]),
    "mpcb_closed_loop_launches": (C.c_int64, [_H]),
    "mpcb_device_alloc": (C.c_int, [C.c_int32, C.c_uint64, C.POINTER(C.c_void_p)]),
    "mpcb_device_free": (C.c_int, [C.c_int32, _vp]),
    "mpcb_device_upload": (C.c_int, [C.c_int32, _vp, _vp, C.c_uint64]),
    "mpcb_device_download": (C.c_int, [C.c_int32, _vp, _vp, C.c_uint64]),
}

_lib = None


class MpcB200Error(RuntimeError):
    """A call into libmpc_b200.so failed; .status is the mpcb_status."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def lib():
    """Loads libmpc_b200.so (built in-tree by `make -C mpc_rs_b200/csrc` or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpcB200Error(CUDA_ERROR, f"{LIB_PATH} is missing: build it with `make -C mpc_rs_b200/csrc` "
                                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if L.mpcb_abi_version() != 1:
        raise MpcB200Error(BAD_ARG, "libmpc_b200.so ABI version mismatch")
    _lib = L
    return L


def status_string(st: int) -> str:
    return lib().mpcb_status_string(st).decode()


def check(st: int, allowed=(OK,)):
    if st in allowed:
        return st
    detail = lib().mpcb_last_error_string().decode()
    msg = status_string(st)
    raise MpcB200Error(st, f"{msg}" + (f": {detail}" if detail and st in (BAD_ARG, CUDA_ERROR, NCCL_ERROR) else ""))
