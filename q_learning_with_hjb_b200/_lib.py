"""ctypes binding of libhjb_b200.so (the C ABI declared in include/hjb_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the symbol
table (``lib()``), and every compute entry point raises ``RuntimeError`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HJB_MAX_N = 10
HJB_MAX_M = 3

# hjb_system_kind
SYS_LINEAR, SYS_CARTPOLE, SYS_ACROBOT, SYS_QUAD2D, SYS_QUAD10D = range(5)
# hjb_control_kind
CTL_FEEDBACK, CTL_CARTPOLE_ES, CTL_ACROBOT_ES, CTL_TRACK, CTL_SWITCH_CURVE, CTL_GRID_SIGN = range(6)
# hjb_integrator
INT_EULER, INT_RK4, INT_DISCRETE = range(3)
INTEGRATORS = {"euler": INT_EULER, "rk4": INT_RK4, "discrete": INT_DISCRETE}

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhjb_b200.so")
_lib = None


class HjbSystem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("m", C.c_int32), ("dt", C.c_float),
                ("umin", C.c_float * HJB_MAX_M), ("umax", C.c_float * HJB_MAX_M),
                ("par", C.c_float * 8), ("A", C.c_float * 16), ("B", C.c_float * 8)]


class HjbControl(C.Structure):
    _fields_ = [("kind", C.c_int32), ("clip", C.c_int32),
                ("K", C.c_float * (HJB_MAX_M * HJB_MAX_N)), ("P", C.c_float * 16),
                ("xf", C.c_float * HJB_MAX_N), ("uf", C.c_float * HJB_MAX_M), ("aux", C.c_float * 8),
                ("ref", C.c_void_p), ("ref_steps", C.c_int32), ("ref_offset", C.c_int32)]


class HjbCost(C.Structure):
    _fields_ = [("Q", C.c_float * (HJB_MAX_N * HJB_MAX_N)), ("R", C.c_float * (HJB_MAX_M * HJB_MAX_M)),
                ("xf", C.c_float * HJB_MAX_N), ("uf", C.c_float * HJB_MAX_M)]


class HjbRolloutOpts(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("record_stride", C.c_int32), ("fast_trig", C.c_int32),
                ("box_enabled", C.c_int32),
                ("box_xf", C.c_float * HJB_MAX_N), ("box_lo", C.c_float * HJB_MAX_N), ("box_hi", C.c_float * HJB_MAX_N)]


ACTIVATIONS = {"relu": 0, "tanh": 1, "sin": 2}
U_CLIPPED, U_BANGBANG = 0, 1
RES_NORMALIZED, RES_MIN_TIME = 0, 1


class HjbVnet(C.Structure):
    _fields_ = [("n", C.c_int32), ("act", C.c_int32), ("features", C.c_int32 * 3), ("impl", C.c_int32),
                ("params", C.c_void_p),
                ("mean", C.c_float * HJB_MAX_N), ("std", C.c_float * HJB_MAX_N), ("xf", C.c_float * HJB_MAX_N),
                ("eps_s", C.c_float)]


class HjbTask(C.Structure):
    _fields_ = [("Q", C.c_float * (HJB_MAX_N * HJB_MAX_N)), ("R", C.c_float * (HJB_MAX_M * HJB_MAX_M)),
                ("Rinv", C.c_float * (HJB_MAX_M * HJB_MAX_M)), ("uf", C.c_float * HJB_MAX_M),
                ("eps", C.c_float), ("control_form", C.c_int32), ("residual_form", C.c_int32)]


class HjbSoftPD(C.Structure):
    _fields_ = [("n", C.c_int32), ("act", C.c_int32), ("normalized_residual", C.c_int32), ("params", C.c_void_p),
                ("xf", C.c_float * HJB_MAX_N), ("uf", C.c_float * HJB_MAX_M),
                ("Q", C.c_float * (HJB_MAX_N * HJB_MAX_N)), ("R", C.c_float * (HJB_MAX_M * HJB_MAX_M)),
                ("Rinv", C.c_float * (HJB_MAX_M * HJB_MAX_M)), ("K", C.c_float * (HJB_MAX_M * HJB_MAX_N)),
                ("P", C.c_float * (HJB_MAX_N * HJB_MAX_N)), ("eps", C.c_float)]


# every symbol include/hjb_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "hjb_abi_version": (C.c_int, []),
    "hjb_status_string": (C.c_char_p, [C.c_int]),
    "hjb_rollout": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbControl), C.POINTER(HjbCost),
                              C.POINTER(HjbRolloutOpts), _P, C.c_int64, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "hjb_rollout_variant": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbControl), C.POINTER(HjbCost),
                                      C.POINTER(HjbRolloutOpts), C.c_int32, C.POINTER(C.c_int32)]),
    "hjb_dynamics": (C.c_int, [C.POINTER(HjbSystem), C.c_int32, C.c_int32, _P, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "hjb_control_efforts": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbControl), C.c_int32, _P, C.c_int64, _P, _P]),
    "hjb_states_wrap": (C.c_int, [C.POINTER(HjbSystem), _P, C.c_int64, _P]),
    "hjb_time_to_goal": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, _P, _P]),
    "hjb_sample_states": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint64, C.c_int64,
                                    C.c_int64, _P, _P]),
    "hjb_fma_peak_probe": (C.c_int, [_P, C.c_int64, C.c_int32, C.POINTER(C.c_double), _P]),
    "hjb_vhjb_param_count": (C.c_int64, [C.c_int32]),
    "hjb_vhjb_workspace_bytes": (C.c_int64, [C.c_int32]),
    "hjb_vhjb_count": (C.c_int, [_P, C.c_int64, C.c_float, _P, _P, _P]),
    "hjb_vhjb_residual": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbVnet), C.POINTER(HjbTask), _P, _P, _P, C.c_int64,
                                    _P, _P, _P, _P, _P, _P, _P]),
    "hjb_vhjb_loss_grad": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbVnet), C.POINTER(HjbTask), _P, _P, _P, C.c_int64,
                                     _P, C.c_float, _P, _P, _P, _P]),
    "hjb_vhjb_loss_grad_accumulate": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbVnet), C.POINTER(HjbTask), _P, _P, _P,
                                                C.c_int64, _P, C.c_float, _P, _P, _P, _P]),
    "hjb_vhjb_stream_batch": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64, _P, _P, _P]),
    "hjb_vhjb_loss_grad_streamed": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbVnet), C.POINTER(HjbTask), _P, _P, _P,
                                              C.c_int64, _P, C.c_float, _P, _P, _P, _P, C.c_int64, _P]),
    "hjb_vhjb_train_step": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbVnet), C.POINTER(HjbTask), _P, _P, _P, C.c_int64,
                                      C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, _P, _P, _P, _P, _P,
                                      _P, _P, _P]),
    "hjb_vhjb_stream_failures": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P]),
    "hjb_vhjb_adam_guarded": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                        _P, _P]),
    "hjb_vhjb_peer_exchange_floats": (C.c_int64, [C.c_int32, C.c_int32]),
    "hjb_vhjb_peer_exchange_flags": (C.c_int64, [C.c_int32, C.c_int32]),
    "hjb_vhjb_train_step_peer": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbVnet), C.POINTER(HjbTask), _P, _P, _P, C.c_int64,
                                           C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, _P, _P, _P, _P, _P,
                                           _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "hjb_vhjb_saturation": (C.c_int, [_P, C.c_int32, _P, _P]),
    "hjb_vhjb_deferred": (C.c_int, [_P, C.c_int32, _P, _P]),
    "hjb_vhjb_saturation_total": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P]),
    "hjb_policy_step": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbTask), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                  C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, _P, _P, _P, _P, _P, _P, _P,
                                  C.c_int64, _P]),
    "hjb_policy_rollout": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbVnet), C.POINTER(HjbTask), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, _P, _P, _P, _P,
                                     _P, _P, _P, _P, _P, C.c_int64, _P, _P]),
    "hjb_replay_append": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, _P, _P,
                                    _P, _P]),
    "hjb_replay_gather": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "hjb_softpd_param_count": (C.c_int64, [C.c_int32]),
    "hjb_softpd_workspace_bytes": (C.c_int64, [C.c_int32]),
    "hjb_softpd_policy": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbSoftPD), _P, C.c_int64, _P, _P, _P, _P, _P]),
    "hjb_softpd_loss_grad": (C.c_int, [C.POINTER(HjbSystem), C.POINTER(HjbSoftPD), _P, C.c_int64, C.c_int32, C.c_float, _P, _P,
                                       _P, _P]),
    "hjb_adam": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, _P]),
}


def c_floats(values, count: int):
    """A ctypes float[count] filled from ``values`` (host-side arguments of the C ABI)."""
    import numpy as np
    a = np.zeros(count, dtype=np.float32)
    v = np.asarray(values, dtype=np.float32).reshape(-1)
    a[: v.size] = v[:count]
    return (C.c_float * count)(*a.tolist())


def lib_path() -> str:
    return _LIB_PATH


def lib():
    """Load libhjb_b200.so (once) and type every exported symbol.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} not found: build it with `python -m q_learning_with_hjb_b200.build` "
                "(there is no CPU fallback for the CUDA hot path)")
        handle = C.CDLL(_LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status: int, what: str = "hjb call"):
    if status != 0:
        msg = lib().hjb_status_string(status)
        raise RuntimeError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


def fill(carray, values):
    """Copy a (possibly nested) sequence / ndarray into a ctypes float array, zero-padding the tail."""
    flat = np.asarray(values, dtype=np.float32).reshape(-1)
    if flat.size > len(carray):
        raise ValueError(f"{flat.size} values do not fit a ctypes array of {len(carray)}")
    for i in range(len(carray)):
        carray[i] = float(flat[i]) if i < flat.size else 0.0


# ------------------------------------------------------------------------------------------------
# device plumbing (torch owns memory and streams)
# ------------------------------------------------------------------------------------------------

def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("q_learning_with_hjb_b200: no CUDA device — the rollout / vhjb hot path is CUDA-only "
                           "(sm_100a); there is no CPU fallback")
    return torch


def dev_f32(x, shape=None):
    """Host or device array -> contiguous fp32 CUDA tensor on the current device."""
    torch = require_cuda()
    if isinstance(x, torch.Tensor):
        t = x.to(device="cuda", dtype=torch.float32)
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to("cuda")
    t = t.contiguous()
    if shape is not None:
        t = t.reshape(shape)
    return t


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
