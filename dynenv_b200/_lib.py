"""ctypes binding of the C ABI (include/dynenv_b200.h).  There is no CPU fallback: a missing
library or a machine without a CUDA device raises."""
import ctypes as C
import os

from . import layout as L

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DYNENV_B200_LIB") or os.path.join(_DIR, "libdynenv_b200.so")
_LIB = None

EXPORTS = ["dynenv_abi_version", "dynenv_last_error", "dynenv_create", "dynenv_destroy", "dynenv_state_words",
           "dynenv_reset", "dynenv_step", "dynenv_step_u8", "dynenv_set_seed", "dynenv_get_state", "dynenv_set_state",
           "dynenv_last_launch_count", "dynenv_debug_substep_range", "dynenv_debug_timing",
           "dynenv_last_kernel_ms", "dynenv_render_images", "dynenv_rollout_insert", "dynenv_discount_rewards", "dynenv_debug_fp32_peak",
           "dynenv_arrange_inputs", "dynenv_arrange_outputs"]


class DynEnvConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("kind", C.c_int32), ("n_envs", C.c_int32), ("n_players", C.c_int32),
                ("obs_type", C.c_int32), ("noise_type", C.c_int32), ("noise_magnitude", C.c_float),
                ("continuous", C.c_int32), ("flags", C.c_uint32), ("seed", C.c_uint64),
                ("env_id_offset", C.c_int64), ("device", C.c_int32)]


class DynEnvRoboCupOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("full", C.c_void_p), ("info", C.c_void_p), ("raw", C.c_void_p)]


class DynEnvDrivingOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("info", C.c_void_p)]


class DynEnvArrangeLayout(C.Structure):
    _fields_ = [("n_types", C.c_int32), ("rec_words", C.c_int32), ("off", C.c_int32 * 8), ("cap", C.c_int32 * 8),
                ("feat", C.c_int32 * 8), ("stride", C.c_int32 * 8), ("count_word", C.c_int32 * 8)]


class DynEnvError(Exception):
    def __init__(self, code, msg):
        Exception.__init__(self, msg)
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("dynenv_b200: %s is missing; build it with `python -m dynenv_b200.build` "
                               "(there is no CPU fallback)" % LIB_PATH)
        lb = C.CDLL(LIB_PATH)
        lb.dynenv_abi_version.restype = C.c_int
        lb.dynenv_last_error.restype = C.c_char_p
        lb.dynenv_create.argtypes = [C.POINTER(DynEnvConfig), C.POINTER(C.c_void_p)]
        lb.dynenv_destroy.argtypes = [C.c_void_p]
        lb.dynenv_state_words.argtypes = [C.c_void_p]
        lb.dynenv_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lb.dynenv_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lb.dynenv_step_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lb.dynenv_set_seed.argtypes = [C.c_void_p, C.c_uint64]
        lb.dynenv_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lb.dynenv_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lb.dynenv_last_launch_count.argtypes = [C.c_void_p]
        lb.dynenv_debug_substep_range.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lb.dynenv_debug_timing.argtypes = [C.c_void_p, C.c_int]
        lb.dynenv_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        lb.dynenv_render_images.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        lb.dynenv_rollout_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        lb.dynenv_discount_rewards.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                               C.c_int64, C.c_float, C.c_void_p]
        lb.dynenv_debug_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_float)]
        lb.dynenv_arrange_inputs.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(DynEnvArrangeLayout),
                                             C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lb.dynenv_arrange_outputs.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        if lb.dynenv_abi_version() != L.DYNENV_ABI_VERSION:
            raise RuntimeError("dynenv_b200: library ABI %d != header ABI %d" % (lb.dynenv_abi_version(),
                                                                                 L.DYNENV_ABI_VERSION))
        _LIB = lb
    return _LIB


def check(rc):
    if rc != 0:
        raise DynEnvError(rc, (lib().dynenv_last_error() or b"").decode())
