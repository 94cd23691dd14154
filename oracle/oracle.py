"""ctypes binding of the CPU oracle (oracle/rc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product package dynenv_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from dynenv_b200 import layout as L

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_DIR, "_build", "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _DIR] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.rc_create.restype = C.c_void_p
        _LIB.rc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint,
                                   C.c_uint64, C.c_uint32]
        for name in ("rc_destroy", "rc_reset"):
            getattr(_LIB, name).argtypes = [C.c_void_p]
            getattr(_LIB, name).restype = None
        _LIB.rc_set_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _LIB.rc_set_tape.restype = None
        for name in ("rc_tape_pos", "rc_tape_underflow", "rc_arb_overflow"):
            getattr(_LIB, name).argtypes = [C.c_void_p]
            getattr(_LIB, name).restype = C.c_int
        _LIB.rc_store_state.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.rc_load_state.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.rc_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.rc_reset_obs.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.rc_stepout_size.restype = C.c_int
        assert _LIB.rc_stepout_size() == C.sizeof(RcStepOut), (
            _LIB.rc_stepout_size(), C.sizeof(RcStepOut))
    return _LIB


_T, _P = L.RC_NTICK, L.RC_MAXP


def _arr(ctype, *shape):
    t = ctype
    for s in reversed(shape):
        t = t * s
    return t


class RcStepOut(C.Structure):
    _fields_ = [
        ("ball", _arr(C.c_double, _T, _P, L.RC_CAP_BALL, L.RC_F_BALL)),
        ("robot", _arr(C.c_double, _T, _P, L.RC_CAP_ROBOT, L.RC_F_ROBOT)),
        ("post", _arr(C.c_double, _T, _P, L.RC_CAP_POST, L.RC_F_POST)),
        ("cross", _arr(C.c_double, _T, _P, L.RC_CAP_CROSS, L.RC_F_CROSS)),
        ("fcross", _arr(C.c_double, _T, _P, L.RC_CAP_FCROSS, L.RC_F_FCROSS)),
        ("line", _arr(C.c_double, _T, _P, L.RC_CAP_LINE, L.RC_F_LINE)),
        ("counts", _arr(C.c_int32, _T, _P, L.RC_NTYPES)),
        ("num_landmarks", _arr(C.c_int32, _T, _P)),
        ("robots_seen", _arr(C.c_int32, _T, _P, _P - 1)),
        ("balls_seen", _arr(C.c_int32, _T, _P)),
        ("full_ball", _arr(C.c_double, _T, _P, 4)),
        ("full_self", _arr(C.c_double, _T, _P, 8)),
        ("full_others", _arr(C.c_double, _T, _P, _P - 1, 6)),
        ("recon_ball", _arr(C.c_double, _P, 4)),
        ("recon_self", _arr(C.c_double, _P, 8)),
        ("recon_others", _arr(C.c_double, _P, _P - 1, 6)),
        ("state_robots", _arr(C.c_double, _P, 6)),
        ("state_ball", _arr(C.c_double, 3)),
        ("rewards", _arr(C.c_double, _P)),
        ("done", C.c_int32),
        ("episode_r", _arr(C.c_double, _P)),
        ("episode_p_r", _arr(C.c_double, _P)),
        ("episode_o_r", _arr(C.c_double, _P)),
        ("episode_g", _arr(C.c_int32, 2)),
    ]

    def to_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = int(v) if isinstance(v, int) else np.ctypeslib.as_array(v).copy()
        return out


class RcRawVision(C.Structure):
    _fields_ = [("ball", _arr(C.c_double, L.RC_CAP_BALL, 3)), ("robot", _arr(C.c_double, L.RC_CAP_ROBOT, 3)),
                ("post", _arr(C.c_double, L.RC_CAP_POST, 3)), ("cross", _arr(C.c_double, L.RC_CAP_CROSS, 3)),
                ("fcross", _arr(C.c_double, L.RC_CAP_FCROSS, 3)), ("line", _arr(C.c_double, L.RC_CAP_LINE, 4)),
                ("circle", _arr(C.c_double, 4)), ("counts", _arr(C.c_int32, L.RC_NTYPES)), ("pad", _arr(C.c_int32, 2))]

    def to_dict(self):
        return {name: np.ctypeslib.as_array(getattr(self, name)).copy() for name, _ in self._fields_ if name != "pad"}


class RcRawOut(C.Structure):
    _fields_ = [("v", (RcRawVision * _P) * _T)]


def raw_to_record(raw):
    """RcRawVision dict -> the float32 raw record of the C ABI (include/dynenv_b200.h, RC_R_*)"""
    rec = np.zeros(L.RC_RAW_WORDS, dtype=np.float32)
    ri = rec.view(np.int32)
    ri[L.RC_R_COUNTS:L.RC_R_COUNTS + L.RC_NTYPES] = raw["counts"]
    ri[L.RC_R_CIRCLE_SEEN] = int(raw["circle"][0] != 0)
    rec[L.RC_R_CIRCLE:L.RC_R_CIRCLE + 3] = raw["circle"][1:4]
    for name, off, cap in (("ball", L.RC_R_BALL, L.RC_CAP_BALL), ("robot", L.RC_R_ROBOT, L.RC_CAP_ROBOT), ("post", L.RC_R_POST, L.RC_CAP_POST),
                           ("cross", L.RC_R_CROSS, L.RC_CAP_CROSS), ("fcross", L.RC_R_FCROSS, L.RC_CAP_FCROSS)):
        rec[off:off + cap * 3] = raw[name].reshape(-1)
    rec[L.RC_R_LINE:L.RC_R_LINE + L.RC_CAP_LINE * 4] = raw["line"].reshape(-1)
    return rec


def record_to_raw(rec):
    """inverse of raw_to_record (values as float64 of the float32 words): what the raster oracle draws for a record"""
    rec = np.asarray(rec, dtype=np.float32)
    ri = rec.view(np.int32)
    raw = {"counts": ri[L.RC_R_COUNTS:L.RC_R_COUNTS + L.RC_NTYPES].copy(),
           "circle": np.concatenate([[float(ri[L.RC_R_CIRCLE_SEEN])], rec[L.RC_R_CIRCLE:L.RC_R_CIRCLE + 3].astype(np.float64)])}
    for name, off, cap in (("ball", L.RC_R_BALL, L.RC_CAP_BALL), ("robot", L.RC_R_ROBOT, L.RC_CAP_ROBOT), ("post", L.RC_R_POST, L.RC_CAP_POST),
                           ("cross", L.RC_R_CROSS, L.RC_CAP_CROSS), ("fcross", L.RC_R_FCROSS, L.RC_CAP_FCROSS)):
        raw[name] = rec[off:off + cap * 3].astype(np.float64).reshape(cap, 3)
    raw["line"] = rec[L.RC_R_LINE:L.RC_R_LINE + L.RC_CAP_LINE * 4].astype(np.float64).reshape(L.RC_CAP_LINE, 4)
    return raw


class RcOracle(object):
    """One RoboCup environment of the CPU oracle."""

    def __init__(self, nPlayers, obsType=1, noiseType=1, noiseMagnitude=2.0, allowHeadTurn=False,
                 flags=L.DYNENV_FLAG_CAN_FALL | L.DYNENV_FLAG_USE_OBS_REWARDS, seed=0, env_id=0):
        self._lib = lib()
        self._h = self._lib.rc_create(int(nPlayers), int(obsType), int(noiseType), float(noiseMagnitude),
                                      int(bool(allowHeadTurn)), int(flags), int(seed), int(env_id))
        self.nPlayers = min(int(nPlayers), 5)
        self.P = 2 * self.nPlayers
        self._tape = None
        self._out = RcStepOut()

    def __del__(self):
        try:
            self._lib.rc_destroy(self._h)
        except Exception:
            pass

    def set_tape(self, tape):
        if tape is None:
            self._tape = None
            self._lib.rc_set_tape(self._h, None, 0)
        else:
            self._tape = np.ascontiguousarray(tape, dtype=np.float64)
            self._lib.rc_set_tape(self._h, self._tape.ctypes.data, len(self._tape))

    def tape_pos(self):
        return self._lib.rc_tape_pos(self._h)

    def tape_underflow(self):
        return bool(self._lib.rc_tape_underflow(self._h))

    def debug_substep_range(self, begin, end):
        self._lib.rc_debug_substep_range(C.c_void_p(self._h), int(begin), int(end))

    def arb_overflow(self):
        return bool(self._lib.rc_arb_overflow(self._h))

    def capture_raw(self, enable=True):
        """the following reset_obs() / step() calls also record the raw detection lists (raw_vision())"""
        assert self._lib.rc_rawout_size() == C.sizeof(RcRawOut)
        self._raw = RcRawOut() if enable else None
        self._lib.rc_set_raw_out(C.c_void_p(self._h), C.byref(self._raw) if enable else None)

    def raw_vision(self, t, agent):
        return self._raw.v[t][agent].to_dict()

    def reset(self):
        self._lib.rc_reset(self._h)

    def reset_obs(self):
        self._lib.rc_reset_obs(self._h, C.byref(self._out))
        return self._out.to_dict()

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.P, 4)
        self._lib.rc_step(self._h, a.ctypes.data, C.byref(self._out))
        return self._out.to_dict()

    def get_state(self):
        w = np.zeros(L.RC_STATE_WORDS, dtype=np.float64)
        self._lib.rc_store_state(self._h, w.ctypes.data)
        return w

    def set_state(self, words):
        w = np.ascontiguousarray(pad_words(words), dtype=np.float64)
        assert w.shape == (L.RC_STATE_WORDS,)
        self._lib.rc_load_state(self._h, w.ctypes.data)


class DrvAgentObs(C.Structure):
    _fields_ = [("self_", _arr(C.c_double, L.DRV_F_SELF)), ("car", _arr(C.c_double, L.DRV_CAP_CAR, L.DRV_F_CAR)),
                ("obs", _arr(C.c_double, L.DRV_CAP_OBS, L.DRV_F_OBS)), ("ped", _arr(C.c_double, L.DRV_CAP_PED, L.DRV_F_PED)),
                ("lane", _arr(C.c_double, L.DRV_CAP_LANE, L.DRV_F_LANE_FULL)), ("counts", _arr(C.c_int32, 4))]


class DrvStepOut(C.Structure):
    _fields_ = [("agent", DrvAgentObs * L.DRV_MAXCARS), ("rewards", _arr(C.c_double, L.DRV_MAXCARS)), ("done", C.c_int32),
                ("episode_r", _arr(C.c_double, L.DRV_MAXCARS)), ("episode_p_r", _arr(C.c_double, L.DRV_MAXCARS)),
                ("episode_g", _arr(C.c_int32, 2)), ("state_cars", _arr(C.c_double, L.DRV_MAXCARS, 7)),
                ("state_obs", _arr(C.c_double, L.DRV_MAXOBS, 4)), ("state_ped", _arr(C.c_double, L.DRV_MAXPED, 2)),
                ("nobs", C.c_int32), ("nped", C.c_int32)]

    def to_dict(self):
        out = {}
        ag = self.agent
        for name, key in (("self_", "self"), ("car", "car"), ("obs", "obs"), ("ped", "ped"), ("lane", "lane"), ("counts", "counts")):
            out[key] = np.stack([np.ctypeslib.as_array(getattr(ag[i], name)).copy() for i in range(L.DRV_MAXCARS)])
        for name, _ in self._fields_[1:]:
            v = getattr(self, name)
            out[name] = int(v) if isinstance(v, int) else np.ctypeslib.as_array(v).copy()
        return out


class DrvOracle(object):
    """One Driving environment of the CPU oracle (oracle/drv_oracle.c)."""

    def __init__(self, nPlayers, obsType=1, noiseType=1, noiseMagnitude=2.0, continuous=False, flags=0, seed=0, env_id=0):
        lb = self._lib = lib()
        lb.drv_create.restype = C.c_void_p
        lb.drv_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint, C.c_uint64, C.c_uint32]
        for name in ("drv_destroy", "drv_reset"):
            getattr(lb, name).argtypes = [C.c_void_p]
            getattr(lb, name).restype = None
        lb.drv_set_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        for name in ("drv_tape_pos", "drv_tape_underflow", "drv_arb_overflow"):
            getattr(lb, name).argtypes = [C.c_void_p]
            getattr(lb, name).restype = C.c_int
        lb.drv_store_state.argtypes = [C.c_void_p, C.c_void_p]
        lb.drv_load_state.argtypes = [C.c_void_p, C.c_void_p]
        lb.drv_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lb.drv_reset_obs.argtypes = [C.c_void_p, C.c_void_p]
        lb.drv_debug_substep_range.argtypes = [C.c_void_p, C.c_int, C.c_int]
        assert lb.drv_stepout_size() == C.sizeof(DrvStepOut), (lb.drv_stepout_size(), C.sizeof(DrvStepOut))
        self._h = lb.drv_create(int(nPlayers), int(obsType), int(noiseType), float(noiseMagnitude), int(bool(continuous)),
                                int(flags), int(seed), int(env_id))
        self.N = min(int(nPlayers), L.DRV_MAXCARS)
        self._tape = None
        self._out = DrvStepOut()

    def __del__(self):
        try:
            self._lib.drv_destroy(self._h)
        except Exception:
            pass

    def set_tape(self, tape):
        if tape is None:
            self._tape = None
            self._lib.drv_set_tape(self._h, None, 0)
        else:
            self._tape = np.ascontiguousarray(tape, dtype=np.float64)
            self._lib.drv_set_tape(self._h, self._tape.ctypes.data, len(self._tape))

    def tape_pos(self):
        return self._lib.drv_tape_pos(self._h)

    def tape_underflow(self):
        return bool(self._lib.drv_tape_underflow(self._h))

    def arb_overflow(self):
        return bool(self._lib.drv_arb_overflow(self._h))

    def debug_substep_range(self, begin, end):
        self._lib.drv_debug_substep_range(self._h, int(begin), int(end))

    def reset(self):
        self._lib.drv_reset(self._h)

    def reset_obs(self):
        self._lib.drv_reset_obs(self._h, C.byref(self._out))
        return self._out.to_dict()

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.N, 2)
        self._lib.drv_step(self._h, a.ctypes.data, C.byref(self._out))
        return self._out.to_dict()

    def get_state(self):
        w = np.zeros(L.DRV_STATE_WORDS, dtype=np.float64)
        self._lib.drv_store_state(self._h, w.ctypes.data)
        return w

    def set_state(self, words):
        w = np.ascontiguousarray(words, dtype=np.float64)
        assert w.shape == (L.DRV_STATE_WORDS,)
        self._lib.drv_load_state(self._h, w.ctypes.data)


def drv_words_to_f32_record(words):
    """Driving: float64 word list -> uint32 record (two-float coordinates split into hi + lo, float64 car angles
    into their high / low 32 bits)"""
    w = np.array(words, dtype=np.float64)
    for hi, lo in L.DRV_SPLIT.items():
        (oh, cnt, _), (ol, _, _) = L.DRV_FIELDS[hi], L.DRV_FIELDS[lo]
        v = w[..., oh:oh + cnt] + w[..., ol:ol + cnt]
        h = v.astype(np.float32).astype(np.float64)
        w[..., oh:oh + cnt] = h
        w[..., ol:ol + cnt] = v - h
    mask = np.array(L.drv_int_mask())
    rec = np.zeros(w.shape, dtype=np.uint32)
    rec[..., mask] = w[..., mask].astype(np.int64).astype(np.int32).view(np.uint32)
    rec[..., ~mask] = w[..., ~mask].astype(np.float32).view(np.uint32)
    for hi, lo in L.DRV_DOUBLE.items():
        (oh, cnt, _), (ol, _, _) = L.DRV_FIELDS[hi], L.DRV_FIELDS[lo]
        bits = np.ascontiguousarray(w[..., oh:oh + cnt] + w[..., ol:ol + cnt]).view(np.uint64)
        rec[..., oh:oh + cnt] = (bits >> np.uint64(32)).astype(np.uint32)
        rec[..., ol:ol + cnt] = (bits & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return rec


def drv_f32_record_to_words(rec):
    r = np.asarray(rec, dtype=np.uint32)
    mask = np.array(L.drv_int_mask())
    w = np.zeros(r.shape, dtype=np.float64)
    w[..., mask] = r[..., mask].view(np.int32).astype(np.float64)
    with np.errstate(invalid="ignore"):  # the float64 angle words are not float32 numbers; they are overwritten below
        w[..., ~mask] = r[..., ~mask].view(np.float32).astype(np.float64)
    for hi, lo in L.DRV_SPLIT.items():
        (oh, cnt, _), (ol, _, _) = L.DRV_FIELDS[hi], L.DRV_FIELDS[lo]
        w[..., oh:oh + cnt] += w[..., ol:ol + cnt]
        w[..., ol:ol + cnt] = 0
    for hi, lo in L.DRV_DOUBLE.items():
        (oh, cnt, _), (ol, _, _) = L.DRV_FIELDS[hi], L.DRV_FIELDS[lo]
        bits = (r[..., oh:oh + cnt].astype(np.uint64) << np.uint64(32)) | r[..., ol:ol + cnt].astype(np.uint64)
        w[..., oh:oh + cnt] = np.ascontiguousarray(bits).view(np.float64)
        w[..., ol:ol + cnt] = 0
    return w


def pad_words(words):
    """float64 word lists of an older (shorter) record layout -> current width (new fields are zero)."""
    w = np.asarray(words, dtype=np.float64)
    if w.shape[-1] == L.RC_STATE_WORDS:
        return w
    out = np.zeros(w.shape[:-1] + (L.RC_STATE_WORDS,), dtype=np.float64)
    out[..., :w.shape[-1]] = w
    return out


def words_to_f32_record(words):
    """float64 word list -> the uint32 record the C-ABI uses: floats rounded to float32, two-float
    coordinates (layout.RC_SPLIT) split into hi = f32(v), lo = f32(v - hi)."""
    w = pad_words(words).copy()
    for hi, lo in L.RC_SPLIT.items():
        (oh, cnt, _), (ol, _, _) = L.RC_FIELDS[hi], L.RC_FIELDS[lo]
        v = w[..., oh:oh + cnt] + w[..., ol:ol + cnt]
        h = v.astype(np.float32).astype(np.float64)
        w[..., oh:oh + cnt] = h
        w[..., ol:ol + cnt] = v - h
    mask = np.array(L.rc_int_mask())
    rec = np.zeros(w.shape, dtype=np.uint32)
    rec[..., mask] = w[..., mask].astype(np.int64).astype(np.int32).view(np.uint32)
    rec[..., ~mask] = w[..., ~mask].astype(np.float32).view(np.uint32)
    return rec


def f32_record_to_words(rec):
    """inverse of words_to_f32_record; two-float coordinates are merged into the hi word (lo = 0)"""
    r = np.asarray(rec, dtype=np.uint32)
    mask = np.array(L.rc_int_mask())
    w = np.zeros(r.shape, dtype=np.float64)
    w[..., mask] = r[..., mask].view(np.int32).astype(np.float64)
    w[..., ~mask] = r[..., ~mask].view(np.float32).astype(np.float64)
    for hi, lo in L.RC_SPLIT.items():
        (oh, cnt, _), (ol, _, _) = L.RC_FIELDS[hi], L.RC_FIELDS[lo]
        w[..., oh:oh + cnt] += w[..., ol:ol + cnt]
        w[..., ol:ol + cnt] = 0
    return w
