"""Batched simulator handles: nEnvs environments stepped by one kernel launch.

Host side only: PyTorch owns the device buffers and the stream, the simulation runs in
dynenv_b200/libdynenv_b200.so through the C ABI (include/dynenv_b200.h).  Outputs are
padded-with-counts device tensors (views into packed per-observer records); the ragged
nested-list format of the reference is produced lazily by `dynenv_b200.reference_format`.
"""
import ctypes as C

import torch

from . import _lib
from . import layout as L
from .enums import NoiseType, ObservationType

_TYPES = [("ball", L.RC_O_BALL, L.RC_CAP_BALL, L.RC_F_BALL), ("robot", L.RC_O_ROBOT, L.RC_CAP_ROBOT, L.RC_F_ROBOT),
          ("post", L.RC_O_POST, L.RC_CAP_POST, L.RC_F_POST), ("cross", L.RC_O_CROSS, L.RC_CAP_CROSS, L.RC_F_CROSS),
          ("fcross", L.RC_O_FCROSS, L.RC_CAP_FCROSS, L.RC_F_FCROSS), ("line", L.RC_O_LINE, L.RC_CAP_LINE, L.RC_F_LINE)]


def _flags(randomInit, deterministicTurn, canFall, useObsRewards, auto_reset, sparse_obs=False):
    return ((L.DYNENV_FLAG_RANDOM_INIT if randomInit else 0) | (L.DYNENV_FLAG_DETERMINISTIC_TURN if deterministicTurn else 0)
            | (L.DYNENV_FLAG_CAN_FALL if canFall else 0) | (L.DYNENV_FLAG_USE_OBS_REWARDS if useObsRewards else 0)
            | (L.DYNENV_FLAG_AUTO_RESET if auto_reset else 0) | (L.DYNENV_FLAG_SPARSE_OBS if sparse_obs else 0))


class RoboCupOutputs(object):
    """Named views of one set of output buffers (no copies)."""

    def __init__(self, nEnvs, P, obs_type, device):
        T = L.RC_NTICK
        self.P = P
        self.obs = self.full = self.raw = None
        if obs_type == int(ObservationType.FULL):
            self.full = torch.zeros((nEnvs, T, P, L.RC_FULL_WORDS), dtype=torch.float32, device=device)
        else:
            self.obs = torch.zeros((nEnvs, T, P, L.RC_OBS_WORDS), dtype=torch.float32, device=device)
        if obs_type == int(ObservationType.IMAGE):  # raw detection records, drawn by render_images()
            self.raw = torch.zeros((nEnvs, T, P, L.RC_RAW_WORDS), dtype=torch.float32, device=device)
        self.info = torch.zeros((nEnvs, L.RC_INFO_WORDS), dtype=torch.float32, device=device)
        self.c_struct = _lib.DynEnvRoboCupOut(self.obs.data_ptr() if self.obs is not None else None,
                                              self.full.data_ptr() if self.full is not None else None,
                                              self.info.data_ptr(), self.raw.data_ptr() if self.raw is not None else None)
        self._build_views()

    def _build_views(self):
        P = self.P
        info_i = self.info.view(torch.int32)
        self.rewards = self.info[:, L.RC_I_REWARDS:L.RC_I_REWARDS + P]
        self.done = info_i[:, L.RC_I_DONE]
        self.diag = info_i[:, L.RC_I_FLAGS]
        self.episode_r = self.info[:, L.RC_I_EPR:L.RC_I_EPR + P]
        self.episode_p_r = self.info[:, L.RC_I_EPPR:L.RC_I_EPPR + P]
        self.episode_o_r = self.info[:, L.RC_I_EPOR:L.RC_I_EPOR + P]
        self.episode_g = info_i[:, L.RC_I_EPG:L.RC_I_EPG + 2]
        E = self.info.shape[0]
        self.state_robots = self.info[:, L.RC_I_STATE_ROBOTS:L.RC_I_STATE_ROBOTS + L.RC_MAXP * 6].view(E, L.RC_MAXP, 6)[:, :P]
        self.state_ball = self.info[:, L.RC_I_STATE_BALL:L.RC_I_STATE_BALL + 3]
        recon = self.info[:, L.RC_I_RECON:L.RC_I_RECON + L.RC_MAXP * L.RC_FULL_WORDS].view(E, L.RC_MAXP, L.RC_FULL_WORDS)[:, :P]
        self.recon_ball = recon[..., L.RC_U_BALL:L.RC_U_BALL + 4]
        self.recon_self = recon[..., L.RC_U_SELF:L.RC_U_SELF + 8]
        self.recon_others = recon[..., L.RC_U_OTHERS:L.RC_U_OTHERS + (L.RC_MAXP - 1) * 6].unflatten(-1, (L.RC_MAXP - 1, 6))
        if self.obs is not None:  # PARTIAL, and IMAGE (the same detections, also emitted as raw records)
            obs_i = self.obs.view(torch.int32)
            for name, off, cap, feat in _TYPES:
                setattr(self, name, self.obs[..., off:off + cap * feat].unflatten(-1, (cap, feat)))
            self.counts = obs_i[..., L.RC_O_COUNTS:L.RC_O_COUNTS + L.RC_NTYPES]
            self.num_landmarks = obs_i[..., L.RC_O_NLANDMARKS]
            self.balls_seen = obs_i[..., L.RC_O_BALLSEEN]
            self.robots_seen = obs_i[..., L.RC_O_ROBSEEN:L.RC_O_ROBSEEN + L.RC_MAXP - 1]
        if self.full is not None:
            self.full_ball = self.full[..., L.RC_U_BALL:L.RC_U_BALL + 4]
            self.full_self = self.full[..., L.RC_U_SELF:L.RC_U_SELF + 8]
            self.full_others = self.full[..., L.RC_U_OTHERS:L.RC_U_OTHERS + (L.RC_MAXP - 1) * 6].unflatten(-1, (L.RC_MAXP - 1, 6))


class BatchedRoboCup(object):
    """nEnvs RoboCup environments on one GPU (reference: DynEnv/RoboCupEnvironment.py:23-64 per env,
    DynEnv/__init__.py:6-25 for the batch)."""

    nTimeSteps = L.RC_NTICK
    stepIterCnt = L.RC_SUBSTEPS

    def __init__(self, nEnvs, nPlayers, observationType=ObservationType.PARTIAL, noiseType=NoiseType.REALISTIC,
                 noiseMagnitude=2.0, allowHeadTurn=False, randomInit=False, deterministicTurn=False, canFall=True,
                 useObsRewards=True, auto_reset=False, seed=0, device=0, env_id_offset=0, terminal_outputs=False,
                 sparse_obs=False):
        """sparse_obs (DYNENV_FLAG_SPARSE_OBS): PARTIAL records are a counts-prefixed stream -- rows behind `counts` keep
        whatever the output buffer held (zeros after construction, stale rows later); read them through `counts`."""
        if not torch.cuda.is_available():
            raise RuntimeError("dynenv_b200 needs a CUDA device (no CPU fallback)")
        self._lib = _lib.lib()
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.nEnvs = int(nEnvs)
        self.sparse_obs = bool(sparse_obs)
        self.nPlayers = min(int(nPlayers), 5)
        self.P = 2 * self.nPlayers
        self.obs_type = int(observationType)
        cfg = _lib.DynEnvConfig(L.DYNENV_ABI_VERSION, L.DYNENV_KIND_ROBOCUP, self.nEnvs, int(nPlayers), self.obs_type,
                                int(noiseType), float(noiseMagnitude), int(bool(allowHeadTurn)),
                                _flags(randomInit, deterministicTurn, canFall, useObsRewards, auto_reset, sparse_obs),
                                int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_id_offset), self.device.index or 0)
        self._h = C.c_void_p()
        _lib.check(self._lib.dynenv_create(C.byref(cfg), C.byref(self._h)))
        self.state_words = self._lib.dynenv_state_words(self._h)
        with torch.cuda.device(self.device):
            self.out = RoboCupOutputs(self.nEnvs, self.P, self.obs_type, self.device)
            self.terminal = RoboCupOutputs(self.nEnvs, self.P, self.obs_type, self.device) if terminal_outputs else None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dynenv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_random_seed(self, seed):
        _lib.check(self._lib.dynenv_set_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def reset(self, env_mask=None):
        """Re-initialises all (or the masked) environments; returns the output views."""
        mask_ptr = None
        if env_mask is not None:
            env_mask = env_mask.to(device=self.device, dtype=torch.uint8).contiguous()
            assert env_mask.numel() == self.nEnvs
            mask_ptr = env_mask.data_ptr()
        _lib.check(self._lib.dynenv_reset(self._h, mask_ptr, C.byref(self.out.c_struct), self._stream()))
        return self.out

    def step(self, actions):
        """actions: float32 CUDA tensor [nEnvs, 2*nPlayers, 4] (move, turn, kick, head); a uint8 tensor of the same shape takes
        the byte entry point (dynenv_step_u8: MultiDiscrete action indices, a quarter of the bytes to bring to the device)."""
        if actions.shape != (self.nEnvs, self.P, 4):
            raise Exception("Error: There must be %d actions of length 4" % self.P)  # RoboCupEnvironment.py:462-463
        if actions.dtype == torch.uint8:
            if actions.device != self.device or not actions.is_contiguous():
                actions = actions.to(device=self.device).contiguous()
            term = C.byref(self.terminal.c_struct) if self.terminal is not None else None
            _lib.check(self._lib.dynenv_step_u8(self._h, actions.data_ptr(), C.byref(self.out.c_struct), term, self._stream()))
            return self.out
        if actions.dtype != torch.float32 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()  # (also a tensor of ANOTHER GPU)
        term = C.byref(self.terminal.c_struct) if self.terminal is not None else None
        _lib.check(self._lib.dynenv_step(self._h, actions.data_ptr(), C.byref(self.out.c_struct), term, self._stream()))
        return self.out

    def render_images(self, env_begin=0, env_end=None, out=None, outputs=None):
        """IMAGE observations of environments [env_begin, env_end): uint32 [n, T, P, 8, 480, 20] bit planes (bit x % 32
        of word x // 32 = pixel x; planes 0..3 top camera, 4..7 bottom camera) drawn from the raw detection records
        of the last reset() / step().  15.4 MB per environment: render the range you need into `out`."""
        o = outputs if outputs is not None else self.out
        if o.raw is None:
            raise RuntimeError("render_images needs observationType IMAGE")
        env_end = self.nEnvs if env_end is None else env_end
        n = env_end - env_begin
        raw = o.raw[env_begin:env_end]
        shape = (n, L.RC_NTICK, self.P, L.RC_IMG_PLANES, L.RC_IMG_H, L.RC_IMG_WORDS_PER_ROW)
        if out is None:
            out = torch.empty(shape, dtype=torch.int32, device=self.device)
        assert out.is_cuda and out.is_contiguous() and out.numel() == n * L.RC_NTICK * self.P * L.RC_IMG_WORDS and out.element_size() == 4
        _lib.check(self._lib.dynenv_render_images(self.device.index or 0, raw.data_ptr(), n * L.RC_NTICK * self.P, out.data_ptr(),
                                                  self._stream()))
        return out.view(shape)

    def debug_timing(self, enable=True, serialise=False):
        """serialise: the handed-over environments' kernels on the caller's stream instead of beside the observations, so that
        every kernel class has a duration of its own (slower than the production step)"""
        _lib.check(self._lib.dynenv_debug_timing(self._h, (2 if serialise else 1) if enable else 0))

    def last_kernel_ms(self):
        """{'phys': (ms, launches), 'obs': ..., 'finish': ..., 'phys_slow': ...} of the last step (needs debug_timing(True));
        'phys_slow' = the substeps of the environments handed over to the record-based code (RoboCup)"""
        ms, n = (C.c_float * 4)(), (C.c_int * 4)()
        _lib.check(self._lib.dynenv_last_kernel_ms(self._h, ms, n))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("phys", "obs", "finish", "phys_slow"))}

    def debug_substep_range(self, begin, end):
        _lib.check(self._lib.dynenv_debug_substep_range(self._h, int(begin), int(end)))

    def last_launch_count(self):
        return self._lib.dynenv_last_launch_count(self._h)

    def get_state(self):
        """[nEnvs, state_words] int32 tensor holding the raw 32-bit words of every state record."""
        st = torch.empty((self.nEnvs, self.state_words), dtype=torch.int32, device=self.device)
        _lib.check(self._lib.dynenv_get_state(self._h, st.data_ptr(), self._stream()))
        return st

    def set_state(self, st):
        st = st.to(device=self.device).contiguous()
        assert st.shape == (self.nEnvs, self.state_words) and st.element_size() == 4
        _lib.check(self._lib.dynenv_set_state(self._h, st.data_ptr(), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # `st` may be a temporary


# ---------------------------------------------------------------------------------------------- Driving
_DRV_TYPES = [("car", L.DRV_O_CAR, L.DRV_CAP_CAR, L.DRV_F_CAR), ("obs", L.DRV_O_OBS, L.DRV_CAP_OBS, L.DRV_F_OBS),
              ("ped", L.DRV_O_PED, L.DRV_CAP_PED, L.DRV_F_PED), ("lane", L.DRV_O_LANE, L.DRV_CAP_LANE, L.DRV_F_LANE_FULL)]


class DrivingOutputs(object):
    """Named views of one set of Driving output buffers (no copies): per (env, car) the padded-with-counts
    record of include/dynenv_b200.h (DRV_O_*), per env the result record (DRV_I_*)."""

    def __init__(self, nEnvs, N, device):
        self.N = N
        self.record = torch.zeros((nEnvs, N, L.DRV_OBS_WORDS), dtype=torch.float32, device=device)
        self.info = torch.zeros((nEnvs, L.DRV_INFO_WORDS), dtype=torch.float32, device=device)
        self.c_struct = _lib.DynEnvDrivingOut(self.record.data_ptr(), self.info.data_ptr())
        rec_i, info_i = self.record.view(torch.int32), self.info.view(torch.int32)
        self.self_ = self.record[..., L.DRV_O_SELF:L.DRV_O_SELF + L.DRV_F_SELF]
        for name, off, cap, feat in _DRV_TYPES:
            setattr(self, name, self.record[..., off:off + cap * feat].unflatten(-1, (cap, feat)))
        self.counts = rec_i[..., L.DRV_O_COUNTS:L.DRV_O_COUNTS + 4]
        self.rewards = self.info[:, L.DRV_I_REWARDS:L.DRV_I_REWARDS + N]
        self.done = info_i[:, L.DRV_I_DONE]
        self.diag = info_i[:, L.DRV_I_FLAGS]
        self.episode_r = self.info[:, L.DRV_I_EPR:L.DRV_I_EPR + N]
        self.episode_p_r = self.info[:, L.DRV_I_EPPR:L.DRV_I_EPPR + N]
        self.episode_g = info_i[:, L.DRV_I_EPG:L.DRV_I_EPG + 2]
        self.state_cars = self.info[:, L.DRV_I_STATE_CARS:L.DRV_I_STATE_CARS + L.DRV_MAXCARS * 7].unflatten(-1, (L.DRV_MAXCARS, 7))[:, :N]
        self.state_obs = self.info[:, L.DRV_I_STATE_OBS:L.DRV_I_STATE_OBS + L.DRV_MAXOBS * 4].unflatten(-1, (L.DRV_MAXOBS, 4))
        self.state_ped = self.info[:, L.DRV_I_STATE_PED:L.DRV_I_STATE_PED + L.DRV_MAXPED * 2].unflatten(-1, (L.DRV_MAXPED, 2))
        self.goals = self.info[:, L.DRV_I_GOALS:L.DRV_I_GOALS + L.DRV_MAXCARS * 2].unflatten(-1, (L.DRV_MAXCARS, 2))[:, :N]
        self.n_obs = info_i[:, L.DRV_I_NOBS]
        self.n_ped = info_i[:, L.DRV_I_NPED]


class BatchedDriving(object):
    """nEnvs Driving environments on one GPU (reference: DynEnv/DrivingEnvironment.py:20-56 per env,
    DynEnv/__init__.py:6-25 for the batch).  continuousActions=True takes acc = action[0], steer = action[1]
    (the reference's own continuous branch raises UnboundLocalError, DrivingEnvironment.py:360-368)."""

    nTimeSteps = 1
    stepIterCnt = L.DRV_SUBSTEPS

    def __init__(self, nEnvs, nPlayers, observationType=ObservationType.PARTIAL, noiseType=NoiseType.REALISTIC,
                 noiseMagnitude=2.0, continuousActions=False, auto_reset=False, seed=0, device=0, env_id_offset=0,
                 terminal_outputs=False):
        if not torch.cuda.is_available():
            raise RuntimeError("dynenv_b200 needs a CUDA device (no CPU fallback)")
        self._lib = _lib.lib()
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.nEnvs = int(nEnvs)
        self.nPlayers = self.N = min(int(nPlayers), L.DRV_MAXCARS)
        self.obs_type = int(observationType)
        cfg = _lib.DynEnvConfig(L.DYNENV_ABI_VERSION, L.DYNENV_KIND_DRIVING, self.nEnvs, int(nPlayers), self.obs_type,
                                int(noiseType), float(noiseMagnitude), int(bool(continuousActions)),
                                L.DYNENV_FLAG_AUTO_RESET if auto_reset else 0,
                                int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_id_offset), self.device.index or 0)
        self._h = C.c_void_p()
        _lib.check(self._lib.dynenv_create(C.byref(cfg), C.byref(self._h)))
        self.state_words = self._lib.dynenv_state_words(self._h)
        with torch.cuda.device(self.device):
            self.out = DrivingOutputs(self.nEnvs, self.N, self.device)
            self.terminal = DrivingOutputs(self.nEnvs, self.N, self.device) if terminal_outputs else None

    close = BatchedRoboCup.close
    __del__ = BatchedRoboCup.__del__
    _stream = BatchedRoboCup._stream
    set_random_seed = BatchedRoboCup.set_random_seed
    reset = BatchedRoboCup.reset
    debug_timing = BatchedRoboCup.debug_timing
    last_kernel_ms = BatchedRoboCup.last_kernel_ms
    debug_substep_range = BatchedRoboCup.debug_substep_range
    last_launch_count = BatchedRoboCup.last_launch_count
    get_state = BatchedRoboCup.get_state
    set_state = BatchedRoboCup.set_state

    def step(self, actions):
        """actions: float32 CUDA tensor [nEnvs, nCars, 2] (acceleration, steering); uint8 (discrete action indices): dynenv_step_u8."""
        if actions.shape != (self.nEnvs, self.N, 2):
            raise Exception("Error: There must be 2 actions for every car")  # DrivingEnvironment.py:262-263
        if actions.dtype == torch.uint8:
            if actions.device != self.device or not actions.is_contiguous():
                actions = actions.to(device=self.device).contiguous()
            term = C.byref(self.terminal.c_struct) if self.terminal is not None else None
            _lib.check(self._lib.dynenv_step_u8(self._h, actions.data_ptr(), C.byref(self.out.c_struct), term, self._stream()))
            return self.out
        if actions.dtype != torch.float32 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()  # (also a tensor of ANOTHER GPU)
        term = C.byref(self.terminal.c_struct) if self.terminal is not None else None
        _lib.check(self._lib.dynenv_step(self._h, actions.data_ptr(), C.byref(self.out.c_struct), term, self._stream()))
        return self.out
