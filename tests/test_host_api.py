"""CPU: host-side logic of the drop-in API (no CUDA calls): action validation with the reference's
messages, spaces, the ragged reference format built from padded records, the C-ABI exports."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dynenv_b200 import api, layout as L, reference_format as RF
from dynenv_b200.enums import DynEnvType, NoiseType, ObservationType

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_enums_match_reference_values():
    assert (int(ObservationType.FULL), int(ObservationType.PARTIAL), int(ObservationType.IMAGE)) == (0, 1, 2)
    assert (int(NoiseType.RANDOM), int(NoiseType.REALISTIC)) == (0, 1)
    assert (int(DynEnvType.ROBO_CUP), int(DynEnvType.DRIVE)) == (0, 1)
    assert str(ObservationType.PARTIAL) == "partial" and ObservationType.argparse("image") is ObservationType.IMAGE


def test_action_validation_messages():
    ok = np.array([[4, 2, 2, 6]] * 4)
    assert api.validate_robocup_actions(ok, 4, False).dtype == np.float32
    with pytest.raises(Exception, match="There must be 4 actions of length 4"):
        api.validate_robocup_actions(np.zeros((3, 4)), 4, False)
    with pytest.raises(Exception, match="Robot motion commands"):
        api.validate_robocup_actions(np.array([[5, 0, 0, 0]] * 2), 2, False)
    with pytest.raises(Exception, match="between 0 and 6"):
        api.validate_robocup_actions(np.array([[0, 0, 0, 7]] * 2), 2, False)
    with pytest.raises(Exception, match="between -3 and \\+3"):
        api.validate_robocup_actions(np.array([[0, 0, 0, 3.5]] * 2), 2, True)
    api.validate_robocup_actions(np.array([[0, 0, 0, -2.5]] * 2), 2, True)


def test_spaces():
    act, obs = api.robocup_spaces(ObservationType.PARTIAL, False)
    assert list(act.spaces[0].nvec) == [5, 3, 3, 7]
    assert len(obs.spaces) == 2 and len(obs.spaces[1].spaces) == 4
    act, _ = api.robocup_spaces(ObservationType.FULL, True)
    assert list(act.spaces[0].nvec) == [5, 3, 3] and act.spaces[1].shape == (1,)


def _fake_views(E, P, rng):
    T = L.RC_NTICK
    v = {}
    caps = dict(ball=(L.RC_CAP_BALL, L.RC_F_BALL), robot=(L.RC_CAP_ROBOT, L.RC_F_ROBOT), post=(L.RC_CAP_POST, L.RC_F_POST),
                cross=(L.RC_CAP_CROSS, L.RC_F_CROSS), fcross=(L.RC_CAP_FCROSS, L.RC_F_FCROSS), line=(L.RC_CAP_LINE, L.RC_F_LINE))
    v["counts"] = np.zeros((E, T, P, 6), np.int32)
    for i, (k, (c, f)) in enumerate(caps.items()):
        v[k] = rng.normal(size=(E, T, P, c, f)).astype(np.float32)
        v["counts"][..., i] = rng.integers(0, c + 1, (E, T, P))
    v["num_landmarks"] = rng.integers(0, 20, (E, T, P)).astype(np.int32)
    v["robots_seen"] = rng.integers(0, 2, (E, T, P, L.RC_MAXP - 1)).astype(np.int32)
    v["balls_seen"] = rng.integers(0, 2, (E, T, P)).astype(np.int32)
    v["rewards"] = rng.normal(size=(E, P)).astype(np.float32)
    v["done"] = np.array([0, 1][:E], np.int32)
    v["state_robots"] = rng.normal(size=(E, P, 6)).astype(np.float32)
    v["state_ball"] = rng.normal(size=(E, 3)).astype(np.float32)
    v["recon_ball"] = rng.normal(size=(E, P, 4)).astype(np.float32)
    v["recon_self"] = rng.normal(size=(E, P, 8)).astype(np.float32)
    v["recon_others"] = rng.normal(size=(E, P, L.RC_MAXP - 1, 6)).astype(np.float32)
    for k in ("episode_r", "episode_p_r", "episode_o_r"):
        v[k] = rng.normal(size=(E, P)).astype(np.float32)
    v["episode_g"] = np.array([[0, 0], [2, 1]][:E], np.int32)
    return v


def test_reference_format_partial():
    P, E = 6, 2
    v = _fake_views(E, P, np.random.default_rng(0))
    obs = RF.observations(v, 1, P, full=False)
    assert len(obs) == L.RC_NTICK and len(obs[0]) == P
    (balls, robots), (posts, crosses, fcs, lines), (nl, rs, bs) = obs[2][3]
    c = v["counts"][1, 2, 3]
    for arr, n, f in ((balls, c[0], 5), (robots, c[1], 7), (posts, c[2], 6), (crosses, c[3], 6), (fcs, c[4], 8), (lines, c[5], 5)):
        assert arr.dtype == np.float32
        assert arr.shape == ((n, f) if n else (0,))  # empty lists are np.array([]) in the reference
    assert np.array_equal(robots, v["robot"][1, 2, 3, :c[1]]) if c[1] else True
    assert rs.dtype == np.uint8 and rs.shape == (P - 1,) and isinstance(bs, bool) and nl == v["num_landmarks"][1, 2, 3]
    info = RF.info_dict(v, 1, P)
    assert info["Full State"][0].shape == (P, 6) and len(info["Recon States"]) == P
    assert info["Recon States"][0][2].shape == (P - 1, 6) and info["episode_g"] == [2, 1]
    assert "episode_r" not in RF.info_dict(v, 0, P)
    stacked = api.BatchedVecEnv._stack([RF.observations(v, e, P, False) for e in range(E)])
    assert stacked.shape == (E, L.RC_NTICK, P, 3) and stacked.dtype == object


def test_library_exports_every_declared_symbol():
    """the C-ABI library loads and exports exactly what include/dynenv_b200.h declares"""
    from dynenv_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dynenv_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dynenv_\w+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dynenv_abi_version() == L.DYNENV_ABI_VERSION


def test_no_device_fails_loudly():
    """without a CUDA device dynenv_create returns DYNENV_E_NODEVICE: there is no CPU fallback"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from dynenv_b200 import _lib
    cfg = _lib.DynEnvConfig(L.DYNENV_ABI_VERSION, 0, 4, 5, 1, 1, 2.0, 0, 12, 0, 0, 0)
    h = C.c_void_p()
    assert _lib.lib().dynenv_create(C.byref(cfg), C.byref(h)) == L.DYNENV_E_NODEVICE
    assert b"no CPU fallback" in _lib.lib().dynenv_last_error()
    with pytest.raises(RuntimeError):
        import dynenv_b200
        dynenv_b200.BatchedRoboCup(2, 5)


def test_driving_validation_and_spaces():
    ok = np.array([[2, 0], [1, 1], [0, 2]])
    assert api.validate_driving_actions(ok, 3, False).dtype == np.float32
    with pytest.raises(Exception, match="2 actions for every car"):
        api.validate_driving_actions(np.zeros((2, 2)), 3, False)
    with pytest.raises(Exception, match="Acceleration must be between"):
        api.validate_driving_actions(np.array([[5, 1]] * 3), 3, False)
    with pytest.raises(Exception, match="Steering must be between"):
        api.validate_driving_actions(np.array([[0.5, -3.5]] * 3), 3, True)
    act, obs = api.driving_spaces(ObservationType.PARTIAL, False)
    assert list(act.spaces[0].nvec) == [3, 3]
    assert set(obs.spaces[1].spaces[1].spaces) == {"signed_distance", "orientation", "type"}
    act, obs = api.driving_spaces(ObservationType.FULL, True)
    assert act.spaces[0].shape == (2,) and set(obs.spaces[0].spaces[1].spaces) == {"position", "width_height"}


def test_driving_reference_format():
    """ragged Driving observations and info from padded records (host logic only)"""
    rng = np.random.default_rng(1)
    E, N = 2, 4
    v = {"self_": rng.normal(size=(E, N, L.DRV_F_SELF)).astype(np.float32),
         "car": rng.normal(size=(E, N, L.DRV_CAP_CAR, L.DRV_F_CAR)).astype(np.float32),
         "obs": rng.normal(size=(E, N, L.DRV_CAP_OBS, L.DRV_F_OBS)).astype(np.float32),
         "ped": rng.normal(size=(E, N, L.DRV_CAP_PED, L.DRV_F_PED)).astype(np.float32),
         "lane": rng.normal(size=(E, N, L.DRV_CAP_LANE, L.DRV_F_LANE_FULL)).astype(np.float32),
         "counts": rng.integers(0, 5, (E, N, 4)).astype(np.int32), "rewards": rng.normal(size=(E, N)).astype(np.float32),
         "done": np.array([0, 1], np.int32), "state_cars": rng.normal(size=(E, N, 7)).astype(np.float32),
         "state_obs": rng.normal(size=(E, L.DRV_MAXOBS, 4)).astype(np.float32),
         "state_ped": rng.normal(size=(E, L.DRV_MAXPED, 2)).astype(np.float32), "n_obs": np.array([11, 0], np.int32),
         "n_ped": np.array([13, 20], np.int32), "goals": rng.normal(size=(E, N, 2)).astype(np.float32),
         "episode_r": rng.normal(size=(E, N)).astype(np.float32), "episode_p_r": rng.normal(size=(E, N)).astype(np.float32),
         "episode_g": np.array([[0, 0], [1, 3]], np.int32)}
    v["counts"][0, 1] = [3, 0, 2, 6]
    obs = RF.drv_observations(v, 0, N, full=False)
    assert len(obs) == 1 and len(obs[0]) == N
    (cars, obst, peds), (self_, lanes), ones = obs[0][1]
    assert cars.shape == (3, 7) and obst.shape == (0,) and peds.shape == (2, 2) and lanes.shape == (6, 4) and self_.shape == (1, 9)
    assert np.array_equal(lanes, v["lane"][0, 1, :6, :4]) and ones == (1, 1, 1)
    info = RF.drv_info_dict(v, 1, N)
    assert info["Full State"][0].shape == (N, 7) and info["Full State"][1].shape == (0,) and info["Full State"][2].shape == (20, 2)
    assert info["Full State"][3].shape == (8, 5) and info["episode_g"] == [1, 3] and info["episode_o_r"] == [0] * N
    rs = info["Recon States"][2]
    assert rs[0].shape == (1, 9) and rs[1].shape == (N - 1, 7) and np.array_equal(rs[0][0, 6:8], v["goals"][1, 2])
    assert "episode_r" not in RF.drv_info_dict(v, 0, N)
    lanes = RF.drv_lanes_full()  # the reference's negative-index quirk: type tags of row i, geometry of another lane line
    assert lanes[:, 4].tolist() == [1, 0, -1, 0, 1, 1, -1, 1]
