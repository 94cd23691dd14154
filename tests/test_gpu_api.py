"""GPU: the drop-in classes (formats, auto-reset, errors) and shard independence."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_single_env_drop_in_full():
    import dynenv_b200
    from dynenv_b200.api import RoboCupEnvironment
    RoboCupEnvironment.canFall = False
    try:
        env = RoboCupEnvironment(nPlayers=1, observationType=dynenv_b200.ObservationType.FULL, noiseMagnitude=0)
    finally:
        RoboCupEnvironment.canFall = True
    obs = env.reset()
    assert len(obs) == 5 and len(obs[0]) == 2
    (ball, others), (self_,), ones = obs[0][0]
    assert ball.shape == (1, 4) and others.shape == (1, 6) and self_.shape == (1, 8) and ones == (1, 1, 1)
    assert ball.dtype == np.float32
    obs, rew, fin, info = env.step(np.array([[1, 0, 0, 3], [2, 1, 0, 3]]))
    assert rew.shape == (2,) and rew.dtype == np.float64 and fin is False
    assert info["Full State"][0].shape == (2, 6) and info["Full State"][1].shape == (3,)
    assert len(info["Recon States"]) == 2 and env.elapsed == 50
    locs = env.get_agent_locs()
    assert len(locs) == 2 and locs[0].shape == (1, 6)
    with pytest.raises(Exception, match="There must be 2 actions"):
        env.step(np.zeros((3, 4)))
    with pytest.raises(Exception, match="Robot motion commands"):
        env.step(np.array([[7, 0, 0, 3], [0, 0, 0, 3]]))
    env.setRandomSeed(5)
    env.close()


def test_single_env_partial_format():
    import dynenv_b200
    env = dynenv_b200.RoboCupEnvironment(nPlayers=2, noiseMagnitude=2)
    obs = env.reset()
    (balls, robots), (posts, crosses, fcs, lines), (nl, rs, bs) = obs[4][1]
    for arr, f in ((balls, 5), (robots, 7), (posts, 6), (crosses, 6), (fcs, 8), (lines, 5)):
        assert arr.dtype == np.float32 and (arr.shape == (0,) or arr.shape[1] == f)
    assert rs.shape == (3,) and rs.dtype == np.uint8 and isinstance(bs, bool) and isinstance(nl, int)
    obs, rew, fin, info = env.step(np.array([[3, 0, 0, 3]] * 4))
    assert len(obs) == 5 and len(obs[0]) == 4 and rew.shape == (4,)


def test_vec_env_auto_reset_and_terminal_observation():
    import torch
    import dynenv_b200
    from dynenv_b200 import layout as L
    venv, name = dynenv_b200.make_dyn_env(dynenv_b200.DynEnvType.ROBO_CUP, 3, 2, False, dynenv_b200.ObservationType.PARTIAL,
                                          dynenv_b200.NoiseType.REALISTIC, 1.0, False)
    assert name == "RoboCup" and venv.num_envs == 3
    obs = venv.reset()
    assert obs.shape == (3, 5, 4, 3)
    # move environment 1 to the last step of its episode
    st = venv.sim.get_state()
    st[1, L.RC_S_ELAPSED] = 12000 - 50
    venv.sim.set_state(st)
    a = np.tile(np.array([1, 0, 0, 3]), (3, 4, 1))
    obs, rew, dones, infos = venv.step(a)
    assert obs.shape == (3, 5, 4, 3) and rew.shape == (3, 4) and dones.tolist() == [False, True, False]
    assert "terminal_observation" in infos[1] and "terminal_observation" not in infos[0]
    assert "episode_r" in infos[1] and len(infos[1]["episode_g"]) == 2
    st = venv.sim.get_state().cpu().numpy()
    assert st[1, L.RC_S_ELAPSED] == 0 and st[1, L.RC_S_EPISODE] == 2 and st[0, L.RC_S_ELAPSED] == 50
    assert venv.get_attr("nTimeSteps") == [5, 5, 5]
    assert len(venv.env_method("get_agent_locs")) == 3
    with pytest.raises(Exception, match="There must be 4 actions"):
        venv.step(np.zeros((3, 5, 4)))
    venv.close()


def test_results_do_not_depend_on_sharding():
    """global env ids feed the RNG: 8 environments on one handle == 2 handles x 4 (what 2 GPUs would run)"""
    import torch
    import dynenv_b200
    kw = dict(nPlayers=3, noiseMagnitude=2.0, canFall=True, seed=11)
    whole = dynenv_b200.BatchedRoboCup(8, env_id_offset=0, **kw)
    parts = [dynenv_b200.BatchedRoboCup(4, env_id_offset=o, **kw) for o in (0, 4)]
    whole.reset()
    for p in parts:
        p.reset()
    g = torch.Generator().manual_seed(3)
    for _ in range(6):
        a = (torch.rand((8, 6, 4), generator=g) * torch.tensor([5, 3, 3, 7])).floor().cuda()
        ow = whole.step(a)
        op = [p.step(a[4 * k:4 * k + 4].contiguous()) for k, p in enumerate(parts)]
        assert torch.equal(ow.obs, torch.cat([o.obs for o in op]))
        assert torch.equal(ow.info, torch.cat([o.info for o in op]))
    assert torch.equal(whole.get_state(), torch.cat([p.get_state() for p in parts]))


def test_driving_single_env_drop_in():
    import dynenv_b200
    env = dynenv_b200.DrivingEnvironment(nPlayers=4, noiseMagnitude=2)
    obs = env.reset()
    assert len(obs) == 1 and len(obs[0]) == 4
    (cars, obst, peds), (self_, lanes), ones = obs[0][0]
    assert self_.shape == (1, 9) and self_.dtype == np.float32 and ones == (1, 1, 1)
    for arr, f in ((cars, 7), (obst, 6), (peds, 2), (lanes, 4)):
        assert arr.dtype == np.float32 and (arr.shape == (0,) or arr.shape[1] == f)
    obs, rew, fin, info = env.step(np.array([[2, 1], [2, 1], [0, 2], [1, 0]]))
    assert rew.shape == (4,) and rew.dtype == np.float64 and fin is False and env.elapsed == 10
    fs = info["Full State"]
    assert fs[0].shape == (4, 7) and fs[1].shape[1] == 4 and fs[2].shape[1] == 2 and fs[3].shape == (8, 5)
    rs = info["Recon States"]
    assert len(rs) == 4 and rs[0][0].shape == (1, 9) and rs[0][1].shape == (3, 7)
    assert np.allclose(rs[1][0][0, :6], fs[0][1, :6]) and rs[1][0][0, 8] == fs[0][1, 6]
    assert env.get_agent_locs()[0].shape == (1, 4)
    with pytest.raises(Exception, match="2 actions for every car"):
        env.step(np.zeros((3, 2)))
    with pytest.raises(Exception, match="Acceleration must be between"):
        env.step(np.array([[9, 1]] * 4))
    env.close()
    # FULL observations: ([others, obstacles, pedestrians], [self, lanes], (1, 1, 1))
    env = dynenv_b200.DrivingEnvironment(nPlayers=3, observationType=dynenv_b200.ObservationType.FULL, noiseMagnitude=0)
    (others, obst, peds), (self_, lanes), _ = env.reset()[0][2]
    assert others.shape == (2, 7) and obst.shape[1] == 4 and lanes.shape == (8, 5) and self_.shape == (1, 9)
    assert np.allclose(lanes, env.getFullState()[3])
    env.close()


def test_driving_vec_env_auto_reset():
    import dynenv_b200
    from dynenv_b200 import layout as L
    venv, name = dynenv_b200.make_dyn_env(dynenv_b200.DynEnvType.DRIVE, 3, 5, False, dynenv_b200.ObservationType.PARTIAL,
                                          dynenv_b200.NoiseType.REALISTIC, 1.0, True)
    assert name == "Driving" and venv.num_envs == 3
    obs = venv.reset()
    assert obs.shape == (3, 1, 5, 3)
    st = venv.sim.get_state()
    st[2, L.DRV_S_ELAPSED] = 6000 - 10
    venv.sim.set_state(st)
    a = np.random.default_rng(0).uniform(-3, 3, (3, 5, 2))
    obs, rew, dones, infos = venv.step(a)
    assert obs.shape == (3, 1, 5, 3) and rew.shape == (3, 5) and dones.tolist() == [False, False, True]
    assert "terminal_observation" in infos[2] and "episode_r" in infos[2] and infos[2]["episode_o_r"] == [0] * 5
    st = venv.sim.get_state().cpu().numpy()
    assert st[2, L.DRV_S_ELAPSED] == 0 and st[2, L.DRV_S_EPISODE] == 2 and st[0, L.DRV_S_ELAPSED] == 10
    assert venv.get_attr("nTimeSteps") == [1, 1, 1] and len(venv.env_method("get_agent_locs")) == 3
    venv.close()


def test_driving_results_do_not_depend_on_sharding():
    import torch
    import dynenv_b200
    kw = dict(nPlayers=6, noiseMagnitude=2.0, continuousActions=True, seed=13)
    whole = dynenv_b200.BatchedDriving(8, env_id_offset=0, **kw)
    parts = [dynenv_b200.BatchedDriving(4, env_id_offset=o, **kw) for o in (0, 4)]
    whole.reset()
    for p in parts:
        p.reset()
    g = torch.Generator().manual_seed(4)
    for _ in range(8):
        a = (torch.rand((8, 6, 2), generator=g) * 6 - 3).cuda()
        ow = whole.step(a)
        op = [p.step(a[4 * k:4 * k + 4].contiguous()) for k, p in enumerate(parts)]
        assert torch.equal(ow.record, torch.cat([o.record for o in op]))
        assert torch.equal(ow.info, torch.cat([o.info for o in op]))
    assert torch.equal(whole.get_state(), torch.cat([p.get_state() for p in parts]))


def test_device_arranger_on_simulator_outputs():
    """the InOutArranger equivalent on the device tensors of a step == the same arrays built from the reference-format
    nested lists (what the reference's model would do in Python)"""
    import torch
    import dynenv_b200
    from dynenv_b200 import reference_format as RF
    from dynenv_b200.arranger import DeviceArranger, robocup_rows
    sim = dynenv_b200.BatchedRoboCup(5, 3, noiseMagnitude=2.0, seed=3)
    sim.reset()
    out = sim.step((torch.rand((5, 6, 4)) * torch.tensor([5, 3, 3, 7])).floor().cuda())
    arr = DeviceArranger(*robocup_rows(out))
    inputs, (counts, maxCount, objCounts) = arr.rearrange_inputs()
    v = RF.host_views(out, False)
    obs = [RF.observations(v, e, 6, False) for e in range(5)]
    for i in range(6):  # flat per-type arrays in (time, env x player, row) order
        ref = [obs[e][t][a][i // 2 if i < 2 else 1][i if i < 2 else i - 2] for t in range(5) for e in range(5) for a in range(6)]
        ref = [r for r in ref if r.ndim == 2]
        ref = np.concatenate(ref) if ref else np.zeros((0, inputs[i].shape[1]), np.float32)
        assert np.array_equal(inputs[i].cpu().numpy(), ref), i
    padded, masks = arr.rearrange_outputs([x[:, :2] for x in inputs])
    assert padded.shape == (5, maxCount, 30, 2) and len(masks) == 5 and masks[0].shape == (30, maxCount)
    assert int((~masks[2]).sum()) == int(objCounts[2].sum())


def test_fp32_peak_probe():
    """the FMA probe behind roofline.fp32.measured_peak_tflops: a plausible FP32 rate for a B200 (nominal 74.4 TFLOP/s)"""
    import ctypes as C
    from dynenv_b200 import _lib
    tf = C.c_float(0)
    assert _lib.lib().dynenv_debug_fp32_peak(0, C.byref(tf)) == 0
    assert 20.0 < tf.value < 90.0, tf.value
    assert _lib.lib().dynenv_debug_fp32_peak(0, None) != 0


def test_bad_actions_are_flagged_on_the_device():
    """the batched device path cannot raise like processAction (RoboCupEnvironment.py:543-550, DrivingEnvironment.py:364-367):
    a step with an out-of-range or non-finite action carries RC_IF_BAD_ACTION in its diag word, the others do not"""
    import torch
    import dynenv_b200
    from dynenv_b200 import layout as L
    sim = dynenv_b200.BatchedRoboCup(4, 2, 1, 1, 2.0, seed=1)
    sim.reset()
    a = torch.tensor([1., 0., 0., 3.]).repeat(4, 4, 1).cuda()
    assert int(sim.step(a).diag.max()) == 0
    a[1, 2, 0] = 7.          # move 7 (the kernels would walk backwards)
    a[2, 0, 3] = float("nan")  # head turn NaN
    a[3, 1, 1] = -1.         # turn -1
    diag = sim.step(a).diag.cpu().numpy()
    assert diag[0] == 0 and all(diag[1:] & L.RC_IF_BAD_ACTION)
    assert int(sim.step(torch.tensor([1., 0., 0., 3.]).repeat(4, 4, 1).cuda()).diag.max()) == 0  # a flag of ONE step
    # allowHeadTurn: the head command is a float in [-3, 3]
    sim = dynenv_b200.BatchedRoboCup(2, 1, 1, 1, 2.0, allowHeadTurn=True, seed=1)
    sim.reset()
    a = torch.tensor([1., 0., 0., -2.5]).repeat(2, 2, 1).cuda()
    assert int(sim.step(a).diag.max()) == 0
    a[1, 0, 3] = 3.5
    diag = sim.step(a).diag.cpu().numpy()
    assert diag[0] == 0 and diag[1] & L.RC_IF_BAD_ACTION
    drv = dynenv_b200.BatchedDriving(3, 4, 1, 1, 2.0, continuousActions=True, seed=1)
    drv.reset()
    a = torch.zeros((3, 4, 2)).cuda()
    assert int(drv.step(a).diag.max()) == 0
    a[1, 3, 0] = 3.5
    a[2, 0, 1] = float("inf")
    diag = drv.step(a).diag.cpu().numpy()
    assert diag[0] == 0 and all(diag[1:] & L.RC_IF_BAD_ACTION)


def test_reco_descriptor_and_render_pass_through():
    """recoDescriptor (environment_base.py:18-38, RoboCupEnvironment.py:101-132, DrivingEnvironment.py:129-168) on the drop-in
    classes and through get_attr (examples/main.py:42); make_dyn_env(render=True) keeps frames instead of raising"""
    import dynenv_b200
    venv, _ = dynenv_b200.make_dyn_env(dynenv_b200.DynEnvType.ROBO_CUP, 2, 2, True, dynenv_b200.ObservationType.PARTIAL,
                                       dynenv_b200.NoiseType.REALISTIC, 1.0, False)
    rd = venv.get_attr("recoDescriptor", 0)[0]
    assert rd.featureGridSize == (1, 1) and len(rd.fullStateSpace) == 2 and len(rd.targetDefs) == 2
    assert rd.fullStateSpace[1].numItemsPerGridCell == 4 and rd.targetDefs[1].contIdx == [2, 3, 4] and rd.targetDefs[1].binaryIdx == [5]
    assert rd.targetDefs[0].numContinuous == 1 and tuple(rd.targetDefs[0].posIdx) == (0, 1)
    venv.reset()
    venv.step(np.tile(np.array([1, 0, 0, 3]), (2, 4, 1)))
    assert venv.get_attr("renderMode")[0] == "memory" and len(venv.screenShots) == 2
    assert venv.screenShots[0].ndim == 3 and venv.screenShots[0].shape[2] == 3
    venv.close()
    denv, name = dynenv_b200.make_dyn_env(dynenv_b200.DynEnvType.DRIVE, 2, 3, False, dynenv_b200.ObservationType.PARTIAL,
                                          dynenv_b200.NoiseType.REALISTIC, 1.0, True)
    rd = denv.get_attr("recoDescriptor", [0, 1])
    assert len(rd) == 2 and rd[0].featureGridSize == (10, 17) and [s.numItemsPerGridCell for s in rd[0].fullStateSpace] == [1, 4, 4, 6]
    assert [t.numContinuous for t in rd[0].targetDefs] == [4, 4, 2, 0]
    denv.close()
    env = dynenv_b200.RoboCupEnvironment(nPlayers=1)
    assert env.recoDescriptor.featureGridSize == (1, 1)
    env.close()


def test_entry_points_leave_the_current_device_alone():
    """a handle on another GPU must not switch the calling thread's current device (needs two GPUs)"""
    import torch
    import dynenv_b200
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU")
    torch.cuda.set_device(0)
    sim = dynenv_b200.BatchedRoboCup(4, 1, 1, 1, 2.0, seed=1, device=1)
    assert torch.cuda.current_device() == 0
    sim.reset()
    out = sim.step(torch.tensor([1., 0., 0., 3.]).repeat(4, 2, 1))  # a host tensor: moved to the simulator's device
    assert torch.cuda.current_device() == 0 and out.rewards.device.index == 1
    sim.get_state()
    sim.close()
    assert torch.cuda.current_device() == 0


def test_kernel_arranger_matches_the_reference_class():
    """dynenv_arrange_inputs / dynenv_arrange_outputs (the InOutArranger kernels) on observation records packed from the golden
    rows, against the outputs of the reference's OWN InOutArranger (tests/golden/arranger.npz, gen_golden_arranger.py)"""
    import os
    import torch
    from dynenv_b200 import layout as L
    from dynenv_b200.arranger import KernelArranger
    from helpers import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "arranger.npz")))
    E, T, P = g["counts"].shape[:3]
    rec = np.zeros((E, T, P, L.RC_OBS_WORDS), np.float32)
    offs = [L.RC_O_BALL, L.RC_O_ROBOT, L.RC_O_POST, L.RC_O_CROSS, L.RC_O_FCROSS, L.RC_O_LINE]
    for k in range(6):
        r = g["rows%d" % k]
        rec[..., offs[k]:offs[k] + r.shape[3] * r.shape[4]] = r.reshape(E, T, P, -1)
    rec.view(np.int32)[..., L.RC_O_COUNTS:L.RC_O_COUNTS + 6] = g["counts"]
    arr = KernelArranger(torch.from_numpy(rec).cuda(), KernelArranger.robocup_layout())
    inputs, (counts, maxCount, objCounts) = arr.rearrange_inputs()
    assert int(maxCount) == int(g["maxCount"])
    assert np.array_equal(counts.cpu().numpy(), g["ref_counts"]) and np.array_equal(objCounts.cpu().numpy(), g["objCounts"])
    exact = arr.compact(inputs)
    for k in range(6):
        assert np.array_equal(exact[k].cpu().numpy(), g["inputs%d" % k]), k
    padded, masks = arr.rearrange_outputs([x[:, :3].contiguous() for x in inputs], max_count=int(g["maxCount"]))
    assert np.array_equal(padded.cpu().numpy(), g["padded"]) and np.array_equal(masks.cpu().numpy(), g["masks"])
    # a static bound larger than maxCount only adds padding
    padded2, masks2 = arr.rearrange_outputs([x[:, :3].contiguous() for x in inputs], max_count=int(g["maxCount"]) + 5)
    assert np.array_equal(padded2.cpu().numpy()[:, :int(g["maxCount"])], g["padded"]) and float(padded2[:, int(g["maxCount"]):].abs().sum()) == 0
    assert bool(masks2[..., int(g["maxCount"]):].all())
    # a type without an embedding block takes no slots
    padded3, masks3 = arr.rearrange_outputs([None] + [x[:, :3].contiguous() for x in inputs[1:]], max_count=20)
    ref = g["objCounts"] - g["ref_counts"][0]
    assert np.array_equal((~masks3.cpu().numpy()).sum(-1), np.minimum(ref, 20))


def test_kernel_arranger_on_simulator_outputs():
    """the kernels on a step's records == the torch formulation on the same step (RoboCup and Driving), at a size with
    several scan blocks"""
    import torch
    import dynenv_b200
    from dynenv_b200.arranger import DeviceArranger, KernelArranger, driving_rows, robocup_rows
    sim = dynenv_b200.BatchedRoboCup(300, 5, noiseMagnitude=2.0, seed=3)
    sim.reset()
    out = sim.step((torch.rand((300, 10, 4)) * torch.tensor([5, 3, 3, 7])).floor().cuda())
    karr = KernelArranger.for_robocup(sim)
    kin, (kc, kmax, ko) = karr.rearrange_inputs()
    tin, (tc, tmax, to) = DeviceArranger(*robocup_rows(out)).rearrange_inputs()
    assert int(kmax) == tmax and torch.equal(kc.long(), tc) and torch.equal(ko.long(), to)
    for a, b in zip(karr.compact(kin), tin):
        assert torch.equal(a, b)
    kp, km = karr.rearrange_outputs([x[:, :2].contiguous() for x in kin], max_count=tmax)
    tp, tm = DeviceArranger(*robocup_rows(out)).rearrange_outputs([x[:, :2] for x in tin])
    assert torch.equal(kp, tp) and torch.equal(km, torch.stack(tm))
    drv = dynenv_b200.BatchedDriving(200, 10, noiseMagnitude=2.0, continuousActions=True, seed=3)
    drv.reset()
    dout = drv.step((torch.rand((200, 10, 2)) * 6 - 3).cuda())
    karr = KernelArranger.for_driving(drv)
    kin, (kc, kmax, ko) = karr.rearrange_inputs()
    tin, (tc, tmax, to) = DeviceArranger(*driving_rows(dout)).rearrange_inputs()
    assert int(kmax) == tmax and torch.equal(kc.long(), tc)
    for a, b in zip(karr.compact(kin), tin):
        assert torch.equal(a, b)


def test_sparse_observation_stream_matches_the_padded_records():
    """DYNENV_FLAG_SPARSE_OBS: the same simulation with rows behind the counts left unwritten.  The output buffer is
    poisoned with NaN first: counts / flags / rewards / states are identical, every row in front of its count is
    bit-identical to the zero-filled run, nothing behind the counts was written (still NaN), and the arranger kernels
    -- a consumer that honours the counts -- produce identical inputs from both."""
    import torch
    import dynenv_b200
    from dynenv_b200 import layout as L
    from dynenv_b200.arranger import KernelArranger
    n, P = 96, 10
    kw = dict(noiseMagnitude=2.0, canFall=True, auto_reset=True, seed=11)
    dense = dynenv_b200.BatchedRoboCup(n, 5, **kw)
    sparse = dynenv_b200.BatchedRoboCup(n, 5, sparse_obs=True, **kw)
    g = torch.Generator().manual_seed(5)
    lists = [("ball", L.RC_CAP_BALL), ("robot", L.RC_CAP_ROBOT), ("post", L.RC_CAP_POST), ("cross", L.RC_CAP_CROSS),
             ("fcross", L.RC_CAP_FCROSS), ("line", L.RC_CAP_LINE)]
    written_bytes, steps = 0, 0
    for s in range(8):
        sparse.out.obs.fill_(float("nan"))
        if s == 0:
            do, so = dense.reset(), sparse.reset()
        else:
            a = (torch.rand((n, P, 4), generator=g) * torch.tensor([5, 3, 3, 7])).floor().cuda()
            do, so = dense.step(a), sparse.step(a)
        torch.cuda.synchronize()
        assert torch.equal(do.counts, so.counts) and torch.equal(do.num_landmarks, so.num_landmarks)
        assert torch.equal(do.balls_seen, so.balls_seen) and torch.equal(do.robots_seen, so.robots_seen)
        assert torch.equal(do.rewards, so.rewards) and torch.equal(dense.get_state(), sparse.get_state())
        for k, (name, cap) in enumerate(lists):
            dr, sr = getattr(do, name), getattr(so, name)                      # [n, T, P, cap, F]
            live = torch.arange(cap, device="cuda")[None, None, None, :] < do.counts[..., k:k + 1]
            assert torch.equal(dr[live], sr[live]), name
            assert torch.isnan(sr[~live]).all(), name                         # nothing behind the counts was touched
            assert (dr[~live] == 0).all(), name
        written_bytes += int((~torch.isnan(so.obs)).sum().item()) * 4
        steps += 1
        ka, kb = KernelArranger.for_robocup(dense), KernelArranger.for_robocup(sparse)
        ia, (ca, ma, oa) = ka.rearrange_inputs()
        ib, (cb, mb, ob) = kb.rearrange_inputs()
        assert int(ma) == int(mb) and torch.equal(ca, cb) and torch.equal(oa, ob)
        for x, y in zip(ka.compact(ia), kb.compact(ib)):
            assert torch.equal(x, y)
    per_env_step = written_bytes / steps / n
    print("sparse stream: %.1f KB of observations written per env-step (padded records: %.1f KB)" % (
        per_env_step / 1e3, 5 * P * L.RC_OBS_WORDS * 4 / 1e3))
    assert per_env_step < 30e3


def test_byte_actions_step_like_float_actions():
    """dynenv_step_u8: MultiDiscrete action indices as bytes give the step the float actions give, bit for bit (RoboCup and
    Driving, an odd number of Driving action words included); continuous action spaces refuse them"""
    import torch
    import dynenv_b200
    kw = dict(noiseMagnitude=2.0, canFall=True, auto_reset=True, seed=21)
    a, b = dynenv_b200.BatchedRoboCup(77, 5, **kw), dynenv_b200.BatchedRoboCup(77, 5, **kw)
    a.reset(); b.reset()
    g = torch.Generator().manual_seed(4)
    for s in range(6):
        act = (torch.rand((77, 10, 4), generator=g) * torch.tensor([5, 3, 3, 7])).floor()
        oa, ob = a.step(act.cuda()), b.step(act.to(torch.uint8).cuda())
        torch.cuda.synchronize()
        assert torch.equal(a.get_state(), b.get_state()) and torch.equal(oa.obs.view(torch.int32), ob.obs.view(torch.int32))
        assert torch.equal(oa.info.view(torch.int32), ob.info.view(torch.int32))
    assert b.last_launch_count() == a.last_launch_count() + 1
    bad = torch.zeros((77, 10, 4), dtype=torch.uint8)
    bad[3, 2, 0] = 9
    assert int(b.step(bad.cuda()).diag[3].item()) & dynenv_b200.layout.RC_IF_BAD_ACTION
    d0, d1 = [dynenv_b200.BatchedDriving(33, 3, noiseMagnitude=2.0, continuousActions=False, auto_reset=True, seed=2) for _ in range(2)]
    d0.reset(); d1.reset()
    for s in range(6):
        act = (torch.rand((33, 3, 2), generator=g) * 3).floor()  # 198 action words: not a multiple of four
        o0, o1 = d0.step(act.cuda()), d1.step(act.to(torch.uint8).cuda())
        torch.cuda.synchronize()
        assert torch.equal(d0.get_state(), d1.get_state()) and torch.equal(o0.record.view(torch.int32), o1.record.view(torch.int32))
    cont = dynenv_b200.BatchedRoboCup(4, 2, allowHeadTurn=True, seed=1)
    cont.reset()
    with pytest.raises(Exception, match="MultiDiscrete"):
        cont.step(torch.zeros((4, 4, 4), dtype=torch.uint8).cuda())


def test_step_replayed_as_a_cuda_graph():
    """the ~31 launches of a RoboCup step (two streams, fork / join events, a device-side work queue) captured once over a
    static action buffer and replayed give what the direct calls give, bit for bit, auto-resets included"""
    import torch
    import dynenv_b200
    kw = dict(noiseMagnitude=2.0, canFall=True, auto_reset=True, seed=8, sparse_obs=True)
    a, b = dynenv_b200.BatchedRoboCup(512, 5, **kw), dynenv_b200.BatchedRoboCup(512, 5, **kw)
    a.reset(); b.reset()
    buf = torch.zeros((512, 10, 4), dtype=torch.uint8, device="cuda")
    b.step(buf)  # (the byte path allocates on first use: outside the capture)
    a.step(buf)
    torch.cuda.synchronize()
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            b.step(buf)
    torch.cuda.current_stream().wait_stream(cap)
    g = torch.Generator().manual_seed(12)
    for s in range(8):
        act = (torch.rand((512, 10, 4), generator=g) * torch.tensor([5, 3, 3, 7])).floor().to(torch.uint8).cuda()
        oa = a.step(act)
        buf.copy_(act)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(a.get_state(), b.get_state())
        assert torch.equal(oa.info.view(torch.int32), b.out.info.view(torch.int32)) and torch.equal(oa.counts, b.out.counts)
    # the byte path's first call allocates: under capture it must refuse (and leave the capture intact), not break it
    c = dynenv_b200.BatchedRoboCup(64, 5, **kw)
    c.reset()
    small = torch.zeros((64, 10, 4), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with pytest.raises(Exception, match="outside stream capture"):
        with torch.cuda.stream(cap):
            with torch.cuda.graph(g2, stream=cap):
                c.step(small)
    torch.cuda.synchronize()
    c.step(small)  # fine outside
    torch.cuda.synchronize()
