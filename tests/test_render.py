"""Top-down renderer: frames drawn from state records.  CPU: records from the oracle's reset; GPU: through
the drop-in classes.  pygame is not available, so the checks are geometric (colours at the places the bodies are)."""
import numpy as np
import pytest

from dynenv_b200 import layout as L
from dynenv_b200 import render as R


def test_robocup_frame_from_oracle_state():
    from oracle.oracle import RcOracle, words_to_f32_record
    orc = RcOracle(5, 1, 1, 2.0, False, L.DYNENV_FLAG_CAN_FALL, seed=3, env_id=0)
    orc.reset()
    w = orc.get_state()
    img = R.render_robocup(words_to_f32_record(w), 5)
    assert img.shape == (740, 1040, 3) and img.dtype == np.uint8
    off = L.RC_FIELDS["px"][0]
    bx, by = w[off + 20], w[L.RC_FIELDS["py"][0] + 20]
    assert tuple(img[int(round(by)), int(round(bx))]) == (255, 0, 0)            # the ball, drawn last
    assert tuple(img[10, 10]) == (0, 255, 0) and tuple(img[70, 500]) == (255, 255, 255)  # grass, the top side line
    for r, col in ((0, (255, 0, 254)), (5, (255, 254, 0))):                      # left foot of a robot of either team
        fx, fy = w[off + 2 * r], w[L.RC_FIELDS["py"][0] + 2 * r]
        assert tuple(img[int(round(fy)), int(round(fx))]) == col, (r, img[int(round(fy)), int(round(fx))])
    assert tuple(img[740 // 2 - 80, 70]) == (0, 0, 255)                          # a goal post


def test_driving_frame_from_oracle_state():
    from oracle.oracle import DrvOracle, drv_words_to_f32_record
    orc = DrvOracle(10, 1, 1, 2.0, True, seed=4, env_id=1)
    orc.reset()
    w = orc.get_state()
    img = R.render_driving(drv_words_to_f32_record(w), 10)
    assert img.shape == (1000, 1750, 3)
    F = L.DRV_FIELDS
    inside = 0
    for k in range(10):  # every car's centre is car-coloured
        x = w[F["cpx"][0] + k] + w[F["cpxLo"][0] + k]
        y = w[F["cpy"][0] + k] + w[F["cpyLo"][0] + k]
        if 0 <= int(round(y)) < 1000 and 0 <= int(round(x)) < 1750:  # cars start at the ends of the roads, on the border
            inside += 1
            assert tuple(img[int(round(y)), int(round(x))]) == (0, 255, 0), k
    assert inside >= 5
    n_ped = int(w[F["nPed"][0]])
    seen = 0
    for k in range(n_ped):
        x = w[F["ppx"][0] + k] + w[F["ppxLo"][0] + k]
        y = w[F["ppy"][0] + k] + w[F["ppyLo"][0] + k]
        if 0 <= int(round(y)) < 1000 and 0 <= int(round(x)) < 1750:
            seen += tuple(img[int(round(y)), int(round(x))]) in ((0, 255, 255), (0, 255, 0))
    assert seen >= n_ped - 2
    assert tuple(img[200, 365]) == (200, 200, 200) and tuple(img[510, 100]) == (50, 50, 50)  # a building, the road surface
    assert tuple(img[100, 875]) == (255, 255, 255)                                      # the middle lane line of road 0


@pytest.mark.gpu
def test_render_through_the_drop_in_classes():
    import dynenv_b200
    env = dynenv_b200.RoboCupEnvironment(2, render=True)
    assert env.renderMode == 'memory' and len(env.screenShots) == 1
    env.step(np.zeros((4, 4)))
    assert len(env.screenShots) == 2 and env.screenShots[1].shape == (740, 1040, 3)
    env.close()
    drv = dynenv_b200.DrivingEnvironment(4, render=False)
    frame = drv.render()
    assert frame.shape == (1000, 1750, 3) and drv.screenShots == [] and int((frame == (0, 255, 0)).all(-1).sum()) > 100
    drv.close()


def _golden_views():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agent_view.npz"))
    return g["records"], g["types"], g["images"], int(g["normal"])


def test_agent_view_is_the_image_the_reference_draws():
    """tests/golden/agent_view.npz: images drawn by the reference's own getAgentVision (RoboCupEnvironment.py:1469-1516)
    and the detection lists they were drawn from, laid out as raw records"""
    recs, types, images, normal = _golden_views()
    assert len(recs) >= 10
    dimmed = 0
    for rec, ty, ref in zip(recs, types, images):
        starts = np.zeros(L.RC_RAW_WORDS, bool)  # first words of the rows in use
        for k, (off, w) in enumerate(((L.RC_R_BALL, 3), (L.RC_R_ROBOT, 3), (L.RC_R_POST, 3), (L.RC_R_CROSS, 3), (L.RC_R_FCROSS, 3), (L.RC_R_LINE, 4))):
            starts[off:off + w * int(rec[L.RC_R_COUNTS + k]):w] = True
        starts[L.RC_R_CIRCLE] = bool(rec[L.RC_R_CIRCLE_SEEN])
        dim = starts & (ty != normal)
        dimmed += int(dim.sum())
        img = R.render_agent_view(rec, dim=dim)
        assert img.shape == ref.shape == (940, 1040, 3) and img.dtype == np.uint8
        assert np.array_equal(img, ref)  # pixel for pixel, colours included
        # without the sighting types (a device record has none): the same pixels, all of them bright
        bright = R.render_agent_view(rec)
        assert np.array_equal(bright.any(axis=2), ref.any(axis=2))
        assert set(np.unique(bright)) <= {0, 255}
    assert dimmed > 0  # the golden does contain Partial / Distant sightings


def test_agent_view_from_float32_record_words():
    """the same record as the float32 / int32 words of the C ABI (DynEnvRoboCupOut.raw): float32 rounding may move an
    int() across a pixel boundary, nothing else"""
    recs, types, images, normal = _golden_views()
    for rec, ref in zip(recs, images):
        w = rec.astype(np.float32)
        wi = w.view(np.int32)
        wi[L.RC_R_COUNTS:L.RC_R_COUNTS + L.RC_NTYPES] = rec[L.RC_R_COUNTS:L.RC_R_COUNTS + L.RC_NTYPES].astype(np.int32)
        wi[L.RC_R_CIRCLE_SEEN] = int(rec[L.RC_R_CIRCLE_SEEN])
        img = R.render_agent_view(w)
        diff = int((img.any(axis=2) != ref.any(axis=2)).sum())
        assert diff <= 0.01 * int(ref.any(axis=2).sum()), diff
        assert np.array_equal(R.render_agent_view(w.view(np.uint32)), img)


def test_agent_view_of_an_oracle_record():
    """layout check against the producer of raw records on the CPU side: the C oracle in IMAGE mode (same RC_R_* words the
    observation kernel writes); every detection of the record is on the canvas where the record says"""
    from oracle.oracle import RcOracle, raw_to_record
    orc = RcOracle(3, L.DYNENV_OBS_IMAGE, 1, 2.0, False, L.DYNENV_FLAG_CAN_FALL, seed=9, env_id=2)
    orc.capture_raw()
    orc.reset()
    orc.reset_obs()
    seen = 0
    for ag in range(6):
        raw = orc.raw_vision(0, ag)
        img = R.render_agent_view(raw_to_record(raw))
        for name, col in (("ball", (0, 0, 255)), ("robot", (0, 255, 0)), ("post", (255, 0, 0))):
            k = ("ball", "robot", "post").index(name)
            for i in range(int(raw["counts"][k])):
                x, y, r = raw[name][i]
                px, py = int(R.AV_XOFFS + np.float32(x)), int(-np.float32(y) + R.AV_H)
                if int(np.float32(r)) >= 1 and 0 <= px < 1040 and 0 <= py < 940:
                    assert tuple(img[py, px]) in ((0, 0, 255), (0, 255, 0), (255, 0, 0)), (ag, name, i)  # (a later list may cover it)
                    seen += 1
        assert tuple(img[R.AV_H, R.AV_XOFFS]) != (0, 0, 0)  # the field-of-view rays start at the observer
    assert seen > 0


@pytest.mark.gpu
def test_agent_view_through_the_drop_in_class():
    """IMAGE mode: the raw detection records of the observation kernel drawn as the reference's agent view (the same
    lists the camera planes are drawn from)"""
    import dynenv_b200
    from dynenv_b200 import ObservationType
    env = dynenv_b200.RoboCupEnvironment(2, render=True, observationType=ObservationType.IMAGE, noiseMagnitude=0)
    env.agentVisID = 1
    obs, _, _, _ = env.step(np.zeros((4, 4)))
    assert obs[0][0].shape == (8, 480, 640)
    assert len(env.obsVis) == 5 and env.obsVis[0].shape == (940, 1040, 3) and env.obsVis[0].dtype == np.uint8
    raw = env._sim.out.raw[0].cpu().numpy()
    for t in range(5):
        for a in range(4):
            img = env.agent_view(a, t)
            if a == 1:
                assert np.array_equal(img, env.obsVis[t])
            assert (img == (255, 255, 0)).all(-1).sum() > 100                      # the field-of-view rays
            n_ball = int(raw[t, a].view(np.int32)[L.RC_R_COUNTS])
            x, y, r = raw[t, a][L.RC_R_BALL:L.RC_R_BALL + 3]
            px, py = int(R.AV_XOFFS + x), int(-y + R.AV_H)
            if n_ball and int(r) >= 1 and 0 <= px < 1040 and 0 <= py < 940:
                assert tuple(img[py, px]) == (0, 0, 255)                           # balls are drawn last
            elif not n_ball:
                assert not (img == (0, 0, 255)).all(-1).any()
    env.close()
    plain = dynenv_b200.RoboCupEnvironment(2)
    with pytest.raises(RuntimeError):
        plain.agent_view(0)
    plain.close()
