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
