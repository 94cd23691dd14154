"""GPU: what stands in for `compute-sanitizer --tool racecheck` (closed on the GPU pool).

The step kernels order their shared-memory phases with __syncwarp (one warp = one environment) and, in rc_slow.cu, with
block barriers between warps that share nothing but the instruction stream.  A missing barrier shows up as a result that
depends on timing, so the same work is repeated under different schedules and must come out bit-identical:
  * the same environments stepped alone, beside a kernel that keeps every SM busy on another stream, and as the first
    half of a batch twice the size (other block-to-SM placement, other co-residents);
  * the handed-over environments through two independent implementations of their kernel, one step from the same states
    60 steps into an episode: the phase-synchronised multi-warp blocks of rc_slow.cu and the first one-warp-per-block
    version (-DRC_SLOW_ONE_WARP), which has no cross-warp barrier at all and visits the environments in list order
    instead of by hand-over bin.  (Separate compilations: integers equal, continuous values to the last bits.)
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_RUN = r"""
import sys
import numpy as np
import torch
sys.path.insert(0, %(root)r)
import dynenv_b200
E, steps = %(envs)d, %(steps)d
sim = dynenv_b200.BatchedRoboCup(E, 5, 1, 1, 2.0, canFall=True, auto_reset=True, seed=77)
sim.reset()
g = torch.Generator().manual_seed(3)
hi = torch.tensor([5, 3, 3, 7])
if sys.argv[2] == "make":      # the states the comparison starts from: `steps` steps into the episode
    for s in range(steps):
        sim.step((torch.rand((E, 10, 4), generator=g) * hi).floor().cuda())
    torch.cuda.synchronize()
    np.save(sys.argv[1], sim.get_state().cpu().numpy())
else:                          # one step from those states
    sim.set_state(torch.from_numpy(np.load(sys.argv[2])).cuda())
    out = sim.step((torch.rand((E, 10, 4), generator=g) * hi).floor().cuda())
    torch.cuda.synchronize()
    np.save(sys.argv[1], sim.get_state().cpu().numpy())
    np.save(sys.argv[1] + ".obs.npy", out.obs.cpu().numpy())
    np.save(sys.argv[1] + ".info.npy", out.info.cpu().numpy())
"""


def _rollout(sim, E, steps, load=False):
    import torch
    g = torch.Generator().manual_seed(3)
    hi = torch.tensor([5, 3, 3, 7])
    sim.reset()
    side = torch.cuda.Stream()
    junk = torch.rand((4096, 4096), device="cuda")
    snaps = []
    for s in range(steps):
        a = (torch.rand((E, 10, 4), generator=g) * hi).floor().cuda()
        if load:  # every SM busy with foreign blocks while the step kernels run
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    junk = torch.tanh(junk @ junk * 1e-3)
        out = sim.step(a)
        if s % 6 == 5 or s == steps - 1:
            snaps.append((sim.get_state().clone(), out.obs.clone(), out.info.clone()))
    torch.cuda.synchronize()
    return snaps


def test_repeated_runs_are_bit_identical_whatever_the_schedule():
    import torch
    import dynenv_b200
    E, steps = 4096, 36
    kw = dict(noiseMagnitude=2.0, canFall=True, auto_reset=True, seed=77)
    ref = _rollout(dynenv_b200.BatchedRoboCup(E, 5, **kw), E, steps)
    again = _rollout(dynenv_b200.BatchedRoboCup(E, 5, **kw), E, steps, load=True)
    for (s0, o0, i0), (s1, o1, i1) in zip(ref, again):
        assert torch.equal(s0, s1) and torch.equal(o0.view(torch.int32), o1.view(torch.int32)) and torch.equal(i0.view(torch.int32), i1.view(torch.int32))
    # the same environments (global ids 0 .. E-1) as the first half of a batch of 2 E: other placement, other neighbours
    big = dynenv_b200.BatchedRoboCup(2 * E, 5, **kw)
    g = torch.Generator().manual_seed(3)
    hi = torch.tensor([5, 3, 3, 7])
    big.reset()
    k = 0
    for s in range(steps):
        a = (torch.rand((E, 10, 4), generator=g) * hi).floor().cuda()
        out = big.step(torch.cat([a, a.flip(0)], 0).contiguous())
        if s % 6 == 5 or s == steps - 1:
            s0, o0, i0 = ref[k]
            k += 1
            assert torch.equal(big.get_state()[:E], s0) and torch.equal(out.obs[:E].view(torch.int32), o0.view(torch.int32))
    # Driving: the same with its two kernels
    drv = [dynenv_b200.BatchedDriving(2048, 10, noiseMagnitude=2.0, continuousActions=True, auto_reset=True, seed=5) for _ in range(2)]
    side = torch.cuda.Stream()
    junk = torch.rand((4096, 4096), device="cuda")
    for d in drv:
        d.reset()
    gd = torch.Generator().manual_seed(8)
    for s in range(40):
        a = (torch.rand((2048, 10, 2), generator=gd) * 6 - 3).cuda()
        o0 = drv[0].step(a)
        with torch.cuda.stream(side):
            junk = torch.tanh(junk @ junk * 1e-3)
        o1 = drv[1].step(a)
        torch.cuda.synchronize()
        assert torch.equal(o0.record.view(torch.int32), o1.record.view(torch.int32)) and torch.equal(drv[0].get_state(), drv[1].get_state())


def test_handed_over_kernel_agrees_with_its_one_warp_version(tmp_path):
    """~10 % of the environments per substep block take the hand-over path by step 60"""
    from dynenv_b200 import build as B
    from dynenv_b200 import layout as L
    from helpers import IMPULSE_SCALE
    import shutil
    if shutil.which(os.environ.get("NVCC", "nvcc")) is None:
        pytest.skip("nvcc not available for the variant build")
    out_dir = os.path.join(ROOT, "tests", "native", "_build")
    os.makedirs(out_dir, exist_ok=True)
    variant = os.path.join(out_dir, "libdynenv_onewarp.so")
    srcs = [os.path.join(B._DIR, s) for s in B.SOURCES]
    newest = max(os.path.getmtime(os.path.join(B._DIR, "csrc", f)) for f in os.listdir(os.path.join(B._DIR, "csrc")))
    if not os.path.exists(variant) or os.path.getmtime(variant) < newest:
        subprocess.check_call([os.environ.get("NVCC", "nvcc")] + B.NVCC_FLAGS + ["-DRC_SLOW_ONE_WARP", "-o", variant] + srcs)
    script = tmp_path / "run.py"
    script.write_text(_RUN % dict(root=ROOT, envs=8192, steps=60))

    def run(lib, *argv):
        p = subprocess.run([sys.executable, str(script)] + list(argv), env=dict(os.environ, DYNENV_B200_LIB=lib), capture_output=True,
                           text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
    start = str(tmp_path / "start.npy")
    run(B.LIB_PATH, start, "make")
    res = {}
    for name, lib in (("product", B.LIB_PATH), ("onewarp", variant)):
        dst = str(tmp_path / (name + ".npy"))
        run(lib, dst, start)
        res[name] = (np.load(dst), np.load(dst + ".obs.npy"), np.load(dst + ".info.npy"))
    s0 = np.load(start)
    assert int((s0[:, L.RC_FIELDS["nArb"][0]] > 0).sum()) > 50  # the path under test is taken: cached pairs present
    # The two kernels are separate compilations of the same source: the compiler contracts multiply-adds differently in
    # them, so continuous values may differ in the last bits (1e-6 relative covers ten sweeps of a contact solve); every
    # integer -- flags, timers, cached-pair keys, counts, seen flags -- must be equal.
    a, b = res["product"][0], res["onewarp"][0]
    for name, (off, cnt, kind) in L.RC_FIELDS.items():
        x, y = a[:, off:off + cnt], b[:, off:off + cnt]
        if kind == "i":
            assert np.array_equal(x, y), name
        else:
            xf, yf = x.view(np.float32).astype(np.float64), y.view(np.float32).astype(np.float64)
            if name in IMPULSE_SCALE:  # solver accumulators in the unit they act in (tests/helpers.py)
                xf, yf = xf * IMPULSE_SCALE[name], yf * IMPULSE_SCALE[name]
            low = name.endswith("Lo")  # low words of two-float numbers: absolute, one ulp of the high word
            err = (np.abs(xf - yf) / (1.0 if low else 1.0 + np.abs(yf))).max()
            assert err < (1e-4 if low else 1e-6), (name, float(err))
    oa, ob = res["product"][1], res["onewarp"][1]
    ia, ib = oa.view(np.int32)[..., L.RC_O_COUNTS:L.RC_OBS_WORDS_RAW], ob.view(np.int32)[..., L.RC_O_COUNTS:L.RC_OBS_WORDS_RAW]
    assert np.array_equal(ia, ib)
    d = np.abs(oa[..., :L.RC_O_COUNTS].astype(np.float64) - ob[..., :L.RC_O_COUNTS])
    # (the direction of a clipped field-line piece shorter than a pixel amplifies the last bit of the head angle: helpers.LINE_TOL)
    assert d[..., :L.RC_O_LINE].max() < 1e-4 and d[..., L.RC_O_LINE:].max() < 5e-3, (float(d[..., :L.RC_O_LINE].max()), float(d[..., L.RC_O_LINE:].max()))
    assert np.abs(res["product"][2][:, :L.RC_I_DONE].astype(np.float64) - res["onewarp"][2][:, :L.RC_I_DONE]).max() < 1e-4
    assert np.array_equal(res["product"][2].view(np.int32)[:, L.RC_I_DONE], res["onewarp"][2].view(np.int32)[:, L.RC_I_DONE])
