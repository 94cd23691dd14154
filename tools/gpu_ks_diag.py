"""GPU box: count histograms of the noise-statistics test, GPU (Philox) vs oracle (sequential tape), larger oracle sample"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from scipy import stats
import dynenv_b200
from dynenv_b200 import layout as L
from oracle.oracle import RcOracle
n_gpu, n_cpu, P = 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 3000, 4
names = ["ball", "robot", "post", "cross", "fcross", "line"]
rng = np.random.default_rng(1234)
flags = L.DYNENV_FLAG_CAN_FALL | L.DYNENV_FLAG_USE_OBS_REWARDS
Cc = []
Crob = []
for i in range(n_cpu):
    orc = RcOracle(2, 1, 1, 3.0, False, flags)
    orc.set_tape(rng.random(6000))
    orc.reset()
    o = orc.reset_obs()
    Cc.append(np.array(o["counts"][:, :P])); Crob.append(np.array(o["robot"][:, :P]))
Cc = np.array(Cc); Crob = np.array(Crob)
for seed in (6, 7):
    sim = dynenv_b200.BatchedRoboCup(n_gpu, 2, 1, 1, 3.0, canFall=True, seed=seed)
    out = sim.reset(); torch.cuda.synchronize()
    Gc = out.counts.cpu().numpy(); Grob = out.robot.cpu().numpy()
    print("seed", seed)
    for a in range(P):
        for k in range(6):
            gc, cc = Gc[:, :, a, k].ravel(), Cc[:, :, a, k].ravel()
            hi = int(max(gc.max(), cc.max())) + 1
            hg, hc = np.bincount(gc, minlength=hi) / len(gc), np.bincount(cc, minlength=hi) / len(cc)
            print("  obs %d %-6s gpu %s | cpu %s" % (a, names[k], np.round(hg, 4).tolist(), np.round(hc, 4).tolist()))
    gm, cm = Gc[..., 1] > 0, Cc[..., 1] > 0
    for col in range(3):
        print("  robot col", col, stats.ks_2samp(Grob[..., 0, col][gm], Crob[..., 0, col][cm]))
