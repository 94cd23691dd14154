"""Dev tool (here): first substep at which the GPU trace and the oracle disagree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden_cfg, load_golden, rel_err  # noqa: E402
from oracle.oracle import RcOracle, f32_record_to_words  # noqa: E402
from dynenv_b200 import layout as L  # noqa: E402

np.set_printoptions(linewidth=220, precision=6, suppress=True)
name, env = sys.argv[1], int(sys.argv[2])
resync = len(sys.argv) > 3 and sys.argv[3] == "resync"
d = np.load(os.path.join(ROOT, "gpurun_out", "trace_%s_%d.npz" % (name, env)))
g = load_golden(name)
cfg = golden_cfg(g)
orc = RcOracle(cfg["nPlayers"], cfg["obsType"], cfg["noiseType"], cfg["noiseMag"], cfg["allowHeadTurn"], cfg["flags"],
               seed=int(d["seed"]), env_id=env)
orc.set_state(f32_record_to_words(d["rec0"]))
tol = float(os.environ.get("TOL", "1e-4"))
for i in range(50):
    orc.debug_substep_range(i, i + 1)
    orc.step(g["actions"][env])
    wo = orc.get_state()
    wg = f32_record_to_words(d["trace"][i])
    bad = []
    for nm, (off, cnt, kind) in L.RC_FIELDS.items():
        a, b = wg[off:off + cnt], wo[off:off + cnt]
        if (kind == "i" and not np.array_equal(a, b)) or (kind == "f" and rel_err(a, b) > tol):
            bad.append(nm)
    if bad:
        print("substep", i, "differs in", bad)
        for nm in bad[:12]:
            off, cnt, kind = L.RC_FIELDS[nm]
            print("  ", nm, "\n     gpu", wg[off:off + cnt], "\n     orc", wo[off:off + cnt])
        if not resync:
            break
    if resync:
        orc.set_state(wg)
else:
    print("no difference over 50 substeps")
