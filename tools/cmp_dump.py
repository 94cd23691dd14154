"""Dev tool (here): compares gpurun_out/dump_<name>.npz with the oracle, field by field."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import compare_step, golden_cfg, golden_names, load_golden, oracle_step_batch  # noqa: E402
from oracle.oracle import f32_record_to_words  # noqa: E402


class _T(object):
    def __init__(self, a):
        self.a = a

    def detach(self):
        return self

    def cpu(self):
        return self

    def numpy(self):
        return self.a


class _Out(object):
    pass


tol = float(os.environ.get("TOL", "1e-4"))
for name in (sys.argv[1:] or golden_names()):
    p = os.path.join(ROOT, "gpurun_out", "dump_%s.npz" % name)
    if not os.path.exists(p):
        continue
    d = np.load(p)
    g = load_golden(name)
    cfg = golden_cfg(g)
    out = _Out()
    for k in d.files:
        setattr(out, k, _T(d[k]))
    outs, ow = oracle_step_batch(cfg, d["recs"], g["actions"], int(d["seed"]))
    bad = compare_step(cfg, out, f32_record_to_words(d["state_after"]), outs, ow, tol)
    envs = sorted(set(b[0] for b in bad))
    print("%-24s envs=%3d bad envs=%3d" % (name, len(d["recs"]), len(envs)))
    for b in bad[:int(os.environ.get("SHOW", "6"))]:
        print("    ", b[0], b[1], str(b[2])[:150])
