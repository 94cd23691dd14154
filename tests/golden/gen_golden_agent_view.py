"""Generates tests/golden/agent_view.npz: the agent-view debug image the REFERENCE's own Python draws in getAgentVision
(RoboCupEnvironment.py:1469-1531; /root/reference/DynEnv, unmodified, real cv2), together with the detection lists it was
drawn from.  Run here only: python tests/golden/gen_golden_agent_view.py

How: the drawing sits inside getAgentVision behind `renderVar and agentVisID == agent.id`; with renderMode 'memory' it
needs neither pygame nor a window and appends the image to env.obsVis.  The environment is built and stepped without
rendering (step() would call pygame's debug draw, which this image does not have); the two attributes are set only
around direct getAgentVision(agent) calls.  The lists the drawing reads are locals of that function: a line tracer copies
them when execution reaches the drawing's `if` (found by its text, not by its number).

Stored per sample: the lists as the raw detection record of the C ABI lays them out (include/dynenv_b200.h, RC_R_*: rows
of x, y, size in the observer's head frame, lines as x1, y1, x2, y2, the centre circle) in float64, their sighting types
(the drawing dims Partial / Distant sightings; the record does not carry the type), and the image."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dynenv_b200 import layout as L  # noqa: E402
from oracle.ref_loader import UniformTape, load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LISTS = (("ball", "ballDets", L.RC_CAP_BALL), ("robot", "robDets", L.RC_CAP_ROBOT), ("post", "goalDets", L.RC_CAP_POST),
         ("cross", "crossDets", L.RC_CAP_CROSS), ("fcross", "fieldCrossDets", L.RC_CAP_FCROSS))
# (nPlayers, noiseMagnitude, steps before the views are taken, seed)
SCENES = ((5, 2.0, 6, 71), (5, 0.0, 3, 72), (2, 3.0, 10, 73))


def capture_views(ns, env, agents):
    """-> [(lists, image)] of getAgentVision(agent) for the given agents"""
    import inspect
    src, first = inspect.getsourcelines(ns.RoboCupEnvironment.getAgentVision)
    draw_line = first + next(i for i, s in enumerate(src) if "self.renderVar and self.agentVisID is not None" in s)
    code = ns.RoboCupEnvironment.getAgentVision.__code__
    got = {}

    def tracer(frame, event, arg):
        if frame.f_code is not code:
            return None

        def local(frame, event, arg):
            if event == "line" and frame.f_lineno == draw_line:
                loc = frame.f_locals
                d = {}
                for name, var, cap in LISTS:
                    d[name] = [(float(r[1].x), float(r[1].y), float(r[2]), int(r[0])) for r in loc[var]]
                d["line"] = [(float(r[1].x), float(r[1].y), float(r[2].x), float(r[2].y), int(r[0])) for r in loc["lineDets"]]
                c = loc["circleDets"]
                d["circle"] = (int(c[0]), float(c[1].x), float(c[1].y), float(c[2])) if int(c[0]) != int(ns.cutils.SightingType.NoSighting) else (int(c[0]), 0.0, 0.0, 0.0)
                got["lists"] = d
            return local
        return local

    views = []
    env.renderMode = "memory"
    for ag in agents:
        env.renderVar, env.agentVisID = True, ag.id
        env.obsVis.clear()
        sys.settrace(tracer)
        try:
            env.getAgentVision(ag)
        finally:
            sys.settrace(None)
            env.renderVar, env.agentVisID = False, None
        assert len(env.obsVis) == 1
        views.append((got.pop("lists"), np.array(env.obsVis[0])))
    return views


def main():
    ns = load_reference()
    ST = ns.cutils.SightingType
    recs, types, images = [], [], []
    for nPlayers, mag, steps, seed in SCENES:
        tape = UniformTape(seed)
        with tape.installed():
            env = ns.RoboCupEnvironment(nPlayers, observationType=ns.ObservationType.PARTIAL, noiseType=ns.NoiseType.REALISTIC,
                                        noiseMagnitude=mag)
            env.reset()
            rng = np.random.default_rng(seed)
            for _ in range(steps):
                env.step(np.stack([rng.integers(0, 5, 2 * nPlayers), rng.integers(0, 3, 2 * nPlayers), rng.integers(0, 3, 2 * nPlayers),
                                   rng.integers(0, 7, 2 * nPlayers)], axis=1))
            views = capture_views(ns, env, env.agents[::2] if nPlayers > 2 else env.agents)
        for lists, img in views:
            if any(len(lists[name]) > cap for name, _, cap in LISTS) or len(lists["line"]) > L.RC_CAP_LINE:
                continue
            rec = np.zeros(L.RC_RAW_WORDS, np.float64)      # counts / circle-seen as numbers here (int32 words in the float32 record)
            ty = np.zeros(L.RC_RAW_WORDS, np.int8)          # sighting type per row, at the row's first word
            for k, (name, _, cap) in enumerate(LISTS):
                off = getattr(L, "RC_R_" + name.upper())
                rec[L.RC_R_COUNTS + k] = len(lists[name])
                for i, (x, y, s, t) in enumerate(lists[name]):
                    rec[off + 3 * i:off + 3 * i + 3] = (x, y, s)
                    ty[off + 3 * i] = t
            rec[L.RC_R_COUNTS + len(LISTS)] = len(lists["line"])
            for i, (x1, y1, x2, y2, t) in enumerate(lists["line"]):
                rec[L.RC_R_LINE + 4 * i:L.RC_R_LINE + 4 * i + 4] = (x1, y1, x2, y2)
                ty[L.RC_R_LINE + 4 * i] = t
            c = lists["circle"]
            rec[L.RC_R_CIRCLE_SEEN] = int(c[0] != int(ST.NoSighting))
            rec[L.RC_R_CIRCLE:L.RC_R_CIRCLE + 3] = c[1:]
            ty[L.RC_R_CIRCLE] = c[0]
            recs.append(rec)
            types.append(ty)
            images.append(img)
    np.savez_compressed(os.path.join(OUT, "agent_view.npz"), records=np.stack(recs), types=np.stack(types),
                        images=np.stack(images), normal=np.array(int(ST.Normal)))
    print("%d views, image %s, rows per view: %s" % (len(recs), images[0].shape, [int(r[L.RC_R_COUNTS:L.RC_R_COUNTS + 6].sum()) for r in recs]))


if __name__ == "__main__":
    main()
