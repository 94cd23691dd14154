"""State-record layout and capacities, read from include/dynenv_b200.h (single source of truth).

The header's `#define NAME value` lines and the `enum { NAME = expr, ... }` block are
evaluated in order, so Python sees exactly the word offsets the CUDA kernels and the C-ABI
use.  `RC_FIELDS` maps a field name to (offset, count, kind) with kind 'f' (float32) or
'i' (int32) for viewing a state record as named arrays.
"""
import os
import re

_HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include",
                       "dynenv_b200.h")


def _parse(path):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    env = {}
    for m in re.finditer(r"^\s*#define\s+(\w+)\s+(.+?)\s*$", txt, flags=re.M):
        name, expr = m.group(1), m.group(2)
        expr = re.sub(r"(\d+)u\b", r"\1", expr)
        try:
            env[name] = eval(expr, {"__builtins__": {}}, env)
        except Exception:
            pass
    for block in re.finditer(r"enum\s*\{(.*?)\}\s*;", txt, flags=re.S):
        nxt = 0
        for item in block.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, expr = [s.strip() for s in item.split("=", 1)]
                val = eval(expr.replace("/", "//"), {"__builtins__": {}}, env)
            else:
                name, val = item, nxt
            env[name] = int(val)
            nxt = int(val) + 1
    return env


C = _parse(_HEADER)
globals().update({k: v for k, v in C.items() if isinstance(v, int)})

_NB, _NP, _NA = C["RC_NBODY"], C["RC_MAXP"], C["RC_MAXARB"]
RC_FIELDS = {
    # bodies
    "px": (C["RC_S_PX"], _NB, "f"), "py": (C["RC_S_PY"], _NB, "f"), "angle": (C["RC_S_ANG"], _NB, "f"),
    "vx": (C["RC_S_VX"], _NB, "f"), "vy": (C["RC_S_VY"], _NB, "f"), "w": (C["RC_S_W"], _NB, "f"),
    "fx": (C["RC_S_FX"], _NB, "f"), "fy": (C["RC_S_FY"], _NB, "f"), "torque": (C["RC_S_TQ"], _NB, "f"),
    "vbx": (C["RC_S_VBX"], _NB, "f"), "vby": (C["RC_S_VBY"], _NB, "f"), "wb": (C["RC_S_WB"], _NB, "f"),
    "cax": (C["RC_S_CAX"], _NB, "f"), "cay": (C["RC_S_CAY"], _NB, "f"),
    "cbx": (C["RC_S_CBX"], _NB, "f"), "cby": (C["RC_S_CBY"], _NB, "f"),
    # robots
    "headAngle": (C["RC_S_HEAD"], _NP, "f"), "headMoving": (C["RC_S_HEADMOV"], _NP, "f"),
    "moveTime": (C["RC_S_MOVET"], _NP, "f"), "initX": (C["RC_S_INITX"], _NP, "f"),
    "initY": (C["RC_S_INITY"], _NP, "f"), "fallTime": (C["RC_S_FALLT"], _NP, "f"),
    "penalTime": (C["RC_S_PENT"], _NP, "f"), "prevX": (C["RC_S_PREVX"], _NP, "f"),
    "prevY": (C["RC_S_PREVY"], _NP, "f"), "pivJx": (C["RC_S_PIVJX"], _NP, "f"),
    "pivJy": (C["RC_S_PIVJY"], _NP, "f"), "rotJ": (C["RC_S_ROTJ"], _NP, "f"),
    "rflags": (C["RC_S_RFLAGS"], _NP, "i"), "fallCntr": (C["RC_S_FALLCNT"], _NP, "i"),
    "touchCntr": (C["RC_S_TOUCHCNT"], _NP, "i"),
    "episodeRewards": (C["RC_S_EPR"], _NP, "f"), "episodePosRewards": (C["RC_S_EPPR"], _NP, "f"),
    "episodeObsRewards": (C["RC_S_EPOR"], _NP, "f"),
    # ball + env
    "ballPrevX": (C["RC_S_BPREVX"], 1, "f"), "ballPrevY": (C["RC_S_BPREVY"], 1, "f"),
    "lastKicked": (C["RC_S_LASTK"], 4, "i"), "nLastKicked": (C["RC_S_NLASTK"], 1, "i"),
    "elapsed": (C["RC_S_ELAPSED"], 1, "i"), "episode": (C["RC_S_EPISODE"], 1, "i"),
    "ballOwned": (C["RC_S_BALLOWNED"], 1, "i"), "ballFreeCntr": (C["RC_S_BALLFREE"], 1, "f"),
    "gracePeriod": (C["RC_S_GRACE"], 1, "f"), "goals": (C["RC_S_GOALS"], 2, "i"),
    "closestID": (C["RC_S_CLOSEST"], 2, "i"), "defenders": (C["RC_S_DEFENDERS"], 4, "i"),
    "nDefenders": (C["RC_S_NDEF"], 2, "i"), "penalTimes": (C["RC_S_PENALTIMES"], 2, "i"),
    "stepped": (C["RC_S_STEPPED"], 1, "i"),
    # cached contact pairs
    "nArb": (C["RC_S_NARB"], 1, "i"), "arbKey": (C["RC_S_ARBKEY"], _NA, "i"),
    "arbCon": (C["RC_S_ARBCON"], _NA, "i"), "arbJn0": (C["RC_S_ARBJN0"], _NA, "f"),
    "arbJt0": (C["RC_S_ARBJT0"], _NA, "f"), "arbJn1": (C["RC_S_ARBJN1"], _NA, "f"),
    "arbJt1": (C["RC_S_ARBJT1"], _NA, "f"),
    # low-order words of the two-float coordinates
    "pxLo": (C["RC_S_PXL"], _NB, "f"), "pyLo": (C["RC_S_PYL"], _NB, "f"), "angleLo": (C["RC_S_ANGL"], _NB, "f"),
    "initXLo": (C["RC_S_INITXL"], _NP, "f"), "initYLo": (C["RC_S_INITYL"], _NP, "f"),
}
# coordinate fields stored as hi + lo float32 pairs: hi field -> lo field
RC_SPLIT = {"px": "pxLo", "py": "pyLo", "angle": "angleLo", "initX": "initXLo", "initY": "initYLo"}


def rc_int_mask():
    """Boolean list over RC_STATE_WORDS: True where the word is an int32."""
    mask = [False] * C["RC_STATE_WORDS"]
    for off, cnt, kind in RC_FIELDS.values():
        if kind == "i":
            for k in range(off, off + cnt):
                mask[k] = True
    return mask


# ---------------------------------------------------------------- Driving state record
_NC, _NPD, _NO, _NDA = C["DRV_MAXCARS"], C["DRV_MAXPED"], C["DRV_MAXOBS"], C["DRV_MAXARB"]
DRV_FIELDS = {
    "cpx": (C["DRV_S_CPX"], _NC, "f"), "cpy": (C["DRV_S_CPY"], _NC, "f"), "cang": (C["DRV_S_CANG"], _NC, "f"),
    "cvx": (C["DRV_S_CVX"], _NC, "f"), "cvy": (C["DRV_S_CVY"], _NC, "f"), "cw": (C["DRV_S_CW"], _NC, "f"),
    "cvbx": (C["DRV_S_CVBX"], _NC, "f"), "cvby": (C["DRV_S_CVBY"], _NC, "f"), "cwb": (C["DRV_S_CWB"], _NC, "f"),
    "cdirx": (C["DRV_S_CDIRX"], _NC, "f"), "cdiry": (C["DRV_S_CDIRY"], _NC, "f"),
    "cprevx": (C["DRV_S_CPREVX"], _NC, "f"), "cprevy": (C["DRV_S_CPREVY"], _NC, "f"),
    "cgoalx": (C["DRV_S_CGOALX"], _NC, "f"), "cgoaly": (C["DRV_S_CGOALY"], _NC, "f"),
    "ctype": (C["DRV_S_CTYPE"], _NC, "i"), "cteam": (C["DRV_S_CTEAM"], _NC, "i"), "cflags": (C["DRV_S_CFLAGS"], _NC, "i"),
    "clane": (C["DRV_S_CLANE"], _NC, "i"),
    "ppx": (C["DRV_S_PPX"], _NPD, "f"), "ppy": (C["DRV_S_PPY"], _NPD, "f"), "pvx": (C["DRV_S_PVX"], _NPD, "f"),
    "pvy": (C["DRV_S_PVY"], _NPD, "f"), "pvbx": (C["DRV_S_PVBX"], _NPD, "f"), "pvby": (C["DRV_S_PVBY"], _NPD, "f"),
    "pmoving": (C["DRV_S_PMOVING"], _NPD, "f"), "pspeed": (C["DRV_S_PSPEED"], _NPD, "i"), "pflags": (C["DRV_S_PFLAGS"], _NPD, "i"),
    "opx": (C["DRV_S_OPX"], _NO, "f"), "opy": (C["DRV_S_OPY"], _NO, "f"),
    "elapsed": (C["DRV_S_ELAPSED"], 1, "i"), "episode": (C["DRV_S_EPISODE"], 1, "i"), "allFinished": (C["DRV_S_ALLFIN"], 1, "i"),
    "nPed": (C["DRV_S_NPED"], 1, "i"), "nObs": (C["DRV_S_NOBS"], 1, "i"), "stepped": (C["DRV_S_STEPPED"], 1, "i"),
    "episodeRewards": (C["DRV_S_EPR"], _NC, "f"), "episodePosRewards": (C["DRV_S_EPPR"], _NC, "f"),
    "nArb": (C["DRV_S_NARB"], 1, "i"), "arbKey": (C["DRV_S_ARBKEY"], _NDA, "i"), "arbCon": (C["DRV_S_ARBCON"], _NDA, "i"),
    "arbJn0": (C["DRV_S_ARBJN0"], _NDA, "f"), "arbJt0": (C["DRV_S_ARBJT0"], _NDA, "f"),
    "arbJn1": (C["DRV_S_ARBJN1"], _NDA, "f"), "arbJt1": (C["DRV_S_ARBJT1"], _NDA, "f"),
    "cpxLo": (C["DRV_S_CPXL"], _NC, "f"), "cpyLo": (C["DRV_S_CPYL"], _NC, "f"), "cangLo": (C["DRV_S_CANGL"], _NC, "f"),
    "ppxLo": (C["DRV_S_PPXL"], _NPD, "f"), "ppyLo": (C["DRV_S_PPYL"], _NPD, "f"),
}
DRV_SPLIT = {"cpx": "cpxLo", "cpy": "cpyLo", "ppx": "ppxLo", "ppy": "ppyLo"}
# float64 fields stored as two 32-bit words: high-word field -> low-word field
DRV_DOUBLE = {"cang": "cangLo"}
DRV_INT_EXTRA = [C["DRV_S_STEP_DIAG"]]


def int_mask(fields, n_words, extra=()):
    mask = [False] * n_words
    for off, cnt, kind in fields.values():
        if kind == "i":
            for k in range(off, off + cnt):
                mask[k] = True
    for k in extra:
        mask[k] = True
    return mask


def drv_int_mask():
    return int_mask(DRV_FIELDS, C["DRV_STATE_WORDS"], DRV_INT_EXTRA)
