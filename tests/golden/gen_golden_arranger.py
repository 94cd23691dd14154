"""Generates tests/golden/arranger.npz with the REFERENCE's own InOutArranger (DynEnv/models/models.py:127-274): the classes
ObsMask, Indexer and InOutArranger are executed from the reference's source text (the module itself imports matplotlib /
h5py / gym, which this image lacks), on a synthetic ragged RoboCup-shaped observation batch.
Run here only: python tests/golden/gen_golden_arranger.py"""
import itertools
import os
import re
import sys
from typing import List

import numpy as np
import torch
import torch.nn.functional as F

SRC = "/root/reference/DynEnv/models/models.py"
OUT = os.path.dirname(os.path.abspath(__file__))
FEATS = [5, 7, 6, 6, 8, 5]
CAPS = [8, 12, 6, 6, 18, 11]


def load_classes():
    txt = open(SRC).read()
    a, b = txt.index("class ObsMask(object):"), txt.index("# Complete input layer")
    ns = {"torch": torch, "F": F, "np": np, "List": List, "itertools": itertools,
          "flatten": lambda l: [item for sublist in l for item in sublist]}  # DynEnv/utils/utils.py:17
    exec(compile(txt[a:b], SRC, "exec"), ns)
    return ns


if __name__ == "__main__":
    ns = load_classes()
    rng = np.random.default_rng(4)
    E, T, P = 3, 5, 4
    rows = [rng.normal(size=(E, T, P, c, f)).astype(np.float32) for c, f in zip(CAPS, FEATS)]
    counts = np.stack([rng.integers(0, min(c, 4) + 1, (E, T, P)) for c in CAPS], -1).astype(np.int32)
    counts[1, 2, 3] = 0  # a player without sightings
    # the reference's nested format: x[env][time][player] = list of per-type arrays
    x = [[[[rows[i][e, t, p, :counts[e, t, p, i]] if counts[e, t, p, i] else np.array([]) for i in range(6)] for p in range(P)]
          for t in range(T)] for e in range(E)]
    arr = ns["InOutArranger"](6, E * P, T)
    inputs, (cnts, maxCount, objCounts) = arr.rearrange_inputs(x)
    outs = [torch.tensor(o[:, :3]) if len(o) else None for o in inputs]  # stand-in embeddings: the first three features
    padded, masks = arr.rearrange_outputs(outs, (cnts, maxCount, objCounts), "cpu")
    rec = {"rows%d" % i: rows[i] for i in range(6)}
    rec.update({"inputs%d" % i: np.asarray(inputs[i], dtype=np.float32) for i in range(6)})
    rec.update(counts=counts, ref_counts=np.array(cnts), maxCount=np.array(maxCount), objCounts=objCounts.numpy(), padded=padded.numpy(),
               masks=np.stack([m.numpy() for m in masks]))
    np.savez_compressed(os.path.join(OUT, "arranger.npz"), **rec)
    print("maxCount", maxCount, "padded", tuple(padded.shape), "inputs", [tuple(np.asarray(i).shape) for i in inputs])
