"""CPU: the device-side InOutArranger equivalent (dynenv_b200/arranger.py) against the outputs of the reference's own
InOutArranger (tests/golden/gen_golden_arranger.py): flat per-type inputs in (time, env x player, row) order, counts,
padded [T, maxCount, players, F] embeddings and padding masks."""
import os

import numpy as np
import torch

from dynenv_b200.arranger import DeviceArranger
from helpers import GOLDEN


def test_arranger_matches_reference():
    g = dict(np.load(os.path.join(GOLDEN, "arranger.npz")))
    rows = [torch.from_numpy(g["rows%d" % i]) for i in range(6)]
    arr = DeviceArranger(rows, torch.from_numpy(g["counts"]))
    inputs, (counts, maxCount, objCounts) = arr.rearrange_inputs()
    assert maxCount == int(g["maxCount"])
    assert np.array_equal(counts.numpy(), g["ref_counts"]) and np.array_equal(objCounts.numpy(), g["objCounts"])
    for i in range(6):
        assert np.array_equal(inputs[i].numpy(), g["inputs%d" % i]), i
    padded, masks = arr.rearrange_outputs([x[:, :3] for x in inputs])
    assert tuple(padded.shape) == tuple(g["padded"].shape) and np.array_equal(padded.numpy(), g["padded"])
    assert np.array_equal(torch.stack(masks).numpy(), g["masks"])


def test_arranger_empty_type():
    """a type without any sighting: empty input array, no rows in the padded tensor"""
    E, T, P = 2, 1, 3
    rows = [torch.randn(E, T, P, 4, 2), torch.randn(E, T, P, 5, 3)]
    counts = torch.zeros((E, T, P, 2), dtype=torch.int32)
    counts[..., 1] = torch.tensor([[1, 0, 2], [3, 5, 0]]).view(E, T, P)
    arr = DeviceArranger(rows, counts)
    inputs, meta = arr.rearrange_inputs()
    assert inputs[0].shape == (0, 2) and inputs[1].shape == (11, 3) and meta[1] == 5
    padded, masks = arr.rearrange_outputs([None, inputs[1]], meta)
    assert padded.shape == (1, 5, 6, 3) and torch.equal(padded[0, :3, 3], rows[1][1, 0, 0, :3])
    assert masks[0].sum().item() == 6 * 5 - 11
