"""Device-side equivalent of the learner's InOutArranger (reference: DynEnv/models/models.py:127-274, use in README.md:228-269).

The reference's model consumes the ragged nested observation lists through pure-Python re-shuffles: per object type one
flat array of all sightings ordered by (time step, environment x player, row), the counts per (type, time, player),
and -- after the per-type embedding blocks -- one padded tensor [T, maxObjCnt, players, features] plus boolean padding
masks.  At tens of thousands of environments that Python would dominate the wall clock; here the same arrays are built
from the simulator's padded-with-counts records with a handful of tensor operations that stay on the device
(boolean-mask compaction, prefix sums, one scatter).  Works on any torch device (the CPU tests use it directly).

    arr = DeviceArranger(rows, counts)          # rows: list per type of [E, T, P, cap, feat]; counts [E, T, P, nTypes]
    inputs, (counts, maxCount, objCounts) = arr.rearrange_inputs()
    outs = [block(x) for block, x in zip(blocks, inputs)]
    padded, masks = arr.rearrange_outputs(outs)  # [T, maxCount, E*P, F], list of T masks [E*P, maxCount]
"""
import torch


def robocup_rows(out):
    """(rows per type, counts) of a RoboCupOutputs in the reference's type order (ball, robot, post, cross, line cross, line)"""
    return [out.ball, out.robot, out.post, out.cross, out.fcross, out.line], out.counts


def driving_rows(out):
    """Driving: types (car, obstacle, pedestrian, self, lane) in the order of the observation tuple
    ((cars, obstacles, pedestrians), (self, lanes)), one tick; PARTIAL feature widths"""
    E, N = out.counts.shape[:2]
    ones = torch.ones((E, N, 1), dtype=out.counts.dtype, device=out.counts.device)
    counts = torch.cat([out.counts[..., :3], ones, out.counts[..., 3:4]], -1)
    rows = [out.car, out.obs, out.ped, out.self_.unsqueeze(-2), out.lane[..., :4]]
    return [r.unsqueeze(1) for r in rows], counts.unsqueeze(1)


class DeviceArranger(object):
    def __init__(self, rows, counts):
        self.nTypes = len(rows)
        E, T, P = counts.shape[:3]
        self.nTime, self.nPlayers = T, E * P
        # [T, E*P, ...]: the reference chains the environments' agents per time step (models.py:221-223)
        self.rows = [r.permute(1, 0, 2, 3, 4).reshape(T, E * P, r.shape[3], r.shape[4]) for r in rows]
        self.counts = counts.permute(3, 1, 0, 2).reshape(self.nTypes, T, E * P).long()  # [type, time, player]
        self.valid = [torch.arange(r.shape[2], device=r.device)[None, None, :] < c[..., None] for r, c in zip(self.rows, self.counts)]

    def rearrange_inputs(self):
        """models.py:220-247: (inputs per type [n_i, feat_i], (counts [type, T, players], maxCount, objCounts [T, players]))"""
        inputs = [r[v] for r, v in zip(self.rows, self.valid)]  # row-major compaction = (time, player, row) order
        objCounts = self.counts.sum(0)
        maxCount = int(objCounts.max().item()) if objCounts.numel() else 0
        return inputs, (self.counts, maxCount, objCounts)

    def rearrange_outputs(self, outs, countArr=None):
        """models.py:249-274: embeddings per type [n_i, F] -> padded [T, maxCount, players, F] (per player: its rows of
        type 0, then type 1, ..., then zeros) and the padding masks (True where the slot is padding), one per time step"""
        counts, maxCount, objCounts = countArr if countArr is not None else self.rearrange_inputs()[1]
        first = next(o for o in outs if o is not None)
        F, dev = first.shape[1], first.device
        padded = torch.zeros((self.nTime, self.nPlayers, maxCount, F), dtype=first.dtype, device=dev)
        before = torch.zeros_like(objCounts)
        for i, o in enumerate(outs):
            if o is not None and o.shape[0]:
                t, q, k = self.valid[i].nonzero(as_tuple=True)
                padded[t, q, before[t, q] + k] = o
            if o is not None:
                before = before + counts[i]
        masks = [torch.arange(maxCount, device=dev)[None, :] >= objCounts[t][:, None] for t in range(self.nTime)]
        return padded.permute(0, 2, 1, 3), masks
