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

`DeviceArranger` is the torch formulation (boolean-mask compaction: data-dependent shapes, hence host synchronisations);
`KernelArranger` is the product path: the same arrays from the observation RECORDS by the CUDA kernels behind
dynenv_arrange_inputs / dynenv_arrange_outputs (csrc/arrange.cuh) -- capacity-sized buffers, row totals and maxCount
left on the device, nothing returns to the host unless the caller asks for exact shapes (`compact`).
"""
import ctypes as C

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


class KernelArranger(object):
    """InOutArranger on the simulator's observation records, by kernels, without host synchronisation.

        arr = KernelArranger.for_robocup(sim)           # or .for_driving(sim): buffers allocated once, reused every step
        inputs, (counts, maxCount, objCounts) = arr.rearrange_inputs()      # after sim.step() / sim.reset()
        #   inputs[k]: [capacity_k, feat_k] -- the first arr.totals[k] rows are type k's sightings in (time, env x player,
        #   row) order, the rest is untouched; counts [type, T, E*P] int32; maxCount: 0-dim device tensor; objCounts [T, E*P]
        outs = [block(x) for block, x in zip(blocks, inputs)]               # embeddings of every (capacity) row
        padded, masks = arr.rearrange_outputs(outs, max_count=32)           # [T, 32, E*P, F], [T, E*P, 32] bool
        exact = arr.compact(inputs)                                          # ONE host read: inputs narrowed to their row totals
    """

    def __init__(self, record, layout, lib=None):
        from . import _lib
        self._lib = lib or _lib.lib()
        self._check = _lib.check
        self.record = record  # [E, T, P, rec_words] float32, CUDA
        E, T, P, W = record.shape
        self.E, self.T, self.P, self.Q = E, T, P, E * P
        self.layout = layout
        self.nTypes = layout.n_types
        dev = record.device
        self.device = dev
        n = T * E * P
        self.n_blocks = (n + 1023) // 1024
        self.counts = torch.zeros((self.nTypes, T, E * P), dtype=torch.int32, device=dev)
        self.objCounts = torch.zeros((T, E * P), dtype=torch.int32, device=dev)
        self.offsets = torch.zeros((self.nTypes, T, E * P), dtype=torch.int32, device=dev)
        self.block_sums = torch.zeros((self.nTypes, self.n_blocks), dtype=torch.int32, device=dev)
        self.totals = torch.zeros((self.nTypes + 1,), dtype=torch.int32, device=dev)  # rows per type, then maxCount
        self.flat = [torch.zeros((n * layout.cap[k], layout.feat[k]), dtype=torch.float32, device=dev) for k in range(self.nTypes)]
        self._flat_ptrs = (C.c_void_p * self.nTypes)(*[f.data_ptr() for f in self.flat])
        self.max_objects = sum(layout.cap[k] for k in range(self.nTypes))

    @staticmethod
    def _layout(rec_words, types):
        from ._lib import DynEnvArrangeLayout
        lay = DynEnvArrangeLayout()
        lay.n_types, lay.rec_words = len(types), rec_words
        for k, (off, cap, feat, stride, count_word) in enumerate(types):
            lay.off[k], lay.cap[k], lay.feat[k], lay.stride[k], lay.count_word[k] = off, cap, feat, stride, count_word
        return lay

    @classmethod
    def robocup_layout(cls):
        from . import layout as L
        t = [(L.RC_O_BALL, L.RC_CAP_BALL, L.RC_F_BALL), (L.RC_O_ROBOT, L.RC_CAP_ROBOT, L.RC_F_ROBOT), (L.RC_O_POST, L.RC_CAP_POST, L.RC_F_POST),
             (L.RC_O_CROSS, L.RC_CAP_CROSS, L.RC_F_CROSS), (L.RC_O_FCROSS, L.RC_CAP_FCROSS, L.RC_F_FCROSS), (L.RC_O_LINE, L.RC_CAP_LINE, L.RC_F_LINE)]
        return cls._layout(L.RC_OBS_WORDS, [(o, c, f, f, L.RC_O_COUNTS + k) for k, (o, c, f) in enumerate(t)])

    @classmethod
    def driving_layout(cls):
        """types (car, obstacle, pedestrian, self, lane) in the order of the observation tuple; PARTIAL feature widths"""
        from . import layout as L
        return cls._layout(L.DRV_OBS_WORDS, [(L.DRV_O_CAR, L.DRV_CAP_CAR, L.DRV_F_CAR, L.DRV_F_CAR, L.DRV_O_COUNTS + 0),
                                             (L.DRV_O_OBS, L.DRV_CAP_OBS, L.DRV_F_OBS, L.DRV_F_OBS, L.DRV_O_COUNTS + 1),
                                             (L.DRV_O_PED, L.DRV_CAP_PED, L.DRV_F_PED, L.DRV_F_PED, L.DRV_O_COUNTS + 2),
                                             (L.DRV_O_SELF, 1, L.DRV_F_SELF, L.DRV_F_SELF, -1),
                                             (L.DRV_O_LANE, L.DRV_CAP_LANE, L.DRV_F_LANE, L.DRV_F_LANE_FULL, L.DRV_O_COUNTS + 3)])

    @classmethod
    def for_robocup(cls, sim, outputs=None):
        o = outputs if outputs is not None else sim.out
        return cls(o.obs, cls.robocup_layout())

    @classmethod
    def for_driving(cls, sim, outputs=None):
        o = outputs if outputs is not None else sim.out
        return cls(o.record.unsqueeze(1), cls.driving_layout())  # one tick

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def rearrange_inputs(self):
        """models.py:220-247 on the records' current content: four kernels on the current stream, no host synchronisation"""
        self._check(self._lib.dynenv_arrange_inputs(self.device.index or 0, self.record.data_ptr(), self.E, self.T, self.P, C.byref(self.layout),
                                                    self._flat_ptrs, self.counts.data_ptr(), self.objCounts.data_ptr(), self.offsets.data_ptr(),
                                                    self.block_sums.data_ptr(), self.totals.data_ptr(), self._stream()))
        return self.flat, (self.counts, self.totals[self.nTypes], self.objCounts)

    def rearrange_outputs(self, outs, max_count=None):
        """models.py:249-274: per-type embeddings [rows >= totals[k], F] (None: the type has no block) -> padded [T, max_count,
        E*P, F] and the padding masks [T, E*P, max_count] (True = padding).  max_count defaults to the records' capacity."""
        first = next(o for o in outs if o is not None)
        F = int(first.shape[1])
        max_count = int(max_count) if max_count is not None else self.max_objects
        outs = [None if o is None else o.contiguous() for o in outs]
        ptrs = (C.c_void_p * self.nTypes)(*[None if o is None else o.data_ptr() for o in outs])
        padded = torch.empty((self.T, max_count, self.Q, F), dtype=torch.float32, device=self.device)
        mask = torch.empty((self.T, self.Q, max_count), dtype=torch.uint8, device=self.device)
        self._check(self._lib.dynenv_arrange_outputs(self.device.index or 0, ptrs, F, self.nTypes, self.E, self.T, self.P, self.counts.data_ptr(),
                                                     self.offsets.data_ptr(), self.block_sums.data_ptr(), self.objCounts.data_ptr(), max_count,
                                                     padded.data_ptr(), mask.data_ptr(), self._stream()))
        return padded, mask.view(torch.bool)

    def compact(self, inputs=None):
        """the flat arrays narrowed to their row totals (exactly what the reference's rearrange_inputs returns): one host read"""
        tot = self.totals.cpu().tolist()
        return [f[:tot[k]] for k, f in enumerate(inputs if inputs is not None else self.flat)]
