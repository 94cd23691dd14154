"""Device-resident rollout buffers: the environment-side half of the reference's RolloutStorage
(DynEnv/models/storage.py:20-205,292-308) fed straight from the step kernels' result records.

The reference's training loop turns every step's rewards / dones into numpy, rebuilds agentFinished from
info['Full State'] in Python and copies all three back to the GPU (models/train.py:269-284); here one kernel per
step scatters them from the result record into [rollout_size, nEnvs * nAgents] device buffers, and the discounted
returns of storage._discount_rewards are one more kernel.  The policy-side buffers (actions, log-probs, values,
features) belong to the learner and are not part of this package.
"""
import ctypes as C

import torch

from . import _lib
from . import layout as L


class DeviceRolloutStorage(object):
    """rewards / agentFinished / dones of `rollout_size` steps of a BatchedRoboCup or BatchedDriving simulator.

    Field names follow models/storage.py:60-118: `rewards` float32 [rollout_size, num_all_players], `agentFinished` bool
    [rollout_size, num_all_players], `dones` uint8 [rollout_size, num_envs]; num_all_players = num_envs * num_players with the
    environment index major (the order of `torch.from_numpy(reward).view(-1)`, :157)."""

    def __init__(self, sim, rollout_size):
        self.sim = sim
        self._lib = _lib.lib()
        self.rollout_size = int(rollout_size)
        self.num_envs = sim.nEnvs
        self.num_players = sim.P if hasattr(sim, "P") else sim.N
        self.num_all_players = self.num_envs * self.num_players
        self.device = sim.device
        self._fin = torch.zeros((self.rollout_size, self.num_all_players), dtype=torch.uint8, device=self.device)
        self.agentFinished = self._fin.view(torch.bool)
        self.rewards = torch.zeros((self.rollout_size, self.num_all_players), dtype=torch.float32, device=self.device)
        self.dones = torch.zeros((self.rollout_size, self.num_envs), dtype=torch.uint8, device=self.device)
        self.episode_acc = torch.zeros((self.num_envs, L.DYNENV_EPISODE_ACC_WORDS), dtype=torch.float32, device=self.device)

    def reset_buffers(self):
        """models/storage.py:74-109 for the environment-side buffers (the episode statistics keep accumulating, like the
        reference's deques)"""
        self.rewards.zero_()
        self._fin.zero_()
        self.dones.zero_()

    after_update = reset_buffers  # models/storage.py:111-122 (the observation carry-over is the learner's)

    def insert(self, step, outputs=None):
        """storage.insert + storage.log_episode_rewards (models/storage.py:133-168,292-308) for the result record of the
        simulator's last step (`outputs`: default sim.out).  Device to device on the current stream."""
        o = outputs if outputs is not None else self.sim.out
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.dynenv_rollout_insert(self.sim._h, C.byref(o.c_struct), int(step), self.rollout_size,
                                                   self.rewards.data_ptr(), self._fin.data_ptr(), self.dones.data_ptr(),
                                                   self.episode_acc.data_ptr(), stream))

    def _discount_rewards(self, final_value, discount=0.99, respect_dones=False, out=None):
        """models/storage.py:170-205: r_discounted [rollout_size, num_all_players].  respect_dones=False is the reference
        (its done mask is all zeros, :194); True restarts the chain where an episode ended."""
        fv = final_value.to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        if fv.numel() != self.num_all_players:
            raise ValueError("final_value must have %d elements" % self.num_all_players)
        if out is None:
            out = torch.empty_like(self.rewards)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.dynenv_discount_rewards(self.device.index or 0, self.rewards.data_ptr(), fv.data_ptr(),
                                                     self.dones.data_ptr() if respect_dones else None, self.num_players,
                                                     out.data_ptr(), self.rollout_size, self.num_all_players, float(discount),
                                                     stream))
        return out

    discount_rewards = _discount_rewards

    def episode_stats(self):
        """What print_reward_stats (models/storage.py:310-330) reports: mean / min / max from the running sums, the median
        over the ring of the last 10 episodes' per-agent rewards of every environment (the reference's deque of
        10 * num_envs entries); one device reduction and one small copy to the host."""
        acc = self.episode_acc
        n = acc[:, 0].sum()
        has = acc[:, 0] > 0
        big = torch.finfo(torch.float32).max
        # ring entries that hold an episode: slot < min(episodes, ring size), agent < num_players
        R, RA, S0 = L.DYNENV_EPISODE_RING, L.DYNENV_EPISODE_RING_AGENTS, L.DYNENV_EPISODE_STAT_WORDS
        ring = acc[:, S0:S0 + R * RA].view(-1, R, RA)[:, :, :self.num_players]
        valid = (torch.arange(R, device=acc.device)[None, :] < acc[:, 0:1].clamp(max=R)).unsqueeze(-1).expand_as(ring)
        med = torch.where(valid, ring, torch.full_like(ring, float("nan"))).flatten().nanmedian()
        vals = torch.stack([n, acc[:, 1].sum(), acc[:, 2].sum(), acc[:, 3].sum(), acc[:, 4].sum(), acc[:, 5].sum(),
                            torch.where(has, acc[:, 6], torch.full_like(acc[:, 6], big)).min(),
                            torch.where(has, acc[:, 7], torch.full_like(acc[:, 7], -big)).max(), med]).cpu().tolist()
        k = max(vals[0], 1.0)
        return {"episodes": int(vals[0]), "mean_reward": vals[1] / k, "mean_pos_reward": vals[2] / k,
                "mean_obs_reward": vals[3] / k, "mean_goals": [vals[4] / k, vals[5] / k],
                "min_reward": vals[6] if vals[0] else 0.0, "max_reward": vals[7] if vals[0] else 0.0,
                "median_reward": vals[8] if vals[0] else 0.0}
