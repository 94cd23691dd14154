"""Multi-GPU use: environments are independent, so a job shards them by contiguous global env-id range,
one process per GPU, with NO collective on the step path (the reference's only parallelism is one OS
process per environment: DynEnv/utils/subproc_vec_env.py:86-100).  Global ids feed the RNG counter, so
results do not depend on how many GPUs the range is split over.  The one optional collective is the
reduction of episode statistics for logging."""
import torch


def shard_range(n_total, rank, world_size):
    """contiguous [begin, end) of global environment ids owned by `rank` (sizes differ by at most one)"""
    base, rem = divmod(int(n_total), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def episode_stats(out):
    """[sum episode_r, sum episode_p_r, sum episode_o_r, goals team 1, goals team 2, #episodes] of the
    environments that finished in the last step, as one float64 device tensor (Driving has no observation reward: 0)"""
    done = out.done.bool()
    f = done.unsqueeze(1).to(out.episode_r.dtype)
    o_r = getattr(out, "episode_o_r", None)
    return torch.stack([(out.episode_r * f).sum(), (out.episode_p_r * f).sum(),
                        (o_r * f).sum() if o_r is not None else out.episode_r.new_zeros(()),
                        (out.episode_g[:, 0] * done).sum().to(out.episode_r.dtype),
                        (out.episode_g[:, 1] * done).sum().to(out.episode_r.dtype),
                        done.sum().to(out.episode_r.dtype)]).double()


def all_reduce_episode_stats(stats, group=None):
    """sum of `stats` over all ranks (NCCL on GPUs, gloo in the CPU tests); a no-op without a process group"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
