"""Sharding helpers and the per-rollout episode-statistics reduction (the only collective on this path).

Envs never interact, so a batch of N_total envs is split into contiguous blocks, one per rank / GPU, with no
per-step communication.  Each rank's kernel accumulates a 16-double statistics vector on its
device (DOCKAUV_STAT_* in include/dockauv.h); once per rollout the vectors are summed with ONE all-reduce
(NCCL over NVLink on GPUs; gloo in the CPU tests) -- this replaces the reference's host-side
FullDataStorage.update() bookkeeping (utils/datastorage.py:65-74).
"""
from .params import N_STATS, STAT_NAMES


def shard_env_range(n_total, rank, world_size):
    """Contiguous block [begin, end) of global env ids owned by `rank` (remainder spread over the first ranks).
    The global id keys the Philox reset stream, so results do not depend on how the batch is sharded."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_stats(stats_tensor, group=None, async_op=False):
    """Sum the statistics vectors of all ranks in place (torch.distributed all-reduce).  Works on the device
    tensor returned by env.stats_tensor() (NCCL) or on a CPU tensor (gloo).  Returns the work handle if async."""
    import torch.distributed as dist
    if stats_tensor.numel() != N_STATS:
        raise ValueError(f"statistics vector must have {N_STATS} entries")
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(stats_tensor, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def summarize(stats):
    """Derived per-rollout quantities from a (reduced) statistics vector or dict: what the reference logs per
    episode (docking3d.py:388-400) aggregated over the batch."""
    if not isinstance(stats, dict):
        vals = [float(x) for x in stats]
        stats = {k: vals[i] for i, k in enumerate(STAT_NAMES)}
    n = max(stats["episodes"], 1.0)
    return {
        "episodes": stats["episodes"],
        "env_steps": stats["env_steps"],
        "mean_return": stats["sum_return"] / n,
        "mean_length": stats["sum_length"] / n,
        "mean_final_delta_d": stats["sum_final_delta_d"] / n,
        "goal_reached_rate": stats["done_goal_reached"] / n,
        "collision_rate": stats["done_collision"] / n,
        "out_pos_rate": stats["done_out_pos"] / n,
        "out_att_rate": stats["done_out_att"] / n,
        "max_t_rate": stats["done_max_t"] / n,
        "nan_envs": stats["nan_envs"],
    }
