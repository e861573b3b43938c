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


def reduce_stats(stats_tensor, group=None):
    """Sum of the statistics vectors of all ranks (torch.distributed all-reduce; NCCL for the device tensor returned by
    ``env.stats_tensor()``, gloo for a CPU tensor).  The input is NOT modified: ``env.stats_tensor()`` is the library's
    live accumulator, and reducing it in place would make every rank keep accumulating on top of the global sum (the
    next reduce would then count earlier episodes world_size times).  A copy is reduced and returned; with one process
    (or without an initialised process group) the copy is returned as is.  Call ``env.clear_stats()`` after the reduce
    for per-rollout statistics, or keep accumulating for running totals -- both are consistent."""
    import torch.distributed as dist
    if stats_tensor.numel() != N_STATS:
        raise ValueError(f"statistics vector must have {N_STATS} entries")
    out = stats_tensor.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


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
