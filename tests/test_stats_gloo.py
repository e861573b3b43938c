"""world_size-2 test (gloo, CPU) of the only collective on the path: the per-rollout statistics all-reduce, and of
the sharding of global env ids."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_dockauv_b200.params import N_STATS
    from gym_dockauv_b200.stats import reduce_stats, shard_env_range, summarize
    b, e = shard_env_range(1001, rank, world)
    local = torch.zeros(N_STATS, dtype=torch.float64)
    local[0] = 10 * (rank + 1)          # episodes
    local[1] = -100.0 * (rank + 1)      # sum_return
    local[2] = 500.0 * (rank + 1)       # sum_length
    local[7] = rank                     # collisions
    local[10] = float(e - b)            # env_steps
    before = local.clone()
    red = reduce_stats(local)
    assert torch.equal(local, before)       # the live accumulator is left alone; a second reduce counts nothing twice
    assert torch.equal(reduce_stats(local), red)
    s = summarize(red)
    out[rank] = (b, e, s["episodes"], s["mean_return"], s["mean_length"], s["env_steps"], s["collision_rate"])
    dist.destroy_process_group()


def test_stats_allreduce_two_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][:2] == (0, 501) and out[1][:2] == (501, 1001)
    for r in range(world):
        _, _, episodes, mean_ret, mean_len, steps, col = out[r]
        assert episodes == 30 and mean_ret == -10.0 and mean_len == 50.0 and steps == 1001
        assert abs(col - 1 / 30) < 1e-15
