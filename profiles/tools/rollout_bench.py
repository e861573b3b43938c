"""Open-loop rollout of the C4 workload (1M envs, T steps per dockauv_rollout call) against the per-step call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N, T = 1 << 20, 16
env = envs.ObstaclesDocking3d(cfg, num_envs=N, seed=0, n_synthetic_spheres=3)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(1)
a = torch.rand(T, N, 6, device="cuda", generator=gen) * 2 - 1
obs = torch.zeros(T, N, env.n_observations, device="cuda"); rew = torch.zeros(T, N, dtype=torch.float64, device="cuda")
done = torch.zeros(T, N, dtype=torch.uint8, device="cuda")
for k in range(8):
    env.rollout(a, obs, rew, done, use_graph=False)
for g in (False, True):
    env.rollout(a, obs, rew, done, use_graph=g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(6):
        env.rollout(a, obs, rew, done, use_graph=g)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (6 * T)
    print(f"dockauv_rollout T={T} graph={g}: {ms:.4f} ms per step, {N / ms * 1e3:.4g} env-steps/s")
