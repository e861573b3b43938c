"""Fraction of envs of the C4 workload (steady state) in which at least one ray really hits within max_dist."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N = 1 << 17
env = envs.ObstaclesDocking3d(cfg, num_envs=N, seed=0, n_synthetic_spheres=3, debug_outputs=True)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(1)
acc, n = 0.0, 0
for k in range(160):
    env.step(torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1)
    if k >= 128:
        acc += float((env.debug["ray_dist"] < 10.0).any(0).float().mean()); n += 1
print(f"envs with at least one ray hit inside max_dist: {acc / n:.4f}")
