"""End-to-end step through host buffers (dockauv_step_host) at 1M envs: ms per step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N = 1 << 20
env = envs.ObstaclesDocking3d(cfg, num_envs=N, seed=0, n_synthetic_spheres=3)
env.reset()
host = [torch.empty(N, 6, dtype=torch.float32).pin_memory() for _ in range(2)]
for h in host:
    h.uniform_(-1, 1)
a = [h.numpy() for h in host]
for k in range(6):
    env.step_host(a[k % 2])
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(20):
    env.step_host(a[k % 2])
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 20 * 1e3
print(f"{ms:.3f} ms per step, {N / ms * 1e3:.4g} env-steps/s")
