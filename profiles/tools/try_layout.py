"""python try_layout.py <layout> <n_envs> <steps>: steps the C4 workload and prints ms/step (CUDA_LAUNCH_BLOCKING=1 to localise faults)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
layout, N, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
env = envs.ObstaclesDocking3d(cfg, num_envs=N, seed=0, layout=layout, n_synthetic_spheres=3, split_chunk_envs=chunk)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(1)
pool = [torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
for k in range(steps):
    env.step(pool[k % 8])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(40):
    env.step(pool[k % 8])
e1.record()
torch.cuda.synchronize()
print(layout, N, chunk, f"{e0.elapsed_time(e1) / 40:.3f} ms/step", env.get_stats())
