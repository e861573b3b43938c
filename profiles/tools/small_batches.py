import os, sys
sys.path.insert(0, os.getcwd())
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
tag = sys.argv[1]
out = []
for n in (65536, 131072, 262144, 524288):
    gen = torch.Generator(device="cuda").manual_seed(1)
    pool = [torch.rand(n, 6, device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
    env = envs.ObstaclesDocking3d(cfg, num_envs=n, seed=0, n_synthetic_spheres=3)
    env.reset()
    for k in range(150): env.step(pool[k % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(200): env.step(pool[k % 8])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    out.append(f"{n}: {ms*1e3:.1f} us ({n/ms*1e3:.3e}/s)")
    env.close()
print(tag, " | ".join(out), flush=True)
