import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N = 1 << 20
gen = torch.Generator(device="cuda").manual_seed(1)
pool = [torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1 for _ in range(16)]
for chunk in [1 << 20, 786432, 524288, 393216, 349568, 262144]:
    env = envs.ObstaclesDocking3d(cfg, num_envs=N, seed=0, layout="split", n_synthetic_spheres=3, split_chunk_envs=chunk)
    env.reset()
    for k in range(100):
        env.step(pool[k % 16])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(40):
        env.step(pool[k % 16])
    e1.record()
    torch.cuda.synchronize()
    print(chunk, f"{e0.elapsed_time(e1) / 40:.3f} ms")
    env.close()
