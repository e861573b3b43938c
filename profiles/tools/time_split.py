"""Times the fused and the split layout on the C4 workload and SimpleDocking3d (1M envs, 40 steps after burn-in)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N = 1 << 20
gen = torch.Generator(device="cuda").manual_seed(1)
pool = [torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1 for _ in range(16)]
out = []
for name, kw in [("ObstaclesDocking3d", dict(layout="warp_rays", n_synthetic_spheres=3)),
                 ("ObstaclesDocking3d", dict(layout="split", n_synthetic_spheres=3, split_chunk_envs=1 << 20)),
                 ("ObstaclesDocking3d", dict(layout="pipeline", n_synthetic_spheres=3)),
                 ("SimpleDocking3d", dict(layout="warp_rays")),
                 ("SimpleDocking3d", dict(layout="split", split_chunk_envs=1 << 20)),
                 ("SimpleDocking3d", dict(layout="pipeline"))]:
    env = envs.SCENARIOS[name](cfg, num_envs=N, seed=0, **kw)
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
    out.append(f"{name[:9]}/{kw['layout'][:5]} {e0.elapsed_time(e1) / 40:.3f}")
    env.close()
print(os.environ.get("DOCKAUV_LIB", "default").split("libdockauv_")[-1], " | ".join(out))
