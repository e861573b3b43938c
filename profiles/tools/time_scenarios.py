"""Kernel time per 1M-env step for several scenario / layout combinations (CUDA events, 40 steps after burn-in)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N = 1 << 20
gen = torch.Generator(device="cuda").manual_seed(1)
pool = [torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1 for _ in range(16)]
for name, kw in [("SimpleDocking3d", dict(layout="thread_per_env")), ("SimpleDocking3d", dict(layout="warp_rays")),
                 ("CapsuleDocking3d", dict(layout="warp_rays")), ("ObstaclesDocking3d", dict(layout="warp_rays")),
                 ("ObstaclesDocking3d", dict(layout="warp_rays", n_synthetic_spheres=3)),
                 ("ObstaclesCurrentDocking3d", dict(layout="warp_rays", n_synthetic_spheres=3)),
                 ("ObstaclesDocking3d", dict(layout="pipeline", n_synthetic_spheres=3)),
                 ("SimpleDocking3d", dict(layout="pipeline")),
                 ("ObstaclesDocking3d", dict(layout="thread_per_env", n_synthetic_spheres=3))][:(int(os.environ.get("N_CASES", "99")))]:
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
    ms = e0.elapsed_time(e1) / 40
    print(f"{name:28s} {str(kw):60s} {ms:7.3f} ms/step  {N / ms * 1e3:.3e} env-steps/s")
    env.close()
