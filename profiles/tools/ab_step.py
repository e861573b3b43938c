"""One line per library variant: device-timed step of the C4 workload at 1M envs (40 steps after a burn-in), its
per-launch CUDA-event times, the C2 workload at 65,536 envs and SimpleDocking3d at 1M envs.

    DOCKAUV_LIB=.../libdockauv_<tag>.so python profiles/tools/ab_step.py [tag]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64

tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("DOCKAUV_LIB", "default"))
quick = "--quick" in sys.argv


def timed(env, pool, steps, burn):
    for k in range(burn):
        env.step(pool[k % len(pool)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        env.step(pool[k % len(pool)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


cfg = dict(BASE_CONFIG)
cfg["radar"] = dict(RADAR_64)
N = 1 << 20
gen = torch.Generator(device="cuda").manual_seed(1)
pool = [torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
env = envs.ObstaclesDocking3d(cfg, num_envs=N, seed=0, n_synthetic_spheres=3)
env.reset()
ms_c4 = timed(env, pool, 40, 140)
ms_c4b = timed(env, pool, 40, 0)
env.enable_timing(True)
per = []
for k in range(16):
    env.step(pool[k % 8])
    per.append(env.last_step_ms()[1])
env.enable_timing(False)
per = np.array(per).mean(axis=0)
env.close()
out = f"{tag:14s} C4 1M {min(ms_c4, ms_c4b):.4f} ms ({N / min(ms_c4, ms_c4b) * 1e3:.3e}/s) launches " + " ".join(f"{x * 1e3:.0f}" for x in per) + " us"
if not quick:
    env = envs.SimpleDocking3d(dict(BASE_CONFIG), num_envs=N, seed=0)
    env.reset()
    ms_s = timed(env, pool, 40, 60)
    env.close()
    n2 = 65536
    pool2 = [p[:n2].contiguous() for p in pool]
    env = envs.SimpleDocking3d(dict(BASE_CONFIG), num_envs=n2, seed=0)
    env.reset()
    ms_c2 = timed(env, pool2, 400, 100)
    env.close()
    out += f" | Simple 1M {ms_s:.4f} ms | C2 65536 {ms_c2 * 1e3:.1f} us ({n2 / ms_c2 * 1e3:.3e}/s)"
print(out, flush=True)
