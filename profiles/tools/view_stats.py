"""In-view statistics of the C4 workload in steady state (needs a library built with -DDOCKAUV_VIEW_STATS)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N = 1 << 20
env = envs.ObstaclesDocking3d(cfg, num_envs=N, seed=0, n_synthetic_spheres=3, layout=sys.argv[1] if len(sys.argv) > 1 else "auto")
env.reset()
gen = torch.Generator(device="cuda").manual_seed(1)
for k in range(128):
    env.step(torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1)
env.clear_stats()
steps = 32
hits = 0.0
for k in range(steps):
    obs, _, _, _ = env.step(torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1)
    hits += float(((obs[:, 16:] < 1.0).any(1)).float().mean())
st = env.stats_tensor().cpu().numpy()
print(f"per env-step: in-view pairs {st[11] / (N * steps):.3f}, envs with a non-empty view {st[12] / (N * steps):.3f}, "
      f"envs with a pooled cell < 1 (real hit within range) {hits / steps:.3f}")
