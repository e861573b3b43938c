"""FP32 variant vs the FP64 oracle: error growth over a short rollout (prints per-step max relative errors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
from oracle import oracle as orc
from tests.golden_utils import rel_err
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
n = 2048
env = envs.ObstaclesDocking3d(cfg, num_envs=n, seed=9, n_synthetic_spheres=3, precision="f32", env_id0=0)
env.reset()
bo = orc.BatchOracle(cfg, "ObstaclesDocking3d", n, seed=9, n_extra_spheres=3)
rng = np.random.default_rng(9)
alive = np.ones(n, bool)
for t in range(60):
    a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
    obs, rew, done, _ = env.step(torch.as_tensor(a, device=env.device))
    robs, rrew, rdone, _ = bo.step(a)
    d = done.cpu().numpy().astype(bool); rd = rdone.astype(bool)
    mism = (d != rd) & alive
    m = alive & ~d & ~rd
    st = env.state.t().cpu().numpy().astype(np.float64); rst = bo.field("state")
    print(t, "alive", int(alive.sum()), "flag mismatches", int(mism.sum()), "state %.2e" % rel_err(st[m], rst[m]),
          "reward %.2e" % rel_err(rew.cpu().numpy().astype(np.float64)[alive], rrew[alive], floor=1.0),
          "obs %.2e" % rel_err(obs.cpu().numpy()[m], robs[m]))
    alive &= ~(d | rd)      # after a reset the two sides may have diverged in episode phase; stop comparing
