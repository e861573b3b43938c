"""Small runs of every layout / scenario for compute-sanitizer (memcheck, racecheck, initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
from gym_dockauv_b200.rollout import DeviceRolloutBuffer
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N = 1000          # not a multiple of any CTA size
gen = torch.Generator(device="cuda").manual_seed(0)
for name, kw in [("ObstaclesDocking3d", dict(n_synthetic_spheres=3)), ("ObstaclesCurrentDocking3d", dict(n_synthetic_spheres=8)),
                 ("SimpleDocking3d", {}), ("CapsuleCurrentDocking3d", {})]:
    for layout in ("pipeline", "warp_rays", "thread_per_env"):
        env = envs.SCENARIOS[name](cfg, num_envs=N, seed=1, layout=layout, **kw)
        env.reset()
        # short episodes so that resets happen inside the run
        env.t_steps += 995
        for k in range(12):
            env.step(torch.rand(N, env.n_actions, device="cuda", generator=gen) * 2 - 1)
        a = np.random.default_rng(0).uniform(-1, 1, (N, env.n_actions)).astype(np.float32)
        env.step_host(a)
        print(name, layout, env.get_stats()["episodes"])
        env.close()
env = envs.ObstaclesDocking3d(cfg, num_envs=N, seed=1, n_synthetic_spheres=3)
buf = DeviceRolloutBuffer(env, 8)
buf.reset_env()
buf.collect_open_loop()
buf.collect_open_loop()
buf.compute_returns_and_advantage(torch.zeros(N, device="cuda"))
torch.cuda.synchronize()
print("rollout ok", int(buf.dones.sum()))
