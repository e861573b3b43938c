"""Throughput of the other BASELINE configs (they are parity-test cases, not bench lines): C2, C3, FP32 C4, and the
stacked-rollout call with / without CUDA-graph replay at a small batch."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
gen = torch.Generator(device="cuda").manual_seed(1)
# C2: SimpleDocking3d, BlueROV2, 65,536 envs
N = 65536
env = envs.SimpleDocking3d(BASE_CONFIG, num_envs=N, seed=0)
env.reset()
pool = [torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
for k in range(100):
    env.step(pool[k % 8])
k = [0]
def one():
    env.step(pool[k[0] % 8]); k[0] += 1
ms = timed(one, 400)
out["C2 step loop (python, one call per step)"] = dict(ms_per_step=ms, env_steps_per_s=N / ms * 1e3)
T = 128
a = torch.rand(T, N, 6, device="cuda", generator=gen) * 2 - 1
obs = torch.zeros(T, N, env.n_observations, device="cuda"); rew = torch.zeros(T, N, dtype=torch.float64, device="cuda")
done = torch.zeros(T, N, dtype=torch.uint8, device="cuda")
for g in (False, True):
    env.rollout(a, obs, rew, done, use_graph=g)
    ms = timed(lambda: env.rollout(a, obs, rew, done, use_graph=g), 5) / T
    out[f"C2 dockauv_rollout T=128 graph={g}"] = dict(ms_per_step=ms, env_steps_per_s=N / ms * 1e3)
env.close()
# C3: CapsuleCurrentDocking3d, LAUV with ocean current, 262,144 envs (stock h = 0.1 and the stable h = 0.02)
for h in (0.1, 0.02):
    cfg = dict(BASE_CONFIG); cfg["vehicle"] = "LAUV"; cfg["t_step_size"] = h
    N = 262144
    env = envs.CapsuleCurrentDocking3d(cfg, num_envs=N, seed=0)
    env.reset()
    pool = [torch.rand(N, 3, device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
    for i in range(100):
        env.step(pool[i % 8])
    k = [0]
    def one3():
        env.step(pool[k[0] % 8]); k[0] += 1
    ms = timed(one3, 200)
    out[f"C3 LAUV h={h}"] = dict(ms_per_step=ms, env_steps_per_s=N / ms * 1e3, nan_envs=env.get_stats()["nan_envs"])
    env.close()
# C4 in FP32
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N = 1 << 20
env = envs.ObstaclesDocking3d(cfg, num_envs=N, seed=0, n_synthetic_spheres=3, precision="f32")
env.reset()
pool = [torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
for i in range(128):
    env.step(pool[i % 8])
k = [0]
def one4():
    env.step(pool[k[0] % 8]); k[0] += 1
ms = timed(one4, 100)
out["C4 FP32 handles"] = dict(ms_per_step=ms, env_steps_per_s=N / ms * 1e3)
env.close()
print(json.dumps(out, indent=1))
