"""Steps one scenario a few times (for ncu): python run_scenario.py <Scenario> <layout> <n_synth_spheres> <steps>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gym_dockauv_b200 import envs
from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
name, layout, nsph, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
chunk = int(sys.argv[5]) if len(sys.argv) > 5 else 0
cfg = dict(BASE_CONFIG); cfg["radar"] = dict(RADAR_64)
N = 1 << 20
env = envs.SCENARIOS[name](cfg, num_envs=N, seed=0, layout=layout, n_synthetic_spheres=nsph, split_chunk_envs=chunk)
env.reset()
env.enable_step_graph(False)      # plain launches: ncu's --launch-skip counts kernels
gen = torch.Generator(device="cuda").manual_seed(1)
pool = [torch.rand(N, 6, device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
for k in range(steps):
    env.step(pool[k % 8])
torch.cuda.synchronize()
print("ok")
