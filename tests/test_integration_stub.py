"""INTEGRATION.md section 3 shows the ctypes stub a maintainer of the reference would add (B200Backend: reset / step /
step_host over the C ABI, the N = 1 gym return shape).  This test executes exactly that code block, so the document
cannot drift from the library."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## 3."):]
    return re.search(r"```python\n(# gym_dockauv/envs/_b200\.py.*?)```", sec, re.S).group(1)


def test_stub_is_present_and_names_every_entry_point_it_uses():
    src = _stub_source()
    for sym in ("dockauv_create", "dockauv_bind", "dockauv_reset", "dockauv_step_host", "dockauv_set_seed", "dockauv_destroy",
                "dockauv_refresh_obstacles", "class B200Backend", "def reset", "def step_host", "def step("):
        assert sym in src, sym
    compile(src, "INTEGRATION.md:_b200.py", "exec")


@pytest.mark.gpu
def test_stub_runs_and_matches_the_packaged_env():
    import torch
    from gym_dockauv_b200 import _capi, envs
    from gym_dockauv_b200.config import BASE_CONFIG
    ns = {}
    exec(compile(_stub_source().replace('"libdockauv_b200.so"', repr(_capi.LIB_PATH)), "INTEGRATION.md:_b200.py", "exec"), ns)
    # ---- N = 1: the gym.Env contract
    b = ns["B200Backend"](dict(BASE_CONFIG), "ObstaclesDocking3d", num_envs=1, seed=5)
    env = envs.ObstaclesDocking3d(dict(BASE_CONFIG), num_envs=1, seed=5, auto_reset=False)
    assert not b.reset(seed=5).any() and not env.reset(seed=5).any()
    rng = np.random.default_rng(0)
    for t in range(30):
        a = rng.uniform(-1, 1, 6).astype(np.float32)
        obs, reward, done, info = b.step(a)
        o2, r2, d2, i2 = env.step(torch.as_tensor(a[None], device=env.device))
        assert obs.shape == (36,) and obs.dtype == np.float32 and isinstance(reward, float) and isinstance(done, bool)
        assert np.array_equal(obs, o2[0].cpu().numpy()) and reward == float(r2[0]) and done == bool(d2[0])
        ref = env.info_dict(0)
        for k in ("t_step", "t_total_steps", "cumulative_reward", "last_reward", "done", "conditions_true",
                  "conditions_true_info", "collision", "goal_reached"):
            assert info[k] == ref[k], k
        assert abs(info["delta_d"] - ref["delta_d"]) < 1e-12 and abs(info["simulation_time"] - ref["simulation_time"]) < 1e-12
    b.close()
    env.close()
    # ---- N > 1: VecEnv-style host step with auto-reset
    b = ns["B200Backend"](dict(BASE_CONFIG), "SimpleDocking3d", num_envs=512, seed=1)
    env = envs.SimpleDocking3d(dict(BASE_CONFIG), num_envs=512, seed=1)
    b.reset(seed=1)
    env.reset(seed=1)
    for t in range(5):
        a = rng.uniform(-1, 1, (512, 6))
        obs, reward, done, cond = b.step_host(a)
        o2, r2, d2, _ = env.step_host(a)
        assert np.array_equal(obs, o2) and np.array_equal(reward, r2) and np.array_equal(done, d2)
    b.close()
    env.close()
