"""The counter-based (Philox) reset against the reference's own ``generate_environment`` (docking3d.py:687-703,
803-988): tests/golden/reset_draws.npz holds, for all seven scenarios and 72 (env id, episode) keys each, what the
UNMODIFIED reference's ``reset()`` produced when its global generator was made to serve exactly the uniforms of the
key's Philox stream (tests/golden/make_reset_golden.py).  Here the oracle (CPU) and the CUDA reset paths (GPU) must
reproduce those initial conditions from the key alone -- slot assignment, distributions and formulas included."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.golden_utils import GOLDEN_DIR, rel_err

TOL = 1e-13       # libm vs numpy sin / cos / arctan2 differ by an ulp or two


def _golden():
    d = np.load(os.path.join(GOLDEN_DIR, "reset_draws.npz"))
    return {k: d[k] for k in d.files}


def _keys(g):
    ids, eps = g["env_ids"], g["episodes"]
    return [(int(gid), int(ep)) for ep in eps for gid in ids]


SCENARIO_NAMES = ["SimpleDocking3d", "SimpleCurrentDocking3d", "CapsuleDocking3d", "CapsuleCurrentDocking3d",
                  "ObstaclesDocking3d", "ObstaclesNoCapDocking3d", "ObstaclesCurrentDocking3d"]


def test_golden_covers_all_scenarios():
    g = _golden()
    assert sorted(g["scenarios"].tolist()) == sorted(SCENARIO_NAMES)
    assert len(_keys(g)) == 72
    # the served uniforms really are in [0, 1) and differ between keys
    u = g["ObstaclesCurrentDocking3d_uniforms"]
    assert u.min() >= 0 and u.max() < 1 and len(np.unique(u[:, 0])) == 72


@pytest.mark.parametrize("scenario", SCENARIO_NAMES)
def test_oracle_reset_matches_reference(scenario):
    from gym_dockauv_b200.config import BASE_CONFIG
    from oracle import oracle as orc
    g = _golden()
    P = orc.make_params(dict(BASE_CONFIG))
    L = orc.lib()
    seed = int(g["seed"])
    for k, (gid, ep) in enumerate(_keys(g)):
        E = orc.OrcEnv()
        E.episode = ep
        L.orc_reset_env(C.byref(P), C.byref(E), orc.SCENARIOS[scenario], seed, gid)
        assert E.episode == ep + 1 and E.t_steps == 0
        assert rel_err(np.array(E.state), g[scenario + "_state"][k]) < TOL, (scenario, k)
        assert rel_err(np.array(E.goal), g[scenario + "_goal"][k]) < TOL
        assert rel_err(E.heading_goal, g[scenario + "_heading"][k]) < TOL
        assert rel_err(np.array(E.cur), g[scenario + "_current"][k]) < TOL
        caps = g[scenario + "_capsules"][k]
        assert E.n_caps == caps.shape[0]
        got = np.array([list(E.caps[c]) for c in range(E.n_caps)]).reshape(-1, 7)
        assert rel_err(got, caps) < TOL


def _gpu_envs(scenario, g):
    """One env instance per contiguous block of golden env ids."""
    from gym_dockauv_b200 import envs
    from gym_dockauv_b200.config import BASE_CONFIG
    ids = g["env_ids"].astype(np.uint64)
    blocks, start = [], 0
    for i in range(1, len(ids) + 1):
        if i == len(ids) or ids[i] != ids[i - 1] + 1:
            blocks.append((start, i))
            start = i
    out = []
    for b, e in blocks:
        env = envs.SCENARIOS[scenario](dict(BASE_CONFIG), num_envs=e - b, seed=int(g["seed"]), env_id0=int(ids[b]))
        out.append((b, e, env))
    return out


def _compare_gpu(env, g, scenario, rows, what):
    st = env.state.t().cpu().numpy()
    assert rel_err(st, g[scenario + "_state"][rows]) < TOL, (scenario, what)
    assert rel_err(env.goal.t().cpu().numpy(), g[scenario + "_goal"][rows]) < TOL, (scenario, what)
    assert rel_err(env.heading_goal.cpu().numpy(), g[scenario + "_heading"][rows]) < TOL, (scenario, what)
    assert rel_err(env.current.t().cpu().numpy(), g[scenario + "_current"][rows]) < TOL, (scenario, what)
    caps = g[scenario + "_capsules"][rows]
    if caps.shape[1]:
        got = env.capsules.t().cpu().numpy().reshape(len(rows), -1, 7)[:, :caps.shape[1]]
        assert rel_err(got, caps) < TOL, (scenario, what)
    assert not env.u_prev.any() and not env.t_steps.any() and not env.ep_return.any()


@pytest.mark.gpu
@pytest.mark.parametrize("scenario", SCENARIO_NAMES)
def test_cuda_reset_kernel_matches_reference(scenario):
    """dockauv_reset (reset_kernel): initial conditions of every golden key."""
    g = _golden()
    n_ids = len(g["env_ids"])
    for b, e, env in _gpu_envs(scenario, g):
        for j, ep in enumerate(g["episodes"].tolist()):
            env.episode.fill_(ep)
            env.reset()
            rows = np.arange(j * n_ids + b, j * n_ids + e)
            _compare_gpu(env, g, scenario, rows, ("reset_kernel", ep))
            assert (env.episode == ep + 1).all()
        env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["pipeline", "warp_rays", "thread_per_env"])
@pytest.mark.parametrize("scenario", SCENARIO_NAMES)
def test_cuda_auto_reset_matches_reference(scenario, layout):
    """The in-step auto-reset (the episode-end launch of the pipeline layout, in-kernel in the others): every env is driven into
    Done-max_t, and the state the step leaves behind must be the golden initial condition of the NEXT episode key."""
    import torch
    from gym_dockauv_b200 import envs
    from gym_dockauv_b200.config import BASE_CONFIG
    g = _golden()
    n_ids = len(g["env_ids"])
    ids = g["env_ids"].astype(np.uint64)
    for b, e in ((0, 12), (12, 24)):
        env = envs.SCENARIOS[scenario](dict(BASE_CONFIG), num_envs=e - b, seed=int(g["seed"]), env_id0=int(ids[b]),
                                       layout=layout)
        for j, ep in enumerate(g["episodes"].tolist()):
            env.episode.fill_(0)
            env.reset()                                   # any valid state
            env.episode.fill_(ep)                         # the key the auto-reset will draw from
            env.t_steps.fill_(env.max_timesteps)          # docking3d.py:612: t_steps >= max_timesteps -> done
            a = torch.zeros(e - b, env.n_actions, device=env.device)
            obs, reward, done, info = env.step(a)
            assert done.all() and not obs.any()
            rows = np.arange(j * n_ids + b, j * n_ids + e)
            _compare_gpu(env, g, scenario, rows, ("auto_reset", layout, ep))
            assert (env.episode == ep + 1).all()
        env.close()
