"""GPU tests of the caller-side rows of SURVEY.md 8f: the stacked rollout call and the device-resident rollout buffer
(8f-1), GAE against a numpy restatement of SB3's RolloutBuffer.compute_returns_and_advantage, and the
EpisodeDataStorage-schema export (8f-3) against arrays recorded from the reference's own pickles."""
import os
import pickle

import numpy as np
import pytest

from tests.golden_utils import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "episode_storage_obstacles.npz")


def _cfg64():
    from gym_dockauv_b200.config import BASE_CONFIG, RADAR_64
    cfg = dict(BASE_CONFIG)
    cfg["radar"] = dict(RADAR_64)
    return cfg


@pytest.mark.parametrize("N,T", [(4096, 24), (1 << 19, 5)])     # the large batch is stepped in two parts on two streams
@pytest.mark.parametrize("use_graph", [False, True])
def test_rollout_equals_individual_steps(use_graph, N, T):
    import torch
    from gym_dockauv_b200 import envs
    kw = dict(num_envs=N, seed=3, n_synthetic_spheres=3)
    a = torch.rand(T, N, 6, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0)) * 2 - 1
    e1 = envs.ObstaclesDocking3d(_cfg64(), **kw)
    e1.reset()
    ref_obs, ref_rew, ref_done = [], [], []
    for t in range(T):
        o, r, d, _ = e1.step(a[t])
        ref_obs.append(o.clone()); ref_rew.append(r.clone()); ref_done.append(d.clone())
    e2 = envs.ObstaclesDocking3d(_cfg64(), **kw)
    e2.reset()
    obs = torch.zeros(T, N, e2.n_observations, device="cuda")
    rew = torch.zeros(T, N, dtype=torch.float64, device="cuda")
    done = torch.zeros(T, N, dtype=torch.uint8, device="cuda")
    ep_len = torch.full((T, N), -7, dtype=torch.int32, device="cuda")
    n0 = e2.launch_count()
    for rep in range(2 if use_graph else 1):     # second call replays the cached graph on a fresh reset
        if rep:
            e2.reset(seed=3)
        e2.rollout(a, obs, rew, done, ep_len_out=ep_len, use_graph=use_graph)
    torch.cuda.synchronize()
    assert e2.launch_count() - n0 >= T
    assert torch.equal(obs, torch.stack(ref_obs)) and torch.equal(rew, torch.stack(ref_rew))
    assert torch.equal(done, torch.stack(ref_done))
    assert torch.equal(e1.state, e2.state) and torch.equal(e1.t_steps, e2.t_steps)
    assert torch.equal(ep_len > 0, done.bool())
    e1.close(); e2.close()


def _gae_numpy(rewards, values, last_values, dones, gamma, lam):
    """stable-baselines3 1.5.0 RolloutBuffer.compute_returns_and_advantage, float32 like its buffers; the flag of
    "next observation starts an episode" is the done flag of the step (episode_starts[t + 1] = dones[t])."""
    T = rewards.shape[0]
    adv = np.zeros_like(values)
    last = np.zeros(values.shape[1], dtype=np.float32)
    for t in reversed(range(T)):
        nv = last_values if t == T - 1 else values[t + 1]
        nt = np.float32(1.0) - dones[t].astype(np.float32)
        delta = rewards[t] + np.float32(gamma) * nv * nt - values[t]
        last = delta + np.float32(gamma) * np.float32(lam) * nt * last
        adv[t] = last
    return adv, adv + values


@pytest.mark.parametrize("rdt", ["f64", "f32"])
def test_gae_matches_sb3_formula(rdt):
    import torch
    from gym_dockauv_b200 import envs
    from gym_dockauv_b200.config import BASE_CONFIG
    T, N = 37, 5000
    rng = np.random.default_rng(0)
    rewards = rng.normal(-1, 2, (T, N)).astype(np.float32)
    values = rng.normal(0, 5, (T, N)).astype(np.float32)
    last_values = rng.normal(0, 5, N).astype(np.float32)
    dones = (rng.random((T, N)) < 0.05).astype(np.uint8)
    env = envs.SimpleDocking3d(BASE_CONFIG, num_envs=8)
    dev = env.device
    adv = torch.zeros(T, N, device=dev)
    ret = torch.zeros(T, N, device=dev)
    r = torch.as_tensor(rewards.astype(np.float64) if rdt == "f64" else rewards, device=dev)
    env.gae(r, torch.as_tensor(values, device=dev), torch.as_tensor(last_values, device=dev),
            torch.as_tensor(dones, device=dev), 0.99, 0.95, adv, ret)
    a_ref, r_ref = _gae_numpy(rewards, values, last_values, dones, 0.99, 0.95)
    # float32 recurrences: FMA contraction on the device vs separate roundings in numpy
    assert np.max(np.abs(adv.cpu().numpy() - a_ref) / np.maximum(np.abs(a_ref), 1.0)) < 2e-5
    assert np.max(np.abs(ret.cpu().numpy() - r_ref) / np.maximum(np.abs(r_ref), 1.0)) < 2e-5
    env.close()


def test_device_rollout_buffer_closed_and_open_loop():
    import torch
    from gym_dockauv_b200 import envs
    from gym_dockauv_b200.rollout import DeviceRolloutBuffer
    N, T = 2048, 16
    env = envs.ObstaclesDocking3d(_cfg64(), num_envs=N, seed=1, n_synthetic_spheres=3)
    buf = DeviceRolloutBuffer(env, T, gamma=0.99, gae_lambda=0.95)
    buf.reset_env()
    w = torch.randn(env.n_observations, env.n_actions, device=env.device) * 0.5

    def policy(obs):
        a = torch.tanh(obs @ w)
        return a, obs.sum(1) * 0.1, -(a * a).sum(1)

    seen_done = 0
    for it in range(3):
        last = buf.collect(policy)
        assert torch.equal(last, buf.last_obs)
        # row t + 1 of the observations is what the env returned at step t: finished envs show the zero reset observation
        d = buf.dones.bool()
        assert torch.all(buf._obs[1:][d] == 0)
        assert torch.equal(buf._starts[1:], buf.dones)
        if it:
            assert torch.equal(buf.observations[0], prev_last) and torch.equal(buf.episode_starts[0], prev_done)
        prev_last, prev_done = buf.last_obs.clone(), buf.dones[-1].clone()
        buf.compute_returns_and_advantage(policy(last)[1])
        assert torch.isfinite(buf.advantages).all()
        info = buf.episode_infos()
        assert info["l"].numel() == int(d.sum()) and torch.all(info["l"] > 0)
        seen_done += int(d.sum())
        n = 0
        for mb in buf.get(batch_size=8192):
            n += mb.observations.shape[0]
            assert mb.observations.shape[1] == env.n_observations and mb.actions.shape[1] == env.n_actions
        assert n == T * N
    buf.collect_open_loop()
    assert torch.all(buf.actions.abs() <= 1)
    stats = env.get_stats()
    assert stats["env_steps"] == 4 * T * N
    env.close()


def test_episode_export_matches_reference_pickles(tmp_path):
    """Replays two episodes recorded from the reference with interval_datastorage = 1 and compares every array of the
    exported pickle with the reference's own EpisodeDataStorage pickle (same T + 2 row layout)."""
    import torch
    from gym_dockauv_b200 import envs
    from gym_dockauv_b200.config import BASE_CONFIG
    from gym_dockauv_b200.episode_export import EpisodeRecorder
    g = np.load(GOLD)
    for e in g["episodes"]:
        env = envs.ObstaclesDocking3d(BASE_CONFIG, num_envs=4, auto_reset=False, debug_outputs=True)
        rec = EpisodeRecorder(env, [2], str(tmp_path), title="t")
        env.reset()
        env.set_state(state=g[f"init_state_{e}"][None], goal=g[f"goal_{e}"][None], heading_goal=[g[f"heading_goal_{e}"]],
                      capsules=g[f"capsules_{e}"][None], u_prev=np.zeros((1, 6)), t_steps=[0], ep_return=[0.0],
                      env_ids=[2])
        rec._start_rows(0)
        rec.episode_no[2] = 1
        acts = g[f"action_{e}"]
        for t in range(len(acts)):
            a = torch.zeros(4, 6, dtype=torch.float64, device=env.device)
            a[2] = torch.as_tensor(acts[t], device=env.device)
            _, _, done, _ = rec.step(a)
            assert bool(done[2]) == (t == len(acts) - 1)
        (i, k, path), = [s for s in rec.saved if s[0] == 2]
        st = pickle.load(open(path, "rb"))
        assert sorted(st.keys()) == list(g["keys"]) and sorted(st["vehicle"].keys()) == list(g["vehicle_keys"])
        assert list(st["meta_data_reward"]) == list(g[f"meta_data_reward_{e}"])
        assert len(st["shapes"]) == int(g[f"n_shapes_{e}"])
        for key in ("states", "states_dot", "u"):
            ref = g[f"{key}_{e}"]
            assert st["vehicle"][key].shape == ref.shape, key
            assert rel_err(st["vehicle"][key], ref) < 1e-9, key
        for key in ("radar", "nu_c", "rewards", "cum_rewards"):
            ref = g[f"{key}_{e}"]
            assert st[key].shape == ref.shape, key
            assert rel_err(st[key], ref) < 1e-9, key
        assert np.max(np.abs(st["observation"] - g[f"observation_{e}"])) < 2e-7
        env.close()
