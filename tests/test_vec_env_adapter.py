"""The SB3-VecEnv shaped adapter: contract checks on CPU with a stub env, and on the GPU with the real env."""
import numpy as np
import pytest

from gym_dockauv_b200.vec_env import DockingVecEnv, LazyInfos


class _StubEnv:
    """Mimics the surface of gym_dockauv_b200.envs.BaseDocking3d that the adapter uses (host path)."""
    num_envs, n_observations, n_actions = 6, 36, 6
    observation_space = action_space = None
    device = "cpu"

    def __init__(self):
        import torch
        self.t = 0
        self.terminal_obs = torch.arange(6 * 36, dtype=torch.float32).reshape(6, 36)
        self.ep_return_out = torch.tensor([-1.0, -2.0, -3.0, -4.0, -5.0, -6.0], dtype=torch.float64)
        self.ep_len_out = torch.tensor([10, 20, 30, 40, 50, 60], dtype=torch.int32)

    def reset(self, seed=None):
        return np.zeros((6, 36), np.float32)

    def step_host(self, a):
        assert a.shape == (6, 6)
        self.t += 1
        done = np.array([0, 1, 0, 0, 1, 0], bool) if self.t == 2 else np.zeros(6, bool)
        obs = np.full((6, 36), 0.5, np.float32)
        obs[done] = 0
        return obs, np.arange(6, dtype=np.float64), done, {"cond_bits": np.array([0, 4, 0, 0, 17, 0], np.uint8)}

    def close(self):
        pass


def test_adapter_contract_with_stub():
    v = DockingVecEnv(_StubEnv())
    obs = v.reset()
    assert obs.shape == (6, 36) and obs.dtype == np.float32 and not obs.any()
    v.step_async(np.zeros((6, 6)))
    obs, rew, dones, infos = v.step_wait()
    assert rew.dtype == np.float32 and dones.dtype == bool and len(infos) == 6 and infos[3] == {}
    obs, rew, dones, infos = v.step(np.zeros((6, 6), np.float32))
    assert dones.tolist() == [False, True, False, False, True, False]
    assert isinstance(infos, LazyInfos) and infos.finished() == [1, 4]
    assert infos[0] == {} and infos[-1] == {}
    assert infos[1]["episode"]["r"] == -2.0 and infos[1]["episode"]["l"] == 20
    assert infos[4]["collision"] and infos[4]["goal_reached"] and infos[4]["conditions_true"] == [0, 4]
    assert np.array_equal(infos[4]["terminal_observation"], np.arange(4 * 36, 5 * 36, dtype=np.float32))
    assert not obs[1].any() and obs[0, 0] == 0.5            # reset observation for finished envs
    assert [bool(i) for i in infos] == [False, True, False, False, True, False]
    assert v.env_is_wrapped(object) == [False] * 6 and len(v.get_attr("n_actions")) == 6
    with pytest.raises(IndexError):
        infos[6]


@pytest.mark.gpu
def test_adapter_on_gpu_matches_env():
    import torch
    from gym_dockauv_b200 import envs
    from gym_dockauv_b200.config import BASE_CONFIG
    n = 4096
    env = envs.CapsuleDocking3d(dict(BASE_CONFIG), num_envs=n, seed=3)
    v = DockingVecEnv(env)
    assert v.reset().shape == (n, 36)
    rng = np.random.default_rng(0)
    episodes, ret = 0, 0.0
    for t in range(150):
        obs, rew, dones, infos = v.step(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
        assert obs.shape == (n, 36) and rew.shape == (n,) and dones.shape == (n,)
        for i in infos.finished():
            assert dones[i] and not obs[i].any()
            info = infos[i]
            assert info["terminal_observation"].shape == (36,) and info["terminal_observation"].any()
            assert 1 <= info["episode"]["l"] <= 1001
            episodes += 1
            ret += info["episode"]["r"]
    st = env.get_stats()
    assert episodes == st["episodes"] > 0
    assert abs(ret - st["sum_return"]) <= 1e-9 * abs(st["sum_return"])
    v.close()
