"""Bit-exact replay of the reference's reset(seed) randomness against the recorded traces: every
episode's initial pose, goal, heading, obstacles and current, across auto-resets (the stream position depends on the
lengths of all earlier episodes because every step consumes one normal draw, current.py:88)."""
import numpy as np
import pytest

from gym_dockauv_b200.reference_reset import ReferenceResetStream
from tests.golden_utils import load_case, rel_err

# traces recorded without post-reset hooks (their initial conditions are the reference's own reset output)
PLAIN = ["simple_bluerov2_f64", "simple_bluerov2_f32", "simple_bluerov2_overdrive", "simplecurrent_bluerov2",
         "capsule_bluerov2", "capsule_bluerov2_seek", "capsulecurrent_bluerov2", "obstacles_bluerov2",
         "obstaclesnocap_bluerov2", "obstaclescurrent_bluerov2", "obstacles_bluerov2_actionfactors",
         "capsulecurrent_lauv_h002", "capsulecurrent_lauv_h01"]


@pytest.mark.parametrize("name", PLAIN)
def test_reset_stream_matches_reference(name):
    g = load_case(name)
    meta = g["meta"]
    s = ReferenceResetStream(meta["env_id"], meta["config"], meta["seed"])
    for e in range(len(g["ep_len"])):
        init = s.generate()
        assert np.array_equal(init["init_state"], g["init_state"][e]), (name, e)
        assert np.array_equal(init["goal"], g["goal"][e]) and init["heading_goal"] == g["heading_goal"][e]
        assert np.array_equal(init["current"], g["current"][e])
        k = g["n_capsules"][e]
        assert init["capsules"].shape[0] == k
        if k:
            assert rel_err(init["capsules"], g["capsules"][e][:k]) < 1e-15
        for _ in range(int(g["ep_len"][e])):
            s.consume_step()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["simple_bluerov2_f64", "obstacles_bluerov2", "capsulecurrent_bluerov2"])
def test_reference_seeded_env_reproduces_whole_trace(name):
    """reset(seed) + auto-resets on the GPU path, driven only by the seed and the recorded actions: every step of every
    episode of the trace matches the reference (flags bit-exact, state / reward 1e-9)."""
    import torch
    from gym_dockauv_b200 import envs
    from gym_dockauv_b200.reference_reset import ReferenceSeededEnv
    g = load_case(name)
    meta = g["meta"]
    has_cur = bool(np.any(g["current"][:, 0] != 0))
    env = envs.SCENARIOS[meta["env_id"]](meta["config"], num_envs=1, auto_reset=False, force_current=has_cur)
    v = ReferenceSeededEnv(env, [meta["seed"]])
    v.reset()
    f32 = meta["action_dtype"] == "f32"
    for e in range(len(g["ep_len"])):
        assert np.array_equal(v.last_init[0]["init_state"], g["init_state"][e])
        for t in range(int(g["ep_len"][e])):
            a = torch.as_tensor(g["action"][e, t][None].astype(np.float32 if f32 else np.float64), device=env.device)
            obs, reward, done, info = v.step(a)
            assert int(done[0]) == g["done"][e, t], (name, e, t)
            assert rel_err(reward.cpu().numpy()[0], g["reward"][e, t]) < 1e-9
            if not g["done"][e, t]:
                assert rel_err(env.state[:, 0].cpu().numpy(), g["state"][e, t]) < 1e-9
    env.close()
