#!/usr/bin/env python
"""Record golden traces from the UNMODIFIED reference (run in the build container, where /root/reference
is mounted; the GPU box only ever sees the committed ``.npz`` files).

    python tests/golden/make_golden.py            # regenerates every tests/golden/*.npz

For each case the reference env (gym_dockauv/envs/docking3d.py, via tests/golden/ref_shims.py) is stepped
with a recorded action sequence and auto-reset on done, exactly as SB3's DummyVecEnv would drive it
(train.py:64-71).  Everything the parity tests need is dumped per step:

  per episode e : init_state[12], goal[3], heading_goal, capsules[Kc,7]=(bot,top,r), spheres[Ks,4],
                  current[5]=(V_c, alpha, beta, V_min, V_max), ep_len
  per step (e,t): action, state[12] (post-step), u[n_u] (low-passed command), state_dot[12] (post-step
                  RHS, auvsim.py:108), nu_c[6] (pre-step, docking3d.py:349), ray_dist[n_rays] (clamped,
                  sensor.py:113-118), obs f32[n_obs], reward, reward_arr[13], conditions[5], collision,
                  done, delta_d, delta_theta, delta_psi

Initial conditions are whatever the reference's own reset() produced (or, for the scripted event cases,
reset() followed by an explicit pose override that is recorded the same way), so the parity contract is
"inject the reference's post-reset initial conditions, then compare step trajectories" (SURVEY.md §9.7).
"""
import contextlib
import copy
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
from gym_dockauv.config.env_config import BASE_CONFIG  # noqa: E402
from gym_dockauv.envs import docking3d  # noqa: E402
from gym_dockauv.objects.shape import Sphere, Spheres  # noqa: E402
from gym_dockauv.objects.current import Current  # noqa: E402


def quiet_config(overrides=None):
    cfg = copy.deepcopy(BASE_CONFIG)
    cfg["verbose"] = 0
    cfg["log_level"] = 50
    cfg["interval_datastorage"] = 10 ** 9
    cfg["interval_episode_log"] = 10 ** 9
    cfg["save_path_folder"] = tempfile.mkdtemp(prefix="dockauv_golden_")
    for k, v in (overrides or {}).items():
        if isinstance(v, dict):
            cfg[k].update(v)
        else:
            cfg[k] = v
    return cfg


RADAR64 = {"alpha": 70 * np.pi / 180, "beta": 70 * np.pi / 180, "ray_per_deg": 10 * np.pi / 180}


# ----------------------------------------------------------------------------------------------------
# post-reset hooks (explicit, recorded overrides of the reference's initial conditions)
# ----------------------------------------------------------------------------------------------------
def hook_three_spheres(env, rng):
    """C4 workload: the stock 5 capsules plus 3 synthetic spheres."""
    sph = []
    for _ in range(3):
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        sph.append(Sphere(position=v * rng.uniform(4.0, 10.0), radius=1.0))
    env.spheres = Spheres(sph)
    env.obstacles = [*env.capsules, *env.spheres()]


def hook_near_pillar(env, rng):
    """Spawn 3.5 m from a random pillar/dock, pointing roughly at it -> collisions and real radar hits."""
    cap = env.capsules[rng.integers(len(env.capsules))]
    ang = rng.uniform(0, 2 * np.pi)
    centre = cap.position
    pos = np.array([centre[0] + 3.5 * np.cos(ang), centre[1] + 3.5 * np.sin(ang), rng.uniform(-1.5, 1.5)])
    env.auv.position = pos
    yaw = np.arctan2(centre[1] - pos[1], centre[0] - pos[0]) + rng.uniform(-0.3, 0.3)
    env.auv.attitude = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), yaw])


def hook_near_goal(env, rng):
    """Spawn 2.5 m from the goal, pointing at it -> Done-Goal_reached within a few dozen steps."""
    v = rng.normal(size=3)
    v[2] *= 0.2
    v /= np.linalg.norm(v)
    pos = env.goal_location + 2.5 * v
    if len(env.capsules) > 0:  # keep clear of the dock capsule
        r = np.hypot(pos[0], pos[1])
        if r < 2.6:
            pos[:2] *= 2.6 / max(r, 1e-6)
    env.auv.position = pos
    d = env.goal_location - pos
    env.auv.attitude = np.array([0.0, 0.0, np.arctan2(d[1], d[0]) + rng.uniform(-0.2, 0.2)])


def hook_near_edge(env, rng):
    """Spawn 19.5 m from the goal heading outwards -> Done-out_pos."""
    v = rng.normal(size=3)
    v[2] *= 0.3
    v /= np.linalg.norm(v)
    env.auv.position = env.goal_location + 19.5 * v
    env.auv.attitude = np.array([0.0, 0.0, np.arctan2(v[1], v[0])])


def hook_stochastic_current(env, rng):
    """white_noise_std > 0 (current.py:88-96); the recorded per-step V_c is fed to the CUDA path."""
    ang = (rng.random(2) - 0.5) * 2 * np.array([np.pi / 2, np.pi])
    env.current = Current(mu=0.01, V_min=0.2, V_max=1.0, Vc_init=0.5, alpha_init=ang[0], beta_init=ang[1],
                          white_noise_std=0.1, step_size=env.auv.step_size)
    env.nu_c = env.current(env.auv.attitude)


# ----------------------------------------------------------------------------------------------------
# policies (their outputs are recorded, so parity does not depend on how they were produced)
# ----------------------------------------------------------------------------------------------------
def policy_random(scale):
    scale = np.asarray(scale, dtype=np.float64)

    def f(env, rng, n_u):
        return rng.uniform(-1, 1, n_u) * scale[:n_u]
    return f


def policy_overdrive(env, rng, n_u):
    """Actions outside [-1, 1]: exercises the clip in unnormalize_input and the RAW action in the penalty."""
    return rng.uniform(-2.0, 2.0, n_u)


def policy_seek(env, rng, n_u):
    """Crude proportional steering towards the goal (BlueROV2 joystick layout) plus noise."""
    diff = env.goal_location - env.auv.position
    yaw_err = docking3d.geom.ssa(np.arctan2(diff[1], diff[0]) - env.auv.attitude[2])
    a = np.zeros(n_u)
    a[0] = 0.8 * np.cos(yaw_err)
    a[2] = np.clip(0.8 * diff[2], -1, 1)
    a[5] = np.clip(1.5 * yaw_err, -1, 1)
    return np.clip(a + rng.normal(0, 0.1, n_u), -1.2, 1.2)


def policy_forward(env, rng, n_u):
    a = rng.normal(0, 0.15, n_u)
    a[0] = 1.0
    return a


# ----------------------------------------------------------------------------------------------------
def record(name, env_id, n_steps, seed, overrides=None, policy=None, hook=None, action_dtype="f64",
           max_episodes=None, note=""):
    cfg = quiet_config(overrides)
    env = getattr(docking3d, env_id)(cfg)
    n_u = env.auv.u_bound.shape[0]
    n_rays = env.radar.n_rays
    n_obs = env.n_observations
    policy = policy or policy_random(np.ones(8))
    rng = np.random.default_rng(seed + 1)
    hook_rng = np.random.default_rng(seed + 2)

    episodes = []

    def start_episode(first):
        with contextlib.redirect_stdout(io.StringIO()):
            obs0 = env.reset(seed=seed) if first else env.reset()
        assert not np.any(obs0), "reset() must return the all-zero observation (docking3d.py:269,322)"
        if hook is not None:
            hook(env, hook_rng)
        caps = np.array([[*c.vec_bot, *c.vec_top, c.radius] for c in env.capsules], dtype=np.float64).reshape(-1, 7)
        sph = np.hstack([env.spheres.position, env.spheres.radius[:, None]]).reshape(-1, 4)
        cur = env.current
        episodes.append({
            "init_state": env.auv.state.copy(), "goal": np.array(env.goal_location, dtype=np.float64),
            "heading_goal": float(env.heading_goal_reached), "capsules": caps, "spheres": sph,
            "current": np.array([cur.V_c, cur.alpha, cur.beta, cur.V_min, cur.V_max], dtype=np.float64),
            "current_mu_sigma": np.array([cur.mu, cur.white_noise_std]),
            "steps": [],
        })

    start_episode(True)
    total = 0
    while total < n_steps:
        a = policy(env, rng, n_u)
        a = a.astype(np.float32) if action_dtype == "f32" else a.astype(np.float64)
        draws = []
        orig_normal = np.random.normal

        def spy_normal(*args, **kwargs):   # records the N(0, sigma) draw of Current.sim (current.py:88)
            w = orig_normal(*args, **kwargs)
            draws.append(float(w))
            return w
        np.random.normal = spy_normal
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                obs, reward, done, info = env.step(a)
        finally:
            np.random.normal = orig_normal
        assert obs.dtype == np.float32 and len(draws) == 1
        episodes[-1]["steps"].append({
            "action": a.astype(np.float64), "state": env.auv.state.copy(), "u": env.auv.u.copy(),
            "state_dot": env.auv._state_dot.copy(), "nu_c": env.nu_c.copy(), "v_c": float(env.current.V_c), "noise_w": draws[0],
            "ray_dist": env.radar.intersec_dist.copy(), "obs": obs.copy(), "reward": float(reward),
            "reward_arr": env.last_reward_arr.copy(), "conditions": np.array(env.conditions, dtype=np.uint8),
            "collision": bool(env.collision), "done": bool(done), "delta_d": float(env.delta_d),
            "delta_theta": float(env.delta_theta), "delta_psi": float(env.delta_psi),
            "t_steps": int(env.t_steps), "cum_reward": float(env.cumulative_reward),
        })
        total += 1
        if done:
            if max_episodes is not None and len(episodes) >= max_episodes:
                break
            if total < n_steps:
                start_episode(False)

    E = len(episodes)
    T = max(len(e["steps"]) for e in episodes)
    Kc = max(e["capsules"].shape[0] for e in episodes)
    Ks = max(e["spheres"].shape[0] for e in episodes)

    def per_ep(key, shape, dtype=np.float64):
        out = np.zeros((E,) + shape, dtype=dtype)
        for i, e in enumerate(episodes):
            v = np.asarray(e[key])
            out[(i,) + tuple(slice(0, s) for s in v.shape)] = v
        return out

    def per_step(key, shape, dtype=np.float64, fill=0):
        out = np.full((E, T) + shape, fill, dtype=dtype)
        for i, e in enumerate(episodes):
            for t, s in enumerate(e["steps"]):
                out[i, t] = s[key]
        return out

    data = {
        "init_state": per_ep("init_state", (12,)), "goal": per_ep("goal", (3,)),
        "heading_goal": per_ep("heading_goal", ()), "capsules": per_ep("capsules", (Kc, 7)),
        "n_capsules": np.array([e["capsules"].shape[0] for e in episodes], dtype=np.int32),
        "spheres": per_ep("spheres", (Ks, 4)),
        "n_spheres": np.array([e["spheres"].shape[0] for e in episodes], dtype=np.int32),
        "current": per_ep("current", (5,)), "current_mu_sigma": per_ep("current_mu_sigma", (2,)),
        "ep_len": np.array([len(e["steps"]) for e in episodes], dtype=np.int32),
        "action": per_step("action", (n_u,)), "state": per_step("state", (12,)), "u": per_step("u", (n_u,)),
        "state_dot": per_step("state_dot", (12,)), "nu_c": per_step("nu_c", (6,)), "v_c": per_step("v_c", ()), "noise_w": per_step("noise_w", ()),
        "ray_dist": per_step("ray_dist", (n_rays,)), "obs": per_step("obs", (n_obs,), np.float32),
        "reward": per_step("reward", ()), "reward_arr": per_step("reward_arr", (13,)),
        "conditions": per_step("conditions", (5,), np.uint8), "collision": per_step("collision", (), np.uint8),
        "done": per_step("done", (), np.uint8), "delta_d": per_step("delta_d", ()),
        "delta_theta": per_step("delta_theta", ()), "delta_psi": per_step("delta_psi", ()),
        "t_steps": per_step("t_steps", (), np.int32), "cum_reward": per_step("cum_reward", ()),
    }
    meta = {
        "name": name, "env_id": env_id, "vehicle": cfg["vehicle"], "seed": seed, "action_dtype": action_dtype,
        "n_u": n_u, "n_rays": n_rays, "n_obs": n_obs, "note": note,
        "overrides": json.loads(json.dumps(overrides or {}, default=float)),
        "config": json.loads(json.dumps({k: v for k, v in cfg.items() if k != "save_path_folder"},
                                        default=lambda o: np.asarray(o).tolist())),
        "radar_max_dist": float(env.radar.max_dist),
        "numpy": np.__version__, "total_steps": total,
    }
    data["meta"] = np.array(json.dumps(meta))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **data)
    conds = data["conditions"].reshape(-1, 5).sum(0)
    valid = np.arange(T)[None, :] < data["ep_len"][:, None]
    hits = int(((data["ray_dist"] < env.radar.max_dist) & valid[:, :, None]).sum())
    nan_steps = int(np.isnan(data["state"]).any(-1).sum())
    print(f"{name:38s} E={E:3d} T={T:5d} steps={total:5d} cond_counts={conds.tolist()} ray_hits={hits} "
          f"nan_steps={nan_steps} size={os.path.getsize(path) / 1024:.0f} KiB")
    return path


def record_vehicle_tables():
    """Physical parameters exactly as the reference's vehicle classes hold them after reading their XML
    (statespace.py:428-448), as plain JSON.  The oracle reads this file; a CPU test checks that the product's
    built-in vehicle tables carry the same numbers."""
    from gym_dockauv.objects.vehicles.BlueROV2 import BlueROV2
    from gym_dockauv.objects.vehicles.LAUV import LAUV
    tables = {}
    for key, v in (("BlueROV2", BlueROV2()), ("BlueROV2_direct", BlueROV2(control_mode="direct")),
                   ("LAUV", LAUV()),
                   ("BlueROV2_test", BlueROV2(os.path.join(ref_shims.REFERENCE_ROOT, "gym_dockauv", "tests",
                                                           "objects", "test_BlueROV2.xml")))):
        d = {k: val for k, val in vars(v).items() if isinstance(val, (int, float)) and not k.startswith("_")}
        d["u_bound"] = np.asarray(v.u_bound, dtype=float).tolist()
        if not key.startswith("LAUV"):
            d["B"] = np.asarray(v.B(None), dtype=float).tolist()
        d["safety_radius"] = float(v.safety_radius)
        d["lowpass_T1"] = float(v.lowpassfilter.T1)
        tables[key] = d
    path = os.path.join(HERE, "vehicles.json")
    with open(path, "w") as f:
        json.dump(tables, f, indent=1, sort_keys=True)
    print(f"vehicles.json                          {list(tables)}")


def record_unit_vectors():
    """Known answers of the reference's building blocks at hand-picked inputs: statespace matrices, RHS,
    one AUVSim.step, shape functions, radar ray table.  Complements the reference's own unit tests."""
    from gym_dockauv.objects.vehicles.BlueROV2 import BlueROV2
    from gym_dockauv.objects.vehicles.LAUV import LAUV
    from gym_dockauv.objects.sensor import Radar
    from gym_dockauv.objects import shape
    out = {}
    rng = np.random.default_rng(7)
    for veh_name, V in (("BlueROV2", BlueROV2), ("LAUV", LAUV)):
        v = V()
        v.step_size = 0.1
        n_u = v.u_bound.shape[0]
        nu = np.array([1.3, -0.4, 0.25, 0.3, -0.2, 0.5])
        eta = np.array([1.0, 2.0, 3.0, 0.2, -0.3, 1.0])
        out[f"{veh_name}_M_inv"] = v.M_inv
        out[f"{veh_name}_M_RB"] = v.M_RB
        out[f"{veh_name}_M_A"] = v.M_A
        out[f"{veh_name}_I_b"] = v.I_b
        out[f"{veh_name}_nu"] = nu
        out[f"{veh_name}_eta"] = eta
        out[f"{veh_name}_C"] = v.C(nu)
        out[f"{veh_name}_D"] = v.D(nu)
        out[f"{veh_name}_G"] = v.G(eta)
        out[f"{veh_name}_B"] = v.B(nu)
        out[f"{veh_name}_u_bound"] = v.u_bound.astype(np.float64)
        # RHS at random states
        states = rng.uniform(-1, 1, (16, 12)) * np.array([10, 10, 10, 1, 1, 3, 2, 1, 1, 1, 1, 1])
        us = rng.uniform(-1, 1, (16, n_u))
        nucs = np.hstack([rng.uniform(-0.5, 0.5, (16, 3)), np.zeros((16, 3))])
        rhs = np.zeros((16, 12))
        for i in range(16):
            v.u = v.unnormalize_input(us[i])
            rhs[i] = v.state_dot(0, states[i], nucs[i])
        out[f"{veh_name}_rhs_state"] = states
        out[f"{veh_name}_rhs_action"] = us
        out[f"{veh_name}_rhs_nu_c"] = nucs
        out[f"{veh_name}_rhs"] = rhs
    # SURVEY.md §8c G1/G2 (BlueROV2, h = 0.1)
    v = BlueROV2()
    v.step_size = 0.1
    for _ in range(100):
        v.step(np.array([1, 0, 0, -0.5, 0, 0]), np.zeros(6))
    out["G1_state"] = v.state.copy()
    out["G1_euler_dot"] = v.euler_dot.copy()
    v = BlueROV2()
    v.step_size = 0.1
    v.state = np.zeros(12)
    v.state[3:6] = [0.2, -0.3, 1.0]
    v.step(np.array([0.5, -0.25, 1.0, 0.1, -0.7, 0.3]), np.array([0.3, -0.1, 0.05, 0, 0, 0]))
    out["G2_state"] = v.state.copy()
    # BlueROV2 control_mode="direct" (8 thrusters, BlueROV2.py:53-72): the env never selects it (docking3d.py:78), so
    # it is pinned at the AUVSim level: 60 steps of vehicle.step() with random 8-dim actions and a constant current
    v = BlueROV2(control_mode="direct")
    v.step_size = 0.1
    acts = rng.uniform(-0.25, 0.25, (60, 8))   # full-scale random thrust blows the explicit RK4 up at h = 0.1
    nu_c = np.array([0.2, -0.1, 0.05, 0, 0, 0])
    v.state = np.zeros(12)
    v.state[3:6] = [0.1, -0.05, 0.7]
    traj, us = [], []
    for a in acts:
        v.step(a, nu_c)
        traj.append(v.state.copy())
        us.append(v.u.copy())
    out["direct_actions"], out["direct_states"], out["direct_u"] = acts, np.array(traj), np.array(us)
    out["direct_nu_c"] = nu_c
    # radar tables
    for tag, kw in (("stock", BASE_CONFIG["radar"]), ("r64", {**BASE_CONFIG["radar"], **RADAR64})):
        r = Radar(eta=np.zeros(6), **kw)
        out[f"radar_{tag}_rd_b"] = r.rd_b
        out[f"radar_{tag}_alpha"] = r.alpha
        out[f"radar_{tag}_beta"] = r.beta
        out[f"radar_{tag}_shape"] = np.array([r.n_vertical, r.n_horizontal, r.n_rays_reduced])
        d = rng.uniform(0.5, 10, r.n_rays)
        r.intersec_dist = d
        out[f"radar_{tag}_pool_in"] = d
        out[f"radar_{tag}_pool_out"] = r.intersec_dist_reduced
        out[f"radar_{tag}_oa"] = np.array(docking3d.Reward.obstacle_avoidance(
            theta_r=r.alpha, psi_r=r.beta, d_r=d, theta_max=r.alpha_max, psi_max=r.beta_max, d_max=r.max_dist,
            gamma_c=1, epsilon_c=0.001, epsilon_oa=0.01))
    # random ray / capsule / sphere cases
    n = 256
    l1 = rng.uniform(-8, 8, (n, 3))
    ld = rng.normal(size=(n, 3))
    cap1 = rng.uniform(-5, 5, (n, 3))
    cap2 = cap1 + rng.uniform(-6, 6, (n, 3))
    rad = rng.uniform(0.3, 2.0, n)
    # aim three quarters of the rays roughly at the capsule (body, end caps, near misses, and "behind" cases)
    aim = cap1 + (cap2 - cap1) * rng.uniform(-0.3, 1.3, (n, 1)) + rng.normal(0, 0.8, (n, 3))
    ld[: 3 * n // 4] = (aim - l1)[: 3 * n // 4] * rng.choice([1.0, 1.0, 1.0, -1.0], (3 * n // 4, 1))
    out["ray_l1"], out["ray_ld"], out["ray_cap1"], out["ray_cap2"], out["ray_rad"] = l1, ld, cap1, cap2, rad
    out["ray_capsule"] = np.array([
        shape.intersec_dist_line_capsule_vectorized(l1[i:i + 1], ld[i:i + 1], cap1[i], cap2[i], rad[i])[0]
        for i in range(n)])
    centres = rng.uniform(-6, 6, (n, 3, 3))
    centres[: n // 2, 0] = (l1 + ld / np.linalg.norm(ld, axis=1)[:, None] * rng.uniform(1, 9, (n, 1))
                            + rng.normal(0, 0.7, (n, 3)))[: n // 2]
    srad = rng.uniform(0.3, 2.0, (n, 3))
    out["ray_sph_c"], out["ray_sph_r"] = centres, srad
    out["ray_spheres"] = np.array([
        shape.intersec_dist_lines_spheres_vectorized(l1[i:i + 1], ld[i:i + 1], centres[i], srad[i])[0]
        for i in range(n)])
    out["dist_line_point"] = np.array([shape.dist_line_point(l1[i], cap1[i], cap2[i]) for i in range(n)])
    out["col_capsule"] = np.array([shape.collision_capsule_sphere(cap1[i], cap2[i], rad[i], l1[i], 1.0)
                                   for i in range(n)], dtype=np.uint8)
    out["col_spheres"] = np.array([shape.collision_sphere_spheres(l1[i], 1.0, centres[i], srad[i])
                                   for i in range(n)], dtype=np.uint8)
    x = np.concatenate([rng.uniform(-20, 20, 64), [np.pi, -np.pi, 3 * np.pi, 0.0, -4 / 3 * np.pi]])
    out["ssa_in"], out["ssa_out"] = x, docking3d.geom.ssa(x)
    path = os.path.join(HERE, "unit_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"unit_vectors                           {len(out)} arrays size={os.path.getsize(path) / 1024:.0f} KiB")


CASES = [
    # name, env_id, n_steps, seed, kwargs
    ("simple_bluerov2_f64", "SimpleDocking3d", 1200, 0, {}),
    ("simple_bluerov2_f32", "SimpleDocking3d", 600, 3, {"action_dtype": "f32"}),
    ("simple_bluerov2_overdrive", "SimpleDocking3d", 300, 4, {"policy": policy_overdrive}),
    ("simple_bluerov2_long", "SimpleDocking3d", 1001, 5,
     {"policy": policy_random([0.5, 0.5, 0.5, 0.1, 0.1, 0.3]), "max_episodes": 1,
      "note": "single 1001-step episode ending by Done-max_t (t_steps >= max_timesteps is tested pre-increment)"}),
    ("simplecurrent_bluerov2", "SimpleCurrentDocking3d", 400, 6, {}),
    ("capsule_bluerov2", "CapsuleDocking3d", 600, 7, {}),
    ("capsule_bluerov2_seek", "CapsuleDocking3d", 600, 8, {"policy": policy_seek}),
    ("capsulecurrent_bluerov2", "CapsuleCurrentDocking3d", 600, 9, {}),
    ("obstacles_bluerov2", "ObstaclesDocking3d", 1000, 0, {}),
    ("obstacles_bluerov2_pillar", "ObstaclesDocking3d", 800, 11, {"policy": policy_forward, "hook": hook_near_pillar}),
    ("obstacles_bluerov2_goal", "ObstaclesDocking3d", 400, 12, {"policy": policy_seek, "hook": hook_near_goal}),
    ("obstacles_bluerov2_edge", "ObstaclesDocking3d", 200, 13, {"policy": policy_forward, "hook": hook_near_edge}),
    ("obstaclesnocap_bluerov2", "ObstaclesNoCapDocking3d", 300, 14, {"policy": policy_seek}),
    ("obstaclescurrent_bluerov2", "ObstaclesCurrentDocking3d", 400, 15, {}),
    ("obstacles64_spheres_bluerov2", "ObstaclesDocking3d", 1000, 16,
     {"overrides": {"radar": RADAR64}, "hook": hook_three_spheres,
      "note": "BASELINE config C4: 64-ray radar, 5 capsules + 3 synthetic spheres"}),
    ("obstacles64_spheres_pillar_f32", "ObstaclesDocking3d", 600, 17,
     {"overrides": {"radar": RADAR64}, "action_dtype": "f32", "policy": policy_forward,
      "hook": lambda env, rng: (hook_three_spheres(env, rng), hook_near_pillar(env, rng))}),
    ("obstacles_bluerov2_rewardset2", "ObstaclesDocking3d", 300, 18,
     {"overrides": {"reward_set": 2}, "policy": policy_forward, "hook": hook_near_pillar}),
    ("obstacles_bluerov2_actionfactors", "ObstaclesDocking3d", 200, 19,
     {"overrides": {"action_reward_factors": [6.0, 3.0, 1.0, 0.5, 2.0, 4.0]}}),
    ("simple_bluerov2_stochcurrent", "SimpleDocking3d", 300, 20, {"hook": hook_stochastic_current}),
    ("capsulecurrent_lauv_h002", "CapsuleCurrentDocking3d", 1500, 21,
     {"overrides": {"vehicle": "LAUV", "t_step_size": 0.02}}),
    ("capsulecurrent_lauv_h01", "CapsuleCurrentDocking3d", 300, 22,
     {"overrides": {"vehicle": "LAUV"},
      "note": "stock h=0.1: explicit RK4 is unstable for LAUV; NaN/inf-aware comparison (SURVEY.md §8c)"}),
    ("capsule_lauv_h002_f32", "CapsuleDocking3d", 600, 23,
     {"overrides": {"vehicle": "LAUV", "t_step_size": 0.02}, "action_dtype": "f32"}),
]


if __name__ == "__main__":
    if not ref_shims.reference_available():
        sys.exit("reference not mounted at " + ref_shims.REFERENCE_ROOT)
    only = set(sys.argv[1:])
    np.seterr(all="ignore")
    if not only or "unit_vectors" in only:
        record_vehicle_tables()
        record_unit_vectors()
    for name, env_id, n_steps, seed, kw in CASES:
        if only and name not in only:
            continue
        record(name, env_id, n_steps, seed, **kw)
