"""Environment configuration accepted by the batched envs.

The key set and the default values are the reference's config contract (gym_dockauv/config/env_config.py:20-91),
so an ``env_config`` dict written for the reference can be passed unchanged.  Keys that do not influence the
arithmetic of the step path (title, log_level, verbose, interval_*, save_path_folder, *_goal_reached_tol other
than the distance, radius, w_t, radar freq) are accepted and ignored, as they are in the reference's step.
"""
import copy
import math
import os

# env ids of the reference registry (config/env_config.py:9-17) -> scenario class names in gym_dockauv_b200.envs
REGISTRATION_DICT = {
    f"{name}-v0": f"gym_dockauv_b200.envs:{name}"
    for name in ("SimpleDocking3d", "SimpleCurrentDocking3d", "CapsuleDocking3d", "CapsuleCurrentDocking3d",
                 "ObstaclesDocking3d", "ObstaclesCurrentDocking3d", "ObstaclesNoCapDocking3d")
}

_DEG = math.pi / 180.0


def _base_config():
    general = dict(config_name="DEFAULT_BASE_CONFIG", title="DEFAULT", log_level=20, verbose=1)
    episode = dict(max_timesteps=1000)
    simulation = dict(t_step_size=0.10, interval_datastorage=100, interval_episode_log=50,
                      save_path_folder=os.path.join(os.getcwd(), "logs"))
    goal = dict(max_dist_from_goal=20, max_attitude=60 / 180 * math.pi, dist_goal_reached_tol=0.5,
                velocity_goal_reached_tol=0.3, ang_rate_goal_reached_tol=20 * _DEG,
                attitude_goal_reached_tol=20 * _DEG)
    vehicle = dict(vehicle="BlueROV2", u_max=2.0, v_max=1.5, w_max=1.5, p_max=90 * _DEG, q_max=90 * _DEG,
                   r_max=120 * _DEG, radius=0.5)
    reward = dict(
        reward_set=1,
        reward_factors=dict(w_d=1.1, w_delta_psi=0.5, w_delta_theta=0.3, w_phi=0.3, w_theta=0.3, w_Thetadot=0.2,
                            w_t=0.05, w_oa=0.20, w_goal=400.0, w_deltad_max=-200.0, w_Theta_max=-200.0,
                            w_t_max=-100.0, w_col=-300.0),
        action_reward_factors=6.0,
    )
    radar = dict(radar=dict(freq=1, alpha=60 * _DEG, beta=80 * _DEG, ray_per_deg=10 * _DEG, max_dist=10,
                            blocksize_reduce=2))
    cfg = {}
    for part in (general, episode, simulation, goal, vehicle, reward, radar):
        cfg.update(part)
    return cfg


BASE_CONFIG = _base_config()


def _derived(title, folder, **over):
    cfg = copy.deepcopy(BASE_CONFIG)
    cfg["title"] = title
    cfg["save_path_folder"] = os.path.join(os.getcwd(), folder)
    cfg.update(over)
    return cfg


TRAIN_CONFIG = _derived("Training Run", "logs")
PREDICT_CONFIG = _derived("Prediction Run", "predict_logs", interval_datastorage=1, interval_episode_log=1)
MANUAL_CONFIG = _derived("Manual Run", "manual_logs", interval_datastorage=1, interval_episode_log=1)

# radar of the BASELINE headline workload: 8 x 8 = 64 rays
RADAR_64 = dict(freq=1, alpha=70 * _DEG, beta=70 * _DEG, ray_per_deg=10 * _DEG, max_dist=10, blocksize_reduce=2)

_REQUIRED = ("t_step_size", "max_timesteps", "max_dist_from_goal", "max_attitude", "dist_goal_reached_tol",
             "vehicle", "u_max", "v_max", "w_max", "p_max", "q_max", "r_max", "reward_set", "reward_factors",
             "action_reward_factors", "radar")
_REQUIRED_REWARD = ("w_d", "w_delta_psi", "w_delta_theta", "w_phi", "w_theta", "w_Thetadot", "w_oa", "w_goal",
                    "w_deltad_max", "w_Theta_max", "w_t_max", "w_col")


def validate(env_config):
    """KeyError with the missing key, like the reference's plain dict lookups (docking3d.py:52-201)."""
    for k in _REQUIRED:
        if k not in env_config:
            raise KeyError(k)
    for k in _REQUIRED_REWARD:
        if k not in env_config["reward_factors"]:
            raise KeyError(k)
    for k in ("alpha", "beta", "ray_per_deg", "max_dist"):
        if k not in env_config["radar"]:
            raise KeyError(k)
    if env_config["reward_set"] not in (1, 2):
        raise ValueError("reward_set must be 1 or 2")
    return env_config
