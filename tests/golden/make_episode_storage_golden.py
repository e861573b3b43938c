#!/usr/bin/env python
"""Records the arrays of the reference's own EpisodeDataStorage pickles (utils/datastorage.py:184-330) for the first
episodes of one env, driven with a recorded action sequence:  python tests/golden/make_episode_storage_golden.py

Output: tests/golden/episode_storage_obstacles.npz with, per episode e, the arrays of the pickle (states, states_dot, u,
nu_c, radar, rewards, cum_rewards, observation), the initial conditions needed to replay it (init_state, goal,
heading_goal, capsules) and the actions.  Only arrays are kept: the pickled reference objects cannot travel."""
import contextlib
import glob
import io
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shims, imports the reference)

cfg = mg.quiet_config({"interval_datastorage": 1})
env = mg.docking3d.ObstaclesDocking3d(cfg)
rng = np.random.default_rng(11)
out = {}
n_ep = 3
with contextlib.redirect_stdout(io.StringIO()):
    env.reset(seed=5)
for e in range(n_ep):
    caps = np.array([[*c.vec_bot, *c.vec_top, c.radius] for c in env.capsules])
    out[f"init_state_{e}"] = env.auv.state.copy()
    out[f"goal_{e}"] = np.array(env.goal_location, dtype=float)
    out[f"heading_goal_{e}"] = float(env.heading_goal_reached)
    out[f"capsules_{e}"] = caps
    acts = []
    done = False
    while not done:
        a = rng.uniform(-1, 1, 6)
        acts.append(a)
        with contextlib.redirect_stdout(io.StringIO()):
            _, _, done, _ = env.step(a)
    out[f"action_{e}"] = np.array(acts)
    with contextlib.redirect_stdout(io.StringIO()):
        env.reset()          # writes the closing row and saves the pickle (docking3d.py:252-254)
files = sorted(glob.glob(os.path.join(cfg["save_path_folder"], "*EPISODE_*_DATA_STORAGE.pkl")),
               key=lambda f: int(f.split("EPISODE_")[1].split("_")[0]))
assert len(files) >= n_ep, files
for e, f in enumerate(files[:n_ep]):
    st = pickle.load(open(f, "rb"))
    assert st["episode"] == e + 1
    for k in ("states", "states_dot", "u"):
        out[f"{k}_{e}"] = np.asarray(st["vehicle"][k])
    for k in ("radar", "rewards", "cum_rewards", "observation"):
        out[f"{k}_{e}"] = np.asarray(st[k])
    out[f"nu_c_{e}"] = np.asarray(st["nu_c"][:])   # save() leaves this one an ArrayList (datastorage.py:311-323)
    out[f"meta_data_reward_{e}"] = np.array(st["meta_data_reward"])
    out[f"n_shapes_{e}"] = len(st["shapes"])
out = {k: v for k, v in out.items() if not k.endswith("_1")}     # the 557-step middle episode is dropped (fixture size)
out["episodes"] = np.array([0, 2])
out["keys"] = np.array(sorted(st.keys()))
out["vehicle_keys"] = np.array(sorted(st["vehicle"].keys()))
np.savez_compressed(os.path.join(HERE, "episode_storage_obstacles.npz"), **out)
print("episodes", [out[f"states_{e}"].shape for e in out["episodes"]])
