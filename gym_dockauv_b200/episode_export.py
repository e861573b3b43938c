"""Trajectory export in the reference's ``EpisodeDataStorage`` pickle schema for selected envs of a
batch, so the reference's post-analysis tooling (``EpisodeDataStorage.load`` and its ``plot_*`` wrappers,
gym_dockauv/utils/datastorage.py:184-420) can read episodes simulated on the GPU.

Row layout, exactly as the reference produces it (datastorage.py:218-288, call sites docking3d.py:252-254, 363-364,
676-684):

* row 0 is written at ``reset()``: initial state, zero ``state_dot`` / ``u``, the reset ``nu_c``, zero rewards and the
  zero observation; its radar end points are those of the radar reset, which the reference performs while the vehicle
  still sits at the origin with zero attitude (docking3d.py:262,285 run before generate_environment, :311);
* one row per ``step()``, appended in the MIDDLE of the step (after the radar / collision update, before
  observe / reward): the post-step ``state``, ``state_dot``, ``u``, radar end points and this step's ``nu_c`` together
  with the rewards / observation of the PREVIOUS step;
* one closing row at the next ``reset()``: the last state again, now with the last step's rewards / observation.

An episode of T steps therefore has T + 2 rows.  The pickle is a dict with the reference's keys.  ``vehicle.object`` and
``shapes`` hold instances of the reference's own classes when ``gym_dockauv`` is importable in the exporting process;
otherwise plain records with the same attribute names (``position``, ``radius``, ``vec_top``, ``vec_bot``).

This is inspection tooling for a handful of envs (one small device-to-host read per step), not part of the step path.
"""
import datetime
import os
import pickle

import numpy as np
import torch

META_DATA_REWARD = ["Nav_delta_d", "Nav_delta_theta", "Nav_delta_psi", "Att_phi", "Att_theta", "Thetadot",
                    "obstacle_avoid", "action", "Done-Goal_reached", "Done-out_pos", "Done-out_att", "Done-max_t",
                    "Done-collision"]                       # docking3d.py:160-182
N_CONT_REWARDS = 8                                          # docking3d.py:153


class ShapeRecord:
    """Stand-in for gym_dockauv.objects.shape.Sphere / Capsule when the reference package is not importable."""

    def __init__(self, kind, position, radius, vec_top=None):
        self.kind = kind
        self.position = np.asarray(position, dtype=float)
        self.radius = float(radius)
        if vec_top is not None:
            self.vec_top = np.asarray(vec_top, dtype=float)
            self.vec_bot = 2 * self.position - self.vec_top        # shape.py:105-108

    def __repr__(self):
        return f"ShapeRecord({self.kind}, position={self.position}, radius={self.radius})"


class VehicleRecord:
    """Stand-in for the AUVSim instance stored under storage['vehicle']['object']."""

    def __init__(self, name, step_size, u_bound):
        self.name, self.step_size, self.u_bound = name, float(step_size), np.asarray(u_bound)


def _rzyx(phi, theta, psi):
    cphi, sphi, cth, sth, cpsi, spsi = np.cos(phi), np.sin(phi), np.cos(theta), np.sin(theta), np.cos(psi), np.sin(psi)
    return np.array([[cpsi * cth, -spsi * cphi + cpsi * sth * sphi, spsi * sphi + cpsi * cphi * sth],
                     [spsi * cth, cpsi * cphi + sphi * sth * spsi, -cpsi * sphi + sth * spsi * cphi],
                     [-sth, cth * sphi, cth * cphi]])


def meta_data_observation(n_rays_reduced):
    return [["delta_d", "delta_theta", "delta_psi"], ["u", "v", "w"], ["phi", "theta", "psi_sin", "psi_cos"],
            ["p", "q", "r"], ["u_c", "v_c", "w_c"], [f"ray_{i}" for i in range(n_rays_reduced)]]   # docking3d.py:128-135


class EpisodeRecorder:
    """Wraps a batched env built with ``auto_reset=False, debug_outputs=True`` and records the episodes of the envs
    in ``env_ids``; every finished episode is written to ``path_folder`` as
    ``<utc>__<title>__EPISODE_<k>_DATA_STORAGE.pkl`` and all finished envs are reset (``env.reset(mask=done)``)."""

    def __init__(self, env, env_ids, path_folder, title="", save=True, shape_module=None):
        """``shape_module``: optionally the reference's own ``gym_dockauv.objects.shape`` module, passed in by a caller
        that has the reference installed and wants its ``Capsule`` / ``Sphere`` classes inside the pickles (what
        ``plot_*`` of the reference draws); by default the files hold ``ShapeRecord`` stand-ins with the same
        attributes.  This package never imports the reference on its own."""
        if env.auto_reset or env.debug is None:
            raise ValueError("create the env with auto_reset=False and debug_outputs=True")
        self.shape_module = shape_module
        self.env, self.ids = env, [int(i) for i in env_ids]
        self.path_folder, self.title, self.save = path_folder, title, save
        self._sel = torch.as_tensor(self.ids, device=env.device)
        self.rd_b = np.ctypeslib.as_array(env._params.rd_b)[:3 * env.n_rays].reshape(env.n_rays, 3).copy()
        self.max_dist = float(env._params.radar_max_dist)
        self.episode_no = {i: 0 for i in self.ids}
        self.rows = {}
        self.saved = []      # (env id, episode number, path or storage dict)

    # ------------------------------------------------------------------ device -> host snapshot of the tracked envs
    def _snapshot(self):
        e, s = self.env, self._sel
        parts = [e.state[:, s], e.u_prev[:e.n_actions, s], e.debug["state_dot"][:, s], e.debug["nu_c"][:, s],
                 e.debug["ray_dist"][:, s], e.debug["reward_arr"][:, s]]
        flat = torch.cat([p.to(torch.float64) for p in parts], dim=0).cpu().numpy()
        obs = e.obs[s].cpu().numpy()
        out, o = [], 0
        for n in (12, e.n_actions, 12, 3, e.n_rays, 13):
            out.append(flat[o:o + n].T)
            o += n
        return (*out, obs)

    def _nu_c_reset(self, k):
        """nu_c right after reset() (docking3d.py:283): the spawned current at the initial attitude."""
        cur = self.env.current[:, self.ids[k]].cpu().numpy()
        att = self.env.state[3:6, self.ids[k]].cpu().numpy()
        vn = cur[0] * np.array([np.cos(cur[1]) * np.cos(cur[2]), np.sin(cur[2]), np.sin(cur[1]) * np.cos(cur[2])])
        return np.concatenate([_rzyx(*att).T.dot(vn), np.zeros(3)])

    def _end_points(self, state, dist):
        rd_n = self.rd_b.dot(_rzyx(*state[3:6]).T)
        rd_n /= np.linalg.norm(rd_n, axis=1)[:, None]                # sensor.py:96-102
        return state[0:3] + rd_n * dist[:, None]                     # sensor.py:120

    def _start_rows(self, k):
        e, i = self.env, self.ids[k]
        state = e.state[:, i].cpu().numpy().astype(float)
        self.episode_no[i] += 1
        self.rows[i] = dict(states=[state], states_dot=[np.zeros(12)], u=[np.zeros(e.n_actions)],
                            nu_c=[self._nu_c_reset(k)],
                            radar=[self._end_points(np.zeros(12), np.full(e.n_rays, self.max_dist))],
                            rewards=[np.zeros(13)], cum_rewards=[np.zeros(13)],
                            observation=[np.zeros(e.n_observations)],
                            goal=e.goal[:, i].cpu().numpy().astype(float),
                            capsules=e.capsules[:e.n_capsules * 7, i].cpu().numpy().astype(float).reshape(-1, 7),
                            spheres=e.spheres[:e.n_spheres * 4, i].cpu().numpy().astype(float).reshape(-1, 4))

    # ------------------------------------------------------------------ gym-like surface
    def reset(self, seed=None):
        obs = self.env.reset(seed=seed)
        for k in range(len(self.ids)):
            self._start_rows(k)
        return obs

    def step(self, actions):
        e = self.env
        obs, reward, done, info = e.step(actions)
        state, u, sdot, nu_c, dist, rarr, obs_sel = self._snapshot()
        d = done.bool()
        d_sel = d[self._sel].cpu().numpy()
        for k, i in enumerate(self.ids):
            r = self.rows[i]
            # mid-step row: this step's kinematics with the previous step's rewards / observation
            r["states"].append(state[k]); r["states_dot"].append(sdot[k]); r["u"].append(u[k])
            r["nu_c"].append(np.concatenate([nu_c[k], np.zeros(3)]))
            r["radar"].append(self._end_points(state[k], dist[k]))
            r["rewards"].append(r.get("_last_r", np.zeros(13))); r["cum_rewards"].append(r.get("_cum", np.zeros(13)))
            r["observation"].append(r.get("_last_obs", np.zeros(e.n_observations)))
            r["_last_r"] = rarr[k]
            r["_cum"] = r.get("_cum", np.zeros(13)) + rarr[k]
            r["_last_obs"] = obs_sel[k].astype(float)
            if d_sel[k]:
                self._finish(k)
        if bool(d.any()):
            e.reset(mask=done)
            for k, i in enumerate(self.ids):
                if d_sel[k]:
                    self._start_rows(k)
        return obs, reward, done, info

    def _finish(self, k):
        e, i = self.env, self.ids[k]
        r = self.rows[i]
        # closing row written by the next reset() (docking3d.py:252-254)
        for key in ("states", "states_dot", "u", "nu_c", "radar"):
            r[key].append(r[key][-1])
        r["rewards"].append(r["_last_r"]); r["cum_rewards"].append(r["_cum"]); r["observation"].append(r["_last_obs"])
        shape = self.shape_module
        if shape is not None:
            shapes = [shape.Capsule(position=(c[0:3] + c[3:6]) / 2, radius=c[6], vec_top=c[3:6]) for c in r["capsules"]]
            shapes += [shape.Sphere(position=s[0:3], radius=s[3]) for s in r["spheres"]]
            shapes.append(shape.Sphere(r["goal"], 0.15))
        else:
            shapes = [ShapeRecord("capsule", (c[0:3] + c[3:6]) / 2, c[6], vec_top=c[3:6]) for c in r["capsules"]]
            shapes += [ShapeRecord("sphere", s[0:3], s[3]) for s in r["spheres"]]
            shapes.append(ShapeRecord("sphere", r["goal"], 0.15))
        storage = {
            "vehicle": {"object": VehicleRecord(e.config.get("vehicle", "BlueROV2"), e.config["t_step_size"],
                                                e._meta["u_bound"]),
                        "states": np.array(r["states"]), "states_dot": np.array(r["states_dot"]), "u": np.array(r["u"])},
            "radar": np.array(r["radar"]), "nu_c": np.array(r["nu_c"]), "shapes": shapes, "title": self.title,
            "episode": self.episode_no[i], "step_size": e.config["t_step_size"],
            "cum_rewards": np.array(r["cum_rewards"]), "rewards": np.array(r["rewards"]),
            "meta_data_reward": list(META_DATA_REWARD), "n_cont_rewards": N_CONT_REWARDS,
            "observation": np.array(r["observation"]),
            "meta_data_observation": meta_data_observation(e.n_observations - 16),
        }
        if self.save:
            utc = datetime.datetime.now(datetime.timezone.utc).strftime('%Y_%m_%dT%H_%M_%S')
            if self.path_folder:
                os.makedirs(self.path_folder, exist_ok=True)
            name = os.path.join(self.path_folder, f"{utc}__{self.title}__ENV_{i}__EPISODE_{self.episode_no[i]}_DATA_STORAGE.pkl")
            with open(name, "wb") as f:
                pickle.dump(storage, f, pickle.HIGHEST_PROTOCOL)
            self.saved.append((i, self.episode_no[i], name))
        else:
            self.saved.append((i, self.episode_no[i], storage))
