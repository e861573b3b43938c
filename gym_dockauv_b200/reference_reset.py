"""Bit-exact replay of the reference's reset randomness, for small batches.

The reference draws every initial condition from the GLOBAL legacy numpy generator (MT19937): ``reset(seed)`` calls
``np.random.seed(seed)`` (docking3d.py:296-298), each ``generate_environment`` then consumes uniforms in a fixed
order (docking3d.py:687-703, 803-988), and every ``step`` consumes one normal draw in ``Current.sim`` even when the
noise is switched off (current.py:88).  The stream position at an auto-reset therefore depends on the lengths of all
earlier episodes.  The in-kernel auto-reset uses a counter-based Philox stream instead (same distributions, results
independent of sharding); this module offers the reference-exact alternative on the host:

* ``ReferenceResetStream``: one env's generator; ``generate()`` returns the same pose / goal / obstacles / current
  the reference would produce at this point of the stream, ``consume_step()`` accounts for the per-step draw;
* ``ReferenceSeededEnv``: wraps a batched env created with ``auto_reset=False`` and emulates N independent
  reference envs (env i seeded with ``seeds[i]``), resetting finished envs from their own stream.  Host-driven, one
  device-to-host read of the done flags per step: meant for validation and small-N use, not for throughput.
"""
import numpy as np

from . import vehicles as _veh
from .params import SCENARIO_IDS


def _ssa(x):
    return (x + np.pi) % (2 * np.pi) - np.pi


class ReferenceResetStream:
    def __init__(self, scenario, env_config, seed):
        if scenario not in SCENARIO_IDS:
            raise KeyError(scenario)
        self.scenario = scenario
        self.max_attitude = env_config["max_attitude"]
        self.max_dist_from_goal = env_config["max_dist_from_goal"]
        self.rs = np.random.RandomState(seed)          # np.random.seed(seed), docking3d.py:296-298
        self.sigma = 0.0                               # white_noise_std of the current spawned by the last reset

    def consume_step(self):
        """Current.sim draws one normal per step, also with sigma = 0 (current.py:88)."""
        return self.rs.normal(0, self.sigma)

    def generate(self):
        rs, scn = self.rs, self.scenario
        out = {"capsules": np.zeros((0, 7)), "spheres": np.zeros((0, 4))}
        # SimpleDocking3d.generate_environment, docking3d.py:803-825
        goal = np.array([0.0, 0.0, 0.0])
        heading = (rs.random_sample() - 0.5) * np.pi
        rnd = rs.random_sample(3) - 0.5                                    # generate_random_pos, :687-696
        rnd[2] = abs(rnd[0] + rnd[1]) / 3 * np.sign(rnd[2])
        pos = goal + rnd * (15 / np.linalg.norm(rnd))
        att = (rs.random_sample(3) - 0.5) * 2 * np.array([self.max_attitude * 0.7, self.max_attitude * 0.7, np.pi])
        current = np.array([0.0, 0.0, 0.0, 0.0, 0.0])                      # V_c, alpha, beta, V_min, V_max
        self.sigma = 0.0
        if scn == "SimpleCurrentDocking3d":                                # :837-849
            ang = (rs.random_sample(2) - 0.5) * 2 * np.array([np.pi / 2, np.pi])
            speed = rs.random_sample() * 1.0
            current = np.array([0.5, ang[0], ang[1], speed, speed])
        if scn.startswith(("Capsule", "Obstacles")):                       # CapsuleDocking3d, :860-886
            theta = rs.rand() * 2 * np.pi
            radius = 1.0 + _veh.SAFETY_RADIUS
            x, y = np.cos(theta) * radius, np.sin(theta) * radius
            goal = np.array([x, y, (rs.rand() - 0.5) * 4.0])
            top = np.array([0.0, 0.0, -2.0])
            bot = np.array([0.0, 0.0, 0.0]) - (top - np.array([0.0, 0.0, 0.0]))      # shape.py:105-108
            caps = [np.array([*bot, *top, 1.0])]
            d_vec = (bot - top) / np.linalg.norm(bot - top)               # vec_line_point(goal, top, bot), shape.py:420-433
            t = np.dot(goal - top, d_vec)
            vec = (top + t * d_vec) - goal
            heading = _ssa(np.arctan2(vec[1], vec[0]))
            if scn.startswith("Obstacles"):                                # ObstaclesDocking3d, :919-946
                theta = rs.rand() * 2 * np.pi
                half = 2 * self.max_dist_from_goal / 2.0
                for _ in range(4):
                    x, y = np.cos(theta) * 6, np.sin(theta) * 6
                    theta += 2 * np.pi / 4
                    p, tp = np.array([x, y, 0.0]), np.array([x, y, -half])
                    caps.append(np.array([*(p - (tp - p)), *tp, 1.0]))
            if scn == "ObstaclesNoCapDocking3d":                           # :957-965
                caps.pop(0)
            out["capsules"] = np.array(caps)
        if scn in ("CapsuleCurrentDocking3d", "ObstaclesCurrentDocking3d"):   # :897-908, :977-988
            ang = (rs.random_sample(2) - 0.5) * 2 * np.array([np.pi / 2, np.pi])
            current = np.array([0.5, ang[0], ang[1], 0.5, 0.5])
        state = np.zeros(12)
        state[0:3] = pos
        state[3:6] = att
        out.update(init_state=state, goal=goal, heading_goal=float(heading), current=current)
        return out


class ReferenceSeededEnv:
    """N independent reference-seeded envs on top of a batched env built with ``auto_reset=False``."""

    def __init__(self, env, seeds):
        if env.auto_reset:
            raise ValueError("create the batched env with auto_reset=False")
        if len(seeds) != env.num_envs:
            raise ValueError("one seed per env")
        self.env = env
        self.streams = [ReferenceResetStream(env.scenario, env.config, s) for s in seeds]
        self.num_envs = env.num_envs

    def _inject(self, ids, inits):
        n_u = self.env.n_actions
        kw = dict(state=np.stack([d["init_state"] for d in inits]), goal=np.stack([d["goal"] for d in inits]),
                  heading_goal=np.array([d["heading_goal"] for d in inits]),
                  current=np.stack([d["current"] for d in inits]), u_prev=np.zeros((len(ids), n_u)),
                  t_steps=np.zeros(len(ids)), ep_return=np.zeros(len(ids)), env_ids=ids)
        if self.env.n_capsules:
            kw["capsules"] = np.stack([d["capsules"] for d in inits])
        self.env.set_state(**kw)

    def reset(self):
        self.last_init = [s.generate() for s in self.streams]
        self._inject(list(range(self.num_envs)), self.last_init)
        self.env.obs.zero_()
        return self.env.obs

    def step(self, actions):
        obs, reward, done, info = self.env.step(actions)
        d = done.cpu().numpy().astype(bool)
        for s in self.streams:
            s.consume_step()
        ids = np.flatnonzero(d).tolist()
        if ids:
            term = obs[ids].clone()
            inits = [self.streams[i].generate() for i in ids]
            for i, init in zip(ids, inits):
                self.last_init[i] = init
            self._inject(ids, inits)
            obs[ids] = 0                       # reset() returns the all-zero observation (docking3d.py:269,322)
            info = dict(info, terminal_observation_rows=(ids, term))
        return obs, reward, done, info
