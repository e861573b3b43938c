#!/usr/bin/env python
"""Pin the counter-based (Philox) reset to the reference's ``generate_environment`` (run in the build container,
where /root/reference is mounted; the GPU box only sees the committed ``reset_draws.npz``).

    python tests/golden/make_reset_golden.py

The batched path draws its initial conditions from a Philox4x32-10 stream keyed by (seed, global env id, episode)
instead of the reference's global MT19937; what must match the reference is everything DOWNSTREAM of the uniforms:
which draw feeds which quantity, and the distributions (docking3d.py:687-703, 803-988).  So, for each of the seven
scenarios and a handful of (env id, episode) keys:

  1. the uniforms of the key's Philox stream are computed here (numpy restatement of the generator below);
  2. ``np.random.random`` / ``np.random.rand`` / ``np.random.random_sample`` are patched to serve exactly those
     uniforms, in the order the reference consumes them, through the slot table of the scenario
     (gym_dockauv_b200/csrc/dockauv_env.cuh: 0 heading, 1-3 position, 4-6 attitude, 7 goal angle, 8 goal depth,
     9 pillar phase, 10-11 current direction, 12 current speed);
  3. the UNMODIFIED reference env's public ``reset()`` runs, and its post-reset initial conditions are recorded.

tests/test_reset_golden.py then asserts that the oracle's ``orc_reset_env`` (CPU) and the CUDA ``reset_kernel`` / the
in-step auto-reset (GPU) reproduce these values from the key alone.  A wrong slot, a swapped angle or a different
formula anywhere in the reset shows up as a mismatch.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

M32 = 0xFFFFFFFF

# reference draw order per scenario, as slots of the Philox stream (dockauv_env.cuh)
BASE = [0, 1, 2, 3, 4, 5, 6]                       # heading | generate_random_pos(3) | generate_random_att(3)
SLOTS = {
    "SimpleDocking3d": BASE,                                         # docking3d.py:803-825
    "SimpleCurrentDocking3d": BASE + [10, 11, 12],                   # :837-849 (direction(2), speed)
    "CapsuleDocking3d": BASE + [7, 8],                               # :860-886 (goal angle, goal depth)
    "CapsuleCurrentDocking3d": BASE + [7, 8, 10, 11],                # :897-908
    "ObstaclesDocking3d": BASE + [7, 8, 9],                          # :919-946 (pillar phase)
    "ObstaclesNoCapDocking3d": BASE + [7, 8, 9],                     # :957-965
    "ObstaclesCurrentDocking3d": BASE + [7, 8, 9, 10, 11],           # :977-988
}
SEED = 0x5EED0123456789AB
ENV_IDS = np.array(list(range(12)) + [(1 << 32) + 5 + k for k in range(12)], dtype=np.uint64)   # two contiguous blocks
EPISODES = [0, 1, 7]


def philox4x32_10(c, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11), the generator of gym_dockauv_b200/csrc/dockauv_device.cuh."""
    c = [int(x) for x in c]
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


def philox_uniform(seed, env_id, episode, idx):
    """Draw `idx` of the stream (seed, env id, episode): 53-bit uniform in [0, 1) from two 32-bit words."""
    c = philox4x32_10([env_id & M32, (env_id >> 32) & M32, episode & M32, idx >> 1], seed & M32, (seed >> 32) & M32)
    hi, lo = (c[2], c[3]) if idx & 1 else (c[0], c[1])
    return ((hi >> 5) * 67108864.0 + (lo >> 6)) / 9007199254740992.0


class Feeder:
    """Stands in for the legacy global numpy generator: serves a fixed queue of uniforms."""

    def __init__(self, values):
        self.q = list(values)
        self.served = 0

    def _take(self, n):
        assert len(self.q) >= n, "the reference asked for more uniforms than the slot table lists"
        out, self.q = self.q[:n], self.q[n:]
        self.served += n
        return out

    def random(self, size=None):
        if size is None:
            return self._take(1)[0]
        return np.array(self._take(int(np.prod(size)))).reshape(size)

    def rand(self, *shape):
        return self.random(shape if shape else None)


@contextlib.contextmanager
def patched_uniforms(feeder):
    names = ("random", "random_sample", "rand", "uniform", "normal", "randn", "randint", "choice")
    saved = {n: getattr(np.random, n) for n in names}

    def forbidden(*a, **k):
        raise AssertionError("reset() drew from a generator function the feeder does not cover")
    try:
        for n in names:
            setattr(np.random, n, forbidden)
        np.random.random = feeder.random
        np.random.random_sample = feeder.random
        np.random.rand = feeder.rand
        yield
    finally:
        for n, f in saved.items():
            setattr(np.random, n, f)


def main():
    if not ref_shims.reference_available():
        sys.exit("reference not mounted at " + ref_shims.REFERENCE_ROOT)
    ref_shims.install()
    from make_golden import quiet_config
    from gym_dockauv.envs import docking3d
    out = {"seed": np.array(SEED, dtype=np.uint64), "env_ids": ENV_IDS, "episodes": np.array(EPISODES, dtype=np.int32),
           "scenarios": np.array(list(SLOTS))}
    for name, slots in SLOTS.items():
        env = getattr(docking3d, name)(quiet_config())
        n_k = len(ENV_IDS) * len(EPISODES)
        state = np.zeros((n_k, 12))
        goal = np.zeros((n_k, 3))
        heading = np.zeros(n_k)
        current = np.zeros((n_k, 5))
        caps = None
        uniforms = np.zeros((n_k, 13))
        k = 0
        for ep in EPISODES:
            for gid in ENV_IDS:
                u = [philox_uniform(SEED, int(gid), ep, s) for s in range(13)]
                uniforms[k] = u
                feeder = Feeder([u[s] for s in slots])
                with patched_uniforms(feeder), contextlib.redirect_stdout(io.StringIO()):
                    obs0 = env.reset()
                assert not feeder.q and feeder.served == len(slots), (name, feeder.served, len(slots))
                assert not np.any(obs0)
                state[k] = env.auv.state
                goal[k] = env.goal_location
                heading[k] = env.heading_goal_reached
                c = env.current
                current[k] = [c.V_c, c.alpha, c.beta, c.V_min, c.V_max]
                cc = np.array([[*q.vec_bot, *q.vec_top, q.radius] for q in env.capsules], dtype=np.float64).reshape(-1, 7)
                if caps is None:
                    caps = np.zeros((n_k,) + cc.shape)
                caps[k] = cc
                k += 1
        out[name + "_state"], out[name + "_goal"], out[name + "_heading"] = state, goal, heading
        out[name + "_current"], out[name + "_capsules"], out[name + "_uniforms"] = current, caps, uniforms
        print(f"{name:28s} keys={n_k} draws per reset={len(slots)} capsules={caps.shape[1]}")
    path = os.path.join(HERE, "reset_draws.npz")
    np.savez_compressed(path, **out)
    print(f"reset_draws.npz size={os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
