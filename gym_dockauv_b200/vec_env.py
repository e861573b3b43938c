"""Stable-Baselines3 `VecEnv`-shaped adapter over a batched docking env.

The reference trains with ``MODEL('MlpPolicy', env=env)`` (gym_dockauv/train.py:64), where SB3 wraps the single env
in ``Monitor`` + ``DummyVecEnv`` and talks to it through ``reset() / step_async() / step_wait()``.  This adapter
offers the same contract for the whole batch in one call:

* ``step_wait()`` returns ``(obs[N, n_obs] float32, rewards[N] float32, dones[N] bool, infos)``;
* finished envs are auto-reset: their ``obs`` row is the reset observation (all zeros for this env family,
  docking3d.py:269,322) and ``infos[i]["terminal_observation"]`` is their last real observation;
* ``infos[i]["episode"] = {"r", "l", "t"}`` like ``Monitor`` adds, from the in-kernel episode counters;
* ``infos`` is a lazy sequence: dicts are only materialised for finished envs, every other index yields an empty
  dict, so a step over 10^5 envs does not allocate 10^5 python dicts.

If stable_baselines3 is importable the class derives from its ``VecEnv`` (so ``isinstance`` checks pass); otherwise it
is a plain class with the same methods.  Host arrays go through ``env.step_host`` (pinned memory, copies pipelined
with the kernel); with ``device_tensors=True`` the adapter hands back CUDA tensors from ``env.step`` instead.
"""
import time
from collections.abc import Sequence

import numpy as np

try:   # optional
    from stable_baselines3.common.vec_env import VecEnv as _SB3VecEnv
except Exception:   # noqa: BLE001
    _SB3VecEnv = object


class LazyInfos(Sequence):
    """infos[i] -> {} unless env i finished this step."""

    def __init__(self, n, filled):
        self._n, self._filled = n, filled

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._filled.get(i, {})

    def finished(self):
        """indices of the envs whose episode ended this step"""
        return sorted(self._filled)


class DockingVecEnv(_SB3VecEnv):
    def __init__(self, env, device_tensors=False):
        self.env = env
        self.device_tensors = bool(device_tensors)
        self._actions = None
        self._t0 = time.time()
        if _SB3VecEnv is not object:
            super().__init__(env.num_envs, env.observation_space, env.action_space)
        else:
            self.num_envs = env.num_envs
            self.observation_space = env.observation_space
            self.action_space = env.action_space

    # ---------------------------------------------------------------- VecEnv API
    def reset(self):
        obs = self.env.reset()
        return obs if self.device_tensors else np.zeros((self.num_envs, self.env.n_observations), dtype=np.float32)

    def seed(self, seed=None):
        if seed is not None:
            self.env.reset(seed=seed)
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        a = self._actions
        self._actions = None
        if self.device_tensors:
            obs, reward, done, info = self.env.step(a)
            idx = done.nonzero().flatten().cpu().numpy()
            if idx.size:
                term = info["terminal_observation"][idx].cpu().numpy()
                ep_r = info["episode_return"][idx].cpu().numpy()
                ep_l = info["episode_length"][idx].cpu().numpy()
                bits = info["cond_bits"][idx].cpu().numpy()
            rewards, dones = reward, done.bool()
        else:
            a = np.ascontiguousarray(a)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float32)
            obs, reward, done, info = self.env.step_host(a)
            idx = np.flatnonzero(done)
            if idx.size:   # episode summaries of the (few) finished envs come from the device buffers
                import torch
                sel = torch.as_tensor(idx, device=self.env.device)
                term = self.env.terminal_obs[sel].cpu().numpy()
                ep_r = self.env.ep_return_out[sel].cpu().numpy()
                ep_l = self.env.ep_len_out[sel].cpu().numpy()
                bits = info["cond_bits"][idx]
            rewards, dones = reward.astype(np.float32), done
        filled = {}
        now = round(time.time() - self._t0, 6)
        for k, i in enumerate(idx.tolist()):
            b = int(bits[k])
            filled[i] = {"terminal_observation": term[k], "episode": {"r": float(ep_r[k]), "l": int(ep_l[k]), "t": now},
                         "conditions_true": [c for c in range(5) if b >> c & 1], "goal_reached": bool(b & 1),
                         "collision": bool(b >> 4 & 1), "TimeLimit.truncated": False}
        return obs, rewards, dones, LazyInfos(self.num_envs, filled)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [getattr(self.env, attr_name)] * n

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.env, attr_name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [getattr(self.env, method_name)(*method_args, **method_kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [False] * n

    def get_images(self):
        raise NotImplementedError("rendering is out of scope of the batched env (SURVEY.md 2, component 13)")

    def render(self, mode="human"):
        raise NotImplementedError("rendering is out of scope of the batched env (SURVEY.md 2, component 13)")

    def _indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return list(indices)
