#!/usr/bin/env python
"""Minimal on-device PPO loop over the batched docking env -- the counterpart of the reference's
``train()`` (gym_dockauv/train.py:21-82: ``MODEL('MlpPolicy', env).learn(...)``) without a host round trip per step:

    python examples/ppo_device_rollout.py [--envs 16384] [--steps 64] [--iters 20]

* rollouts: ``DeviceRolloutBuffer.collect`` -- the step kernels write observation / reward / done of step t straight
  into row t of the buffer;
* advantages: ``dockauv_gae`` (one kernel over the stacked rows);
* the policy / value networks and the clipped-surrogate update are plain PyTorch (they are not part of the step path).

Prints, per iteration, the mean return of the episodes that ended inside the rollout and the env-steps/s of collection.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from gym_dockauv_b200 import envs  # noqa: E402
from gym_dockauv_b200.config import TRAIN_CONFIG  # noqa: E402
from gym_dockauv_b200.rollout import DeviceRolloutBuffer  # noqa: E402


class ActorCritic(nn.Module):
    def __init__(self, n_obs, n_act, hidden=64):
        super().__init__()
        self.pi = nn.Sequential(nn.Linear(n_obs, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, n_act))
        self.v = nn.Sequential(nn.Linear(n_obs, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, 1))
        self.log_std = nn.Parameter(torch.full((n_act,), -0.5))

    def dist(self, obs):
        return torch.distributions.Normal(self.pi(obs), self.log_std.exp())

    @torch.no_grad()
    def act(self, obs):
        d = self.dist(obs)
        a = d.sample()
        return a, self.v(obs).squeeze(-1), d.log_prob(a).sum(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--scenario", default="ObstaclesDocking3d")
    args = ap.parse_args()
    env = envs.SCENARIOS[args.scenario](TRAIN_CONFIG, num_envs=args.envs, seed=0)
    buf = DeviceRolloutBuffer(env, args.steps, gamma=0.99, gae_lambda=0.95)
    buf.reset_env()
    net = ActorCritic(env.n_observations, env.n_actions).to(env.device)
    opt = torch.optim.Adam(net.parameters(), lr=3e-4)
    for it in range(args.iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        last_obs = buf.collect(net.act)
        with torch.no_grad():
            buf.compute_returns_and_advantage(net.v(last_obs).squeeze(-1))
        torch.cuda.synchronize()
        t_collect = time.perf_counter() - t0
        info = buf.episode_infos()
        for _ in range(4):
            for mb in buf.get(batch_size=1 << 16):
                adv = (mb.advantages - mb.advantages.mean()) / (mb.advantages.std() + 1e-8)
                d = net.dist(mb.observations)
                ratio = (d.log_prob(mb.actions).sum(-1) - mb.old_log_prob).exp()
                loss_pi = -torch.min(ratio * adv, ratio.clamp(0.8, 1.2) * adv).mean()
                loss_v = (net.v(mb.observations).squeeze(-1) - mb.returns).pow(2).mean()
                loss = loss_pi + 0.5 * loss_v - 0.0 * d.entropy().sum(-1).mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(net.parameters(), 0.5)
                opt.step()
        n_ep = info["l"].numel()
        mean_r = float(info["r"].mean()) if n_ep else float("nan")
        mean_l = float(info["l"].float().mean()) if n_ep else float("nan")
        print(f"iter {it:3d}  episodes {n_ep:7d}  mean return {mean_r:10.2f}  mean length {mean_l:7.1f}  "
              f"collect {args.envs * args.steps / t_collect:.3g} env-steps/s (policy included)", flush=True)
    env.close()


if __name__ == "__main__":
    main()
