"""The reference's actor-critic training loop restated on the CPU: marlbase/ac/train.py:24-119 (`_collect_trajectories` over `parallel_envs`
env instances with the `running` mask), :170-204 (collect -> `model.update(batch, step)` -> step += t * parallel_envs), marlbase/ac/model.py:147-153
(`act`: a Categorical sample per agent), 1 torch thread as marlbase/run.py:29 mandates.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (tools/learning_curve_ac.py): the env is the pure-Python restatement (oracle/lbf_ref.py), the learners are
oracle/learner_ref.py's `a2c_update` / `ppo_update` (pinned against the live A2CNetwork / PPONetwork, tests/test_ppo.py, tests/golden/ia2c_*.npz).
Covers ia2c / ippo (independent critics) and maa2c / mappo (`critic.centralised`: every critic reads all observations)."""
from __future__ import annotations

import random

import numpy as np
import torch

from . import learner_ref as lr
from .lbf_ref import LBFConfig, WrappedForaging

ALGS = {"ia2c": (False, False), "ippo": (True, False), "maa2c": (False, True), "mappo": (True, True)}   # (PPO update, centralised critic)


class CpuAC:
    def __init__(self, alg: str, cfg: LBFConfig, parallel_envs: int = 64, seed: int = 0, hp: lr.A2CHP | None = None, num_epochs: int = 4, ppo_clip: float = 0.2):
        torch.set_num_threads(1)
        torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
        self.ppo, central = ALGS[alg]
        self.cfg, self.P, self.hp, self.num_epochs, self.ppo_clip = cfg, parallel_envs, hp or lr.A2CHP(), num_epochs, ppo_clip
        self.N, self.D, self.A, self.T = cfg.n_agents, cfg.obs_dim, 6, cfg.time_limit
        self.envs = [WrappedForaging(cfg, seed, env_gid=i) for i in range(parallel_envs)]
        critic_in = self.N * self.D if central else self.D
        actor, critic = lr.init_flat(self.N, self.D, self.A), lr.init_flat(self.N, critic_in, 1)
        self.st = lr.A2CState(actor, critic, critic.clone(), list(range(self.N)), list(range(self.N)), self.D, self.A, centralised=central)
        self.step = self.updates = 0

    def act(self, obss):
        """ac/model.py:147-153 for all envs at once: obss = per-agent (P, D) tensors -> actions (P, N)"""
        with torch.no_grad():
            logits = lr.agents_forward(self.st.actor, self.st.actor_net, obss, self.D, self.A)
            return torch.stack([torch.distributions.Categorical(logits=l).sample() for l in logits], dim=-1)

    def collect(self):
        P, N, D, T = self.P, self.N, self.D, self.T
        running = np.ones(P, bool)
        obs = np.stack([np.concatenate(e.reset()[0]) for e in self.envs]).astype(np.float32)           # (P, N*D)
        b = dict(obss=torch.zeros(T + 1, P, N * D), actions=torch.zeros(T, P, N, dtype=torch.long), rewards=torch.zeros(T, P, N),
                 dones=torch.zeros(T + 1, P), filled=torch.zeros(T, P))
        b["obss"][0] = torch.tensor(obs)
        t, infos = 0, []
        while running.any():
            actions = self.act(list(torch.split(torch.tensor(obs), D, dim=-1))).numpy()
            for i, e in enumerate(self.envs):
                if not running[i]:
                    continue   # the reference keeps stepping the auto-reset env; nothing of it reaches the batch (masked by `running`)
                o, rew, done, trunc, info = e.step(actions[i].tolist())
                obs[i] = np.concatenate(o)
                b["obss"][t + 1, i] = torch.tensor(obs[i]); b["actions"][t, i] = torch.tensor(actions[i]); b["rewards"][t, i] = torch.tensor(np.asarray(rew, np.float32))
                b["dones"][t + 1, i] = float(done or trunc); b["filled"][t, i] = 1.0       # use_proper_termination: False (ia2c.yaml)
                if done or trunc:
                    infos.append(info); running[i] = False
            t += 1
        return t, b, infos

    def iteration(self):
        t, batch, infos = self.collect()
        if self.ppo:
            lr.ppo_update(self.st, batch, self.hp, self.step, self.num_epochs, self.ppo_clip)
        else:
            lr.a2c_update(self.st, batch, self.hp, self.step)
        self.updates += 1
        at = self.step
        self.step += t * self.P
        return at, float(np.mean([np.sum(i["episode_returns"]) for i in infos]))
