"""`ac.eval.main(env, ckpt_path, **cfg)` -- marlbase/ac/eval.py:8-27 without the video: load the checkpoint into the run's model class and sample one
episode per env instance from the policy (the reference's `model.act` samples from the categorical as well)."""
from __future__ import annotations

import torch

from ..config import Config, instantiate
from ..dqn.eval import summarise
from .train import Collector


def main(env, ckpt_path, time_limit, **cfg):
    cfg = Config(cfg)
    model = instantiate(cfg.model, env.single_observation_space, env.single_action_space, cfg, max_envs=env.num_envs, max_episode_length=time_limit)
    print(f"Loading model from {ckpt_path}")
    model.load_state_dict(torch.load(ckpt_path, weights_only=True))
    ln, ret = Collector(env, model, time_limit).collect()
    torch.cuda.synchronize()
    out = summarise(ln, ret)
    env.close()
    return out
