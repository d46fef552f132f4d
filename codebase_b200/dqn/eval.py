"""`dqn.eval.main(env, ckpt_path, **cfg)` -- marlbase/dqn/eval.py:8-28 without the video: instantiate the model from the run's config, load the
checkpoint (`torch.load(..., weights_only=True)` + `load_state_dict`, the reference's two lines), play one greedy-ish episode
(`eps_evaluation`) per env instance on the device."""
from __future__ import annotations

import numpy as np
import torch

from ..config import Config, instantiate
from .train import Collector


def summarise(final_len, final_ret):
    ln, ret = final_len.cpu().numpy(), final_ret.cpu().numpy().sum(-1)   # the logged episode return is the sum over agents (utils/wrappers.py:33-41)
    return dict(episodes=int(len(ln)), mean_episode_returns=float(ret.mean()), std_episode_returns=float(ret.std()), mean_episode_length=float(ln.mean()),
                episode_returns=[float(x) for x in ret])


def main(env, ckpt_path, time_limit, **cfg):
    cfg = Config(cfg)
    model = instantiate(cfg.model, env.single_observation_space, env.single_action_space, cfg, max_batch=cfg.batch_size, max_episode_length=time_limit)
    print(f"Loading model from {ckpt_path}")
    model.load_state_dict(torch.load(ckpt_path, weights_only=True))
    ln, ret = Collector(env, model, time_limit).collect(None, 0, cfg.eps_evaluation)
    torch.cuda.synchronize()
    out = summarise(ln, ret)
    env.close()
    return out
