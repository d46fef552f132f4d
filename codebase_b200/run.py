"""Entry point -- drop-in for marlbase/run.py:14-47 with the same command line:

    python -m codebase_b200.run +algorithm=idqn env.name="lbforaging:Foraging-8x8-2p-3f-v3" env.time_limit=25 seed=0

Builds logger, env, a second evaluation env, seeds torch/numpy, dispatches `algorithm._target_`, and works in
outputs/<env.name>/<algorithm.name>/<random hex>/ (results.csv, config.yaml) exactly as the reference's Hydra run dir
(configs/default.yaml:7-9)."""
from __future__ import annotations

import logging
import os
import sys

import numpy as np
import torch

from .config import Config, call, compose, instantiate


def main(argv=None):
    cfg = compose(list(sys.argv[1:] if argv is None else argv))
    run_dir = cfg.get("run_dir") or os.path.join("outputs", str(cfg.env.name), str(cfg.algorithm.get("name", "algorithm")), os.urandom(4).hex())
    os.makedirs(run_dir, exist_ok=True)
    os.chdir(run_dir)
    logging.basicConfig(level=logging.INFO, format="[%(asctime)s][%(levelname)s] - %(message)s",
                        handlers=[logging.FileHandler("run.log"), logging.StreamHandler(sys.stdout)], force=True)
    logger = instantiate(cfg.logger, cfg=cfg)
    env = call(cfg.env, seed=cfg.seed)
    # evaluation envs: the reference builds ONE extra env (run.py:21-27); here `eval_episodes` instances run one episode each
    eval_cfg = Config(cfg.env.to_dict())
    eval_cfg["parallel_envs"] = int(cfg.algorithm.eval_episodes)
    eval_env = call(eval_cfg, seed=cfg.seed, env_gid0=1 << 30)
    torch.set_num_threads(1)
    if cfg.seed is not None:
        torch.manual_seed(cfg.seed)
        np.random.seed(cfg.seed)
    else:
        logger.warning("No seed has been set.")
    assert cfg.env.time_limit is not None, "Time limit must be set."
    algo = cfg.algorithm
    algo["seed_for_sampling"] = cfg.seed
    call(algo, env, eval_env, logger, time_limit=cfg.env.time_limit)
    return logger.get_state()


if __name__ == "__main__":
    main()
