"""GPU: dqn.train.main end to end through run.py's command line (Hydra-style overrides), results.csv schema as the
reference's FileSystemLogger writes it (marlbase/utils/loggers.py:149-165), and a short learning sanity check."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

IDQN_COLS = ["environment_steps", "agent0/mean_episode_returns", "agent0/std_episode_returns", "agent1/mean_episode_returns", "agent1/std_episode_returns", "epsilon",
             "loss", "mean_episode_length", "mean_episode_returns", "mean_episode_time", "std_episode_length", "std_episode_returns", "std_episode_time", "updates"]


@pytest.mark.parametrize("alg", ["idqn", "vdn", "qmix"])
def test_driver_writes_reference_schema(tmp_path, monkeypatch, alg):
    from codebase_b200 import run

    monkeypatch.chdir(tmp_path)
    run.main([f"+algorithm={alg}", "env.name=lbforaging:Foraging-8x8-2p-3f-v3", "env.time_limit=25", "env.parallel_envs=256", "seed=0",
              "algorithm.total_steps=60000", "algorithm.eval_interval=20000", "algorithm.batch_size=128", "algorithm.buffer_size=4096",
              "algorithm.updates_per_iteration=16", f"run_dir={tmp_path}/out"])
    df = pd.read_csv(tmp_path / "out" / "results.csv")
    assert list(df.columns) == IDQN_COLS  # environment_steps first, the rest sorted
    assert len(df) >= 2 and df["updates"].iloc[-1] > 0 and np.isfinite(df["loss"].iloc[-1])
    assert (tmp_path / "out" / "config.yaml").exists()
    assert df["mean_episode_length"].between(1, 25).all()


def test_driver_with_observe_id_shared_parameters_and_standardised_rewards(tmp_path, monkeypatch):
    """env.observe_id=True (one-hot agent id in front of the observation: the reference's companion of full parameter sharing) and
    env.standardise_rewards=True through the same command line; the networks take obs_dim + n_agents = 17 inputs."""
    from codebase_b200 import run

    monkeypatch.chdir(tmp_path)
    run.main(["+algorithm=idqn", "env.name=lbforaging:Foraging-8x8-2p-3f-v3", "env.time_limit=25", "env.parallel_envs=256", "env.observe_id=True",
              "env.standardise_rewards=True", "algorithm.model.parameter_sharing=True", "seed=0", "algorithm.total_steps=60000",
              "algorithm.eval_interval=20000", "algorithm.batch_size=128", "algorithm.buffer_size=4096", "algorithm.updates_per_iteration=16",
              f"run_dir={tmp_path}/out"])
    df = pd.read_csv(tmp_path / "out" / "results.csv")
    assert list(df.columns) == IDQN_COLS and len(df) >= 2 and np.isfinite(df["loss"].iloc[-1])
    # RecordEpisodeStatistics sits inside StandardiseReward: the logged returns are the raw ones (LBF: within [0, 1] per episode in total)
    assert df["mean_episode_returns"].between(0.0, 1.0 + 1e-6).all()


@pytest.mark.parametrize("alg", ["idqn", "qmix", "ia2c"])
def test_checkpoint_eval_round_trip(tmp_path, monkeypatch, alg):
    """save_interval writes checkpoints/model_s<step>.pt with the reference's state_dict keys; `python -m codebase_b200.eval path=<run dir>`
    (marlbase/eval.py's arguments) rebuilds env + model from <run dir>/config.yaml, loads the latest checkpoint and plays evaluation episodes."""
    import json
    import os

    import torch

    from codebase_b200 import eval as ev
    from codebase_b200 import run

    monkeypatch.chdir(tmp_path)
    out = f"{tmp_path}/out"
    extra = ["algorithm.batch_size=128", "algorithm.buffer_size=4096", "algorithm.updates_per_iteration=16"] if alg != "ia2c" else []
    run.main([f"+algorithm={alg}", "env.name=lbforaging:Foraging-8x8-2p-3f-v3", "env.time_limit=25", "env.parallel_envs=256", "seed=0",
              "algorithm.total_steps=40000", "algorithm.eval_interval=20000", "algorithm.save_interval=15000", f"run_dir={out}"] + extra)
    monkeypatch.chdir(tmp_path)
    steps = sorted(int(f[7:-3]) for f in os.listdir(f"{out}/checkpoints"))
    assert len(steps) >= 2
    sd = torch.load(f"{out}/checkpoints/model_s{steps[-1]}.pt", weights_only=True)
    want_key = {"idqn": "critic.independent.0.network.0.weight", "qmix": "target_mixer.hyper_w_1.2.weight", "ia2c": "actor.independent.1.network.4.bias"}[alg]
    assert want_key in sd, sorted(sd)[:8]
    res = ev.main([f"path={out}", "episodes=64", "seed=3"])
    assert res["load_step"] == steps[-1] and res["episodes"] == 64 and np.isfinite(res["mean_episode_returns"]) and 0.0 <= res["mean_episode_returns"] <= 1.0 + 1e-6
    assert json.load(open(f"{out}/eval_s{steps[-1]}.json"))["episode_returns"] == res["episode_returns"]
    first = ev.main([f"path={out}", f"load_step={steps[0]}", "episodes=64", "seed=3"])
    assert first["load_step"] == steps[0] and os.path.exists(f"{out}/eval_s{steps[0]}.json")
