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
