"""CPU: host-side logic of the product package that needs no GPU -- the PRODUCT epsilon schedule against the golden vector the
live reference produced (tests/golden/misc.npz), config validation of the drivers."""
import os
import warnings

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_product_epsilon_schedule_matches_reference_golden():
    from codebase_b200.dqn.train import epsilon_schedule

    g = np.load(os.path.join(GOLD, "misc.npz"))
    lin = epsilon_schedule("linear", 0.5, 1.0, 0.05, 6.5, 100000)
    exp = epsilon_schedule("exponential", 0.5, 1.0, 0.05, 6.5, 100000)
    assert np.array_equal(np.array([lin(int(s)) for s in g["steps"]]), g["eps_linear"])      # bit-exact Python floats
    assert np.array_equal(np.array([exp(int(s)) for s in g["steps"]]), g["eps_exp"])
    for bad in (dict(decay_style="cosine"), dict(eps_start=0.01), dict(decay_over=0.0), dict(total_steps=0), dict(exp_decay_rate=0.0)):
        kw = dict(decay_style="linear", decay_over=0.5, eps_start=1.0, eps_end=0.05, exp_decay_rate=6.5, total_steps=1000)
        kw.update(bad)
        with pytest.raises(AssertionError):   # the reference's own validation (dqn/train.py:140-150)
            epsilon_schedule(**kw)


def test_iteration_budget_warning():
    """One vectorised iteration must not swallow the whole step budget / epsilon decay (degenerate defaults)."""
    from codebase_b200.dqn.train import check_iteration_budget

    with warnings.catch_warnings():
        warnings.simplefilter("error")
        check_iteration_budget(64, 25, total_steps=100_000, eval_interval=10_000, eps_decay_over=0.5)   # the shipped defaults: quiet
    with pytest.warns(UserWarning, match="parallel_envs"):
        check_iteration_budget(4096, 25, total_steps=100_000, eval_interval=10_000, eps_decay_over=0.5)


def test_shipped_overlays_are_not_degenerate():
    from codebase_b200.config import compose
    from codebase_b200.dqn.train import check_iteration_budget

    for algo in ("idqn", "vdn", "qmix", "ia2c"):
        cfg = compose([f"+algorithm={algo}", "env.name=lbforaging:Foraging-8x8-2p-3f-v3", "env.time_limit=25"])
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            check_iteration_budget(int(cfg["env"]["parallel_envs"]), 25, total_steps=int(cfg["algorithm"]["total_steps"]),
                                   eval_interval=int(cfg["algorithm"]["eval_interval"]), eps_decay_over=float(cfg["algorithm"].get("eps_decay_over", 1.0)))


def test_eval_entry_point_helpers(tmp_path):
    """codebase_b200.eval (marlbase/eval.py): argument parsing and the latest-checkpoint rule (largest N over checkpoints/model_sN.pt)."""
    from codebase_b200 import eval as ev

    d = tmp_path / "checkpoints"
    d.mkdir()
    for n in (100, 25600, 9000):
        (d / f"model_s{n}.pt").write_bytes(b"")
    (d / "model_sX.pt").write_bytes(b""); (d / "notes.txt").write_text("")
    assert ev.latest_step(str(d)) == 25600
    with pytest.raises(FileNotFoundError):
        ev.latest_step(str(tmp_path))
    a = ev.parse_args(["path=outputs/x", "load_step=9000", "seed=null", "episodes=64"])
    assert a == dict(path="outputs/x", load_step=9000, seed=None, episodes=64)
    with pytest.raises(ValueError):
        ev.parse_args(["checkpoint=foo"])


def test_qmix_host_class_rejects_what_the_kernels_do_not_implement():
    """Unsupported reference options fail loudly in Python, before any native call (no silent fallback): single-layer hypernetworks,
    standardise_returns with the mixer."""
    import types

    from codebase_b200.dqn import model as M

    mixing = dict(embed_dim=64, hypernet_layers=1, hypernet_embed=32)
    with pytest.raises(NotImplementedError, match="hypernet_layers"):
        M.QMixNetwork([], [], types.SimpleNamespace(standardise_returns=False), [128, 128], False, False, True, mixing, "cuda")
    with pytest.raises(NotImplementedError, match="standardise_returns"):
        M.QMixNetwork([], [], types.SimpleNamespace(standardise_returns=True), [128, 128], False, False, True, dict(mixing, hypernet_layers=2), "cuda")
