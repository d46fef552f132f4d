"""GPU: learner parity for episode lengths other than the BASELINE's time_limit=25 -- the training kernel walks 128-row tiles
over (T+1)-row episodes, so tile / episode alignment changes with T (carry of q(t+1) across tile boundaries, ragged tiles)."""
import types

import numpy as np
import pytest
import torch

from oracle import learner_ref as lr

from tests.helpers import assert_grad_close, check_margin, redraw_on_near_tie

pytestmark = pytest.mark.gpu
D, A, N = 15, 6, 2


def _close(a, b, rtol=1e-5, atol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.allclose(a, b, rtol=rtol, atol=atol), float(np.abs(a - b).max())


def _space(shape=None, n=None):
    return types.SimpleNamespace(shape=shape, n=n)


def _store(rng, cap, T, coop=False):
    obs = rng.integers(-1, 8, size=(cap, N, T + 1, D)).astype(np.float32)
    act = rng.integers(0, A, size=(cap, N, T)).astype(np.int32)
    rew = (rng.random((cap, N, T)) < 0.2).astype(np.float32) * rng.random((cap, N, T)).astype(np.float32)
    if coop:
        rew[:] = rew[:, :1]
    length = rng.integers(1, T + 1, size=cap)
    done = np.zeros((cap, T + 1), np.uint8); filled = np.zeros((cap, T), np.uint8)
    for e in range(cap):
        filled[e, : length[e]] = 1
        done[e, length[e]] = rng.random() < 0.8
    return dict(obs=obs, act=act, rew=rew, done=done, filled=filled)


def _to_dev(s, T, device):
    from codebase_b200.lbf import TrajStore

    ts = TrajStore(s["obs"].shape[0], N, T, D, device)
    for k in ("obs", "act", "rew", "done", "filled"):
        getattr(ts, k).copy_(torch.as_tensor(s[k]))
    return ts


@pytest.mark.parametrize("T,B,mixer", [(50, 128, 0), (50, 700, 1), (7, 1024, 0), (100, 33, 0), (1, 512, 0), (127, 16, 1), (128, 9, 0)])
@redraw_on_near_tie
def test_dqn_update_various_T(T, B, mixer):
    from codebase_b200.dqn import model as M

    rng = np.random.default_rng(T * 1000 + B)
    hp = lr.DqnHP(mixer=mixer)
    cfg = types.SimpleNamespace(optimizer="Adam", lr=hp.lr, gamma=hp.gamma, grad_clip=hp.grad_clip, double_q=True, target_update_interval_or_tau=200, standardise_returns=False)
    m = (M.VDNetwork if mixer else M.QNetwork)([_space(shape=(D,))] * N, [_space(n=A)] * N, cfg, [128, 128], False, False, True, "cuda", max_batch=B, max_episode_length=T)
    st = lr.DqnState(m.theta.cpu().clone(), m.theta_tgt.cpu().clone(), m.agent_net, D, A)
    cap = 300
    s = _store(rng, cap, T, bool(mixer))
    idx = rng.integers(0, cap, size=B).astype(np.int32)
    batch = lr.batch_from_store(s, idx)
    check_margin(lr, st, batch, hp)   # near-tie in the double-Q argmax: re-drawn by the decorator
    st0 = lr.DqnState(st.theta.clone(), st.theta_tgt.clone(), st.agent_net, D, A)   # dqn_update steps st in place
    want = lr.dqn_update(st, batch, hp)
    m.update_grads(_to_dev(s, T, m.device), torch.tensor(idx, device="cuda"))
    gr = m.grad.cpu().numpy()
    n = m.n_params
    assert_grad_close(lr, st0, batch, hp, gr[:n] / gr[n + 1], want["grad"].numpy())   # re-drawn when a ReLU unit on its kink explains the mismatch
    _close(m.update_apply().cpu().numpy()[0], want["loss"])


@pytest.mark.parametrize("T,P", [(50, 200), (7, 1000), (100, 40)])
def test_a2c_update_various_T(T, P):
    from codebase_b200.ac.model import A2CNetwork

    rng = np.random.default_rng(T + P)
    hp = lr.A2CHP()
    cfg = types.SimpleNamespace(optimizer="Adam", lr=hp.lr, gamma=hp.gamma, grad_clip=hp.grad_clip, n_steps=hp.n_steps, entropy_coef=hp.entropy_coef,
                                value_loss_coef=hp.value_loss_coef, target_update_interval_or_tau=hp.target_update_interval_or_tau, standardise_returns=False)
    net = types.SimpleNamespace(layers=[128, 128], parameter_sharing=False, use_rnn=False, use_orthogonal_init=True, centralised=False)
    m = A2CNetwork([_space(shape=(D,))] * N, [_space(n=A)] * N, cfg, net, net, "cuda", max_envs=P, max_episode_length=T)
    st = lr.A2CState(m.theta[: m.n_actor].cpu().clone(), m.theta[m.n_actor:].cpu().clone(), m.theta_tgt.cpu().clone(), [0, 1], [0, 1], D, A)
    s = _store(rng, P, T)
    s["done"][:] = 0
    for e in range(P):
        s["done"][e, int(s["filled"][e].sum())] = 1
    t = {k: torch.as_tensor(v) for k, v in s.items()}
    batch = dict(obss=t["obs"].permute(2, 0, 1, 3).reshape(T + 1, P, N * D).float(), actions=t["act"].permute(2, 0, 1).long(),
                 rewards=t["rew"].permute(2, 0, 1).float(), dones=t["done"].permute(1, 0).float(), filled=t["filled"].permute(1, 0).float())
    want = lr.a2c_update(st, batch, hp, 0)
    met = m.metrics_dict(m.update_from_store(_to_dev(s, T, m.device), P, 0))
    _close([met["loss"], met["actor_loss"], met["value_loss"], met["entropy"]], [want["loss"], want["actor_loss"], want["value_loss"], want["entropy"]])
    _, ret, _ = m.scratch(P, T)
    _close(ret.permute(2, 1, 0).cpu().numpy(), want["returns"].numpy())
    d = np.abs(m.theta.cpu().numpy() - np.concatenate([st.actor.numpy(), st.critic.numpy()]))
    assert np.quantile(d, 0.999) < 1e-5
