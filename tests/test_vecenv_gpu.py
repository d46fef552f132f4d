"""GPU: the gym-style surface of the product env (B200VecEnv.reset / step / final_info, the reference's vector protocol:
marlbase/ac/train.py:30-34,79-110, marlbase/utils/wrappers.py:36-41) against the wrapped CPU oracle; the product QNetwork.act;
the on-policy collector re-used across iterations (fresh zero batch every call, marlbase/ac/train.py:36-52)."""
import random
import types

import numpy as np
import pytest
import torch

from oracle import learner_ref as lr
from oracle.lbf_ref import LBFConfig, WrappedForaging

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,kw,P", [("lbforaging:Foraging-8x8-2p-3f-v3", dict(), 6), ("lbforaging:Foraging-5x5-3p-1f-v3", dict(), 1),
                                       ("lbforaging:Foraging-2s-10x10-3p-3f-coop-v3", dict(wrappers=["CooperativeReward"]), 5)])
def test_vecenv_protocol_matches_wrapped_oracle(name, kw, P):
    from codebase_b200.lbf import parse_env_id
    from codebase_b200.utils.envs import make_env

    seed, T = 5, 25
    env = make_env(seed, name=name, time_limit=T, parallel_envs=P, **kw)
    c = parse_env_id(name, T)
    ocfg = LBFConfig(rows=c.rows, cols=c.cols, n_agents=c.n_agents, max_num_food=c.max_num_food, sight=c.sight, max_player_level=c.max_player_level,
                     force_coop=c.force_coop, penalty=c.penalty, time_limit=T, cooperative_reward=int("wrappers" in kw))
    orc = [WrappedForaging(ocfg, seed, env_gid=i) for i in range(P)]
    N = c.n_agents
    assert env.unwrapped.n_agents == N and len(env.single_observation_space) == N and env.observation_space[0].shape[0] == P
    obs, info = env.reset()
    want = [o.reset()[0] for o in orc]
    assert info == {} and len(obs) == N
    for i in range(N):
        assert obs[i].dtype == np.float32 and np.array_equal(obs[i], np.stack([w[i] for w in want]))
    rng = np.random.default_rng(0)
    finished = 0
    for _ in range(70):
        acts = rng.integers(0, 6, size=(N, P))
        obs, rew, done, trunc, info = env.step(acts.tolist() if P > 1 else acts[:, 0].tolist())
        rew = np.asarray(rew, np.float32)
        assert rew.shape == (P, N) and done.shape == (P,) and done.dtype == bool and trunc.dtype == bool
        for e in range(P):
            o, r, d, tr, inf = orc[e].step(acts[:, e].tolist())
            assert np.array_equal(rew[e], np.asarray(r, np.float32)) and bool(done[e]) == d and bool(trunc[e]) == tr
            if d or tr:
                finished += 1
                fi = info["final_info"][e]
                assert info["_final_info"][e] and fi is not None
                assert np.array_equal(fi["episode_returns"], inf["episode_returns"]) and fi["episode_returns"].dtype == np.float32
                assert fi["episode_length"] == inf["episode_length"] and fi["episode_time"] >= 0.0
                for i in range(N):
                    assert fi[f"agent{i}/episode_returns"] == inf[f"agent{i}/episode_returns"]
                o = orc[e].reset()[0]      # same-step autoreset: the returned observation opens the next episode
            elif "final_info" in info:
                assert info["final_info"][e] is None and not info["_final_info"][e]
            for i in range(N):
                assert np.array_equal(obs[i][e], o[i])
    assert finished >= 2 * P
    with pytest.raises(ValueError):
        env.step(np.zeros((P + 1, N), np.int64))
    env.close()


def _qmodel(n_agents=2):
    from codebase_b200.dqn.model import QNetwork

    sp = lambda **k: types.SimpleNamespace(**{"shape": None, "n": None, **k})
    cfg = types.SimpleNamespace(optimizer="Adam", lr=3e-4, gamma=0.99, grad_clip=1.0, double_q=True, target_update_interval_or_tau=200, standardise_returns=False)
    return QNetwork([sp(shape=(15,))] * n_agents, [sp(n=6)] * n_agents, cfg, [128, 128], False, False, True, "cuda", max_batch=8, max_episode_length=25)


def test_product_act_follows_the_reference_stream():
    """dqn/model.py:94-116: greedy = per-agent argmax of the network; exploring = ONE random.random() test, then a random joint action."""
    m = _qmodel()
    rng = np.random.default_rng(1)
    obs = [rng.integers(-1, 8, size=15).astype(np.float32) for _ in range(2)]
    want_q = lr.agents_forward(m.theta.cpu(), m.agent_net, [torch.tensor(o).view(1, 1, -1) for o in obs], 15, 6)
    acts, hid = m.act(obs, m.init_hiddens(1), 0.0)
    assert acts == [int(q.argmax(-1)) for q in want_q] and hid == [None, None]
    random.seed(3)
    a1, _ = m.act(obs, None, 1.0)
    random.seed(3)
    assert random.random() < 1.0
    assert a1 == [random.randrange(6) for _ in range(2)]
    random.seed(3)
    a2, _ = m.act(obs, None, 1.0)
    assert a1 == a2   # seeded by Python's `random`, as in the reference
    m.close()
    assert m.theta is None and m.grad is None   # views of freed library memory are dropped


def test_onpolicy_collector_hands_out_a_fresh_zero_batch_every_call():
    """The reference allocates zero batch_* tensors per call (ac/train.py:36-52) and compute_nstep_returns never looks at `filled`:
    after an early episode end the tail must read 0, not the previous batch's longer episode."""
    from codebase_b200.ac.model import A2CNetwork
    from codebase_b200.ac.train import Collector
    from codebase_b200.utils.envs import make_env

    P, T, D, N = 512, 25, 9, 2
    envs = make_env(11, name="lbforaging:Foraging-5x5-2p-1f-v3", time_limit=T, parallel_envs=P)
    sp = lambda **k: types.SimpleNamespace(**{"shape": None, "n": None, **k})
    hp = lr.A2CHP()
    cfg = types.SimpleNamespace(optimizer="Adam", lr=hp.lr, gamma=hp.gamma, grad_clip=hp.grad_clip, n_steps=hp.n_steps, entropy_coef=hp.entropy_coef,
                                value_loss_coef=hp.value_loss_coef, target_update_interval_or_tau=hp.target_update_interval_or_tau, standardise_returns=False)
    net = types.SimpleNamespace(layers=[128, 128], parameter_sharing=False, use_rnn=False, use_orthogonal_init=True, centralised=False)
    m = A2CNetwork([sp(shape=(D,))] * N, [sp(n=6)] * N, cfg, net, net, "cuda", max_envs=P, max_episode_length=T)
    coll = Collector(envs, m, T)
    len1, _ = coll.collect()
    len1 = len1.cpu().numpy().copy()
    len2, _ = coll.collect()
    len2 = len2.cpu().numpy()
    assert (len2 < len1).sum() > 20, "the test needs slots whose second episode is shorter than the first"
    b = coll.batch
    s = {k: getattr(b, k).cpu().numpy() for k in ("obs", "act", "rew", "done", "filled")}
    for e in range(P):
        L = int(len2[e])
        assert s["filled"][e, :L].all() and not s["filled"][e, L:].any()
        assert not s["obs"][e, :, L + 1:].any() and not s["rew"][e, :, L:].any() and not s["act"][e, :, L:].any() and not s["done"][e, L + 1:].any()
    # and the learner's n-step returns on this batch equal the oracle's on the same (zero-tailed) data
    st = lr.A2CState(m.theta[: m.n_actor].cpu().clone(), m.theta[m.n_actor:].cpu().clone(), m.theta_tgt.cpu().clone(), [0, 1], [0, 1], D, 6)
    t = {k: torch.as_tensor(v) for k, v in s.items()}
    ob = dict(obss=t["obs"].permute(2, 0, 1, 3).reshape(T + 1, P, N * D).float(), actions=t["act"].permute(2, 0, 1).long(),
              rewards=t["rew"].permute(2, 0, 1).float(), dones=t["done"].permute(1, 0).float(), filled=t["filled"].permute(1, 0).float())
    want = lr.a2c_update(st, ob, hp, 7)
    met = m.metrics_dict(m.update_from_store(b, P, 7))
    _, ret, _ = m.scratch(P, T)
    assert np.allclose(ret.permute(2, 1, 0).cpu().numpy(), want["returns"].numpy(), rtol=1e-5, atol=1e-5)
    assert np.allclose([met["loss"], met["value_loss"]], [want["loss"], want["value_loss"]], rtol=1e-5, atol=1e-5)
