"""Diagnostic (GPU): gradient error per parameter block for one (T, B, mixer) case of tests/test_episode_lengths_gpu.py, on-chip vs streaming pass."""
import ctypes as C
import sys
import types

import numpy as np
import torch

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_episode_lengths_gpu as t   # noqa: E402
from oracle import learner_ref as lr   # noqa: E402
from codebase_b200 import _native as nat   # noqa: E402
from codebase_b200.dqn import model as M   # noqa: E402

T, B, mixer = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (127, 16, 1)
D, A, N = t.D, t.A, t.N
from tests.helpers import TIE
attempt = 0
while True:   # the test's redraw_on_near_tie: first attempt whose oracle argmax margin is healthy
    torch.manual_seed(7919 * attempt + 17)
    rng = np.random.default_rng(T * 1000 + B)
    hp = lr.DqnHP(mixer=mixer)
    cfg = types.SimpleNamespace(optimizer="Adam", lr=hp.lr, gamma=hp.gamma, grad_clip=hp.grad_clip, double_q=True, target_update_interval_or_tau=200, standardise_returns=False)
    m = (M.VDNetwork if mixer else M.QNetwork)([t._space(shape=(D,))] * N, [t._space(n=A)] * N, cfg, [128, 128], False, False, True, "cuda", max_batch=B, max_episode_length=T)
    st = lr.DqnState(m.theta.cpu().clone(), m.theta_tgt.cpu().clone(), m.agent_net, D, A)
    s = t._store(rng, 300, T, bool(mixer))
    idx = rng.integers(0, 300, size=B).astype(np.int32)
    mg = lr.double_q_margin(st, lr.batch_from_store(s, idx), hp)
    print("attempt", attempt, "margin", mg)
    if mg >= TIE:
        break
    attempt += 1
for onchip in (0, 1):
    nat.check(nat.lib().marl_set_option(b"tensor_core_onchip", C.c_int32(onchip)), "opt")
    torch.manual_seed(7919 * attempt + 17)
    rng = np.random.default_rng(T * 1000 + B)
    hp = lr.DqnHP(mixer=mixer)
    cfg = types.SimpleNamespace(optimizer="Adam", lr=hp.lr, gamma=hp.gamma, grad_clip=hp.grad_clip, double_q=True, target_update_interval_or_tau=200, standardise_returns=False)
    m = (M.VDNetwork if mixer else M.QNetwork)([t._space(shape=(D,))] * N, [t._space(n=A)] * N, cfg, [128, 128], False, False, True, "cuda", max_batch=B, max_episode_length=T)
    st = lr.DqnState(m.theta.cpu().clone(), m.theta_tgt.cpu().clone(), m.agent_net, D, A)
    s = t._store(rng, 300, T, bool(mixer))
    idx = rng.integers(0, 300, size=B).astype(np.int32)
    batch = lr.batch_from_store(s, idx)
    want = lr.dqn_update(st, batch, hp)["grad"].numpy()
    m.update_grads(t._to_dev(s, T, m.device), torch.tensor(idx, device="cuda"))
    gr = m.grad.cpu().numpy(); n = m.n_params
    g = gr[:n] / gr[n + 1]
    P = n // 2
    blocks = [("w1", 0, 128 * D), ("b1", 128 * D, 128 * D + 128), ("w2", 128 * D + 128, 128 * D + 128 + 16384), ("b2", 128 * D + 128 + 16384, 128 * D + 256 + 16384),
              ("w3", 128 * D + 256 + 16384, 128 * D + 256 + 16384 + A * 128), ("b3", 128 * D + 256 + 16384 + A * 128, P)]
    for net in range(2):
        print(f"onchip={onchip} net {net}: " + "  ".join(f"{nm} {np.abs(g[net * P + a: net * P + b] - want[net * P + a: net * P + b]).max():.2e}/{np.abs(want[net * P + a: net * P + b]).max():.2e}" for nm, a, b in blocks), flush=True)
