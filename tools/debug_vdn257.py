"""Diagnostic (GPU): the seeded VDN case of tests/test_dqn_gpu.py::test_update_matches_oracle_on_random_batches[1-False-257-2], update by update, under
every kernel selection -- where does the gradient leave the oracle, and do the forward outputs agree on the rows that feed the double-Q argmax?"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_dqn_gpu as t   # noqa: E402
from oracle import learner_ref as lr   # noqa: E402
from codebase_b200 import _native as nat   # noqa: E402


def opt(name, v):
    nat.check(nat.lib().marl_set_option(name, C.c_int32(int(v))), "opt")


def run(mixer, sharing, B, n_agents, attempt, bwd, pp, onchip=1):
    opt(b"tensor_core_backward", bwd); opt(b"tensor_core_pingpong", pp); opt(b"tensor_core_onchip", onchip)
    torch.manual_seed(1000 * attempt + B)
    rng = np.random.default_rng(B)
    hp = lr.DqnHP(mixer=mixer, target_update_interval_or_tau=2)
    m = t._model("VDNetwork" if mixer else "QNetwork", sharing, hp, n_agents=n_agents, max_batch=B)
    st = lr.DqnState(m.theta.cpu().clone(), m.theta_tgt.cpu().clone(), m.agent_net, t.D, t.A)
    cap, T, D, A = 300, t.T, t.D, t.A
    for u in range(3):
        obs = rng.integers(-1, 8, size=(cap, n_agents, T + 1, D)).astype(np.float32)
        act = rng.integers(0, A, size=(cap, n_agents, T)).astype(np.int32)
        rew = (rng.random((cap, n_agents, T)) < 0.2).astype(np.float32) * rng.random((cap, n_agents, T)).astype(np.float32)
        if mixer:
            rew[:] = rew[:, :1]
        length = rng.integers(1, T + 1, size=cap)
        done = np.zeros((cap, T + 1), np.uint8); filled = np.zeros((cap, T), np.uint8)
        for e in range(cap):
            filled[e, : length[e]] = 1
            done[e, length[e]] = rng.random() < 0.7
        store = dict(obs=obs, act=act, rew=rew, done=done, filled=filled)
        idx = rng.integers(0, cap, size=B).astype(np.int32)
        batch = lr.batch_from_store(store, idx)
        margin = lr.double_q_margin(st, batch, hp)
        # forward agreement on the batch rows: (B, T+1, N, D) dense rows through the product forward
        rows = torch.tensor(obs[idx].transpose(0, 2, 1, 3).reshape(-1, n_agents, D), device="cuda")
        q_gpu = m.q_values(rows).cpu().numpy().reshape(B, T + 1, n_agents, A)
        with torch.no_grad():
            q_or = torch.stack(lr.agents_forward(st.theta, st.agent_net, list(batch["obss"]), D, A)).numpy()   # (N, T+1, B, A)
        q_or = q_or.transpose(2, 1, 0, 3)
        dq = np.abs(q_gpu - q_or).max()
        flips = int((q_gpu.argmax(-1) != q_or.argmax(-1))[:, 1:][np.repeat(filled[idx][:, :, None], n_agents, 2) > 0].sum())
        want = lr.dqn_update(st, batch, hp)
        ts = t._store_to_device(store, m.device)
        m.update_grads(ts, torch.tensor(idx, device="cuda"))
        gr = m.grad.cpu().numpy()
        n = m.n_params
        scale = max(1.0, float(np.abs(want["grad"].numpy()).max()))
        err = np.abs(gr[:n] / gr[n + 1] - want["grad"].numpy()) / scale
        w = int(err.argmax())
        print(f"bwd={bwd} pp={pp} onchip={onchip} mixer={mixer} B={B} attempt={attempt} u={u} margin={margin:.2e} |dq|max={dq:.2e} argmax flips on filled rows={flips} grad err max={err.max():.2e} at {w} (P={n // 2}) "
              f"n>1e-6: {(err > 1e-6).sum()} loss {gr[n] / gr[n + 1]:.8f} vs {want['loss']:.8f}", flush=True)
        m.update_apply()
        m.theta.copy_(st.theta); m.theta_tgt.copy_(st.theta_tgt); m.adam_m.copy_(st.m); m.adam_v.copy_(st.v)
        m.params_changed()


if __name__ == "__main__":
    for bwd, pp, oc in ((0, 0, 0), (1, 2, 0), (1, 2, 1)):
        run(1, False, 257, 2, 0, bwd, pp, oc)
        run(0, False, 1024, 2, 0, bwd, pp, oc)
        run(0, [0, 1, 0], 33, 3, 0, bwd, pp, oc)
