#!/usr/bin/env python
"""Stress (GPU): gradient of one IDQN / VDN update against the CPU oracle for many seeded random initialisations, each run twice
(run-to-run determinism).  Prints the worst relative error per (config, seed)."""
import sys
import types

import numpy as np
import torch

sys.path.insert(0, ".")
from codebase_b200.dqn import model as M  # noqa: E402
from codebase_b200.lbf import TrajStore  # noqa: E402
from oracle import learner_ref as lr  # noqa: E402

N_SEEDS = int(sys.argv[1]) if len(sys.argv) > 1 else 6
D, A, T = 15, 6, 25
sp = lambda **k: types.SimpleNamespace(**{"shape": None, "n": None, **k})
bad = 0
for mixer, sharing, B, n_agents in ((0, False, 1024, 2), (1, False, 257, 2), (0, True, 100, 2), (0, False, 64, 2)):
    for seed in range(N_SEEDS):
        torch.manual_seed(seed)
        rng = np.random.default_rng(B)
        hp = lr.DqnHP(mixer=mixer, target_update_interval_or_tau=2)
        cfg = types.SimpleNamespace(optimizer="Adam", lr=hp.lr, gamma=hp.gamma, grad_clip=hp.grad_clip, double_q=True, target_update_interval_or_tau=2.0, standardise_returns=False)
        m = (M.VDNetwork if mixer else M.QNetwork)([sp(shape=(D,))] * n_agents, [sp(n=A)] * n_agents, cfg, [128, 128], sharing, False, True, "cuda", max_batch=B, max_episode_length=T)
        st = lr.DqnState(m.theta.cpu().clone(), m.theta_tgt.cpu().clone(), m.agent_net, D, A)
        cap = 300
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
        want = lr.dqn_update(st, lr.batch_from_store(store, idx), hp)
        ts = TrajStore(cap, n_agents, T, D, m.device)
        for k in ("obs", "act", "rew", "done", "filled"):
            getattr(ts, k).copy_(torch.as_tensor(store[k]))
        errs = []
        for rep in range(2):
            m.update_grads(ts, torch.tensor(idx, device="cuda"))
            gr = m.grad.cpu().numpy()
            g = gr[:m.n_params] / gr[m.n_params + 1]
            w = want["grad"].numpy()
            errs.append(float(np.abs(g - w).max() / max(1.0, np.abs(w).max())))
        flag = "" if max(errs) < 1e-5 else "   <-- BAD"
        bad += bool(flag)
        print(f"mixer={mixer} sharing={sharing} B={B} seed={seed}: rel err {errs[0]:.2e} {errs[1]:.2e} loss {gr[m.n_params] / gr[m.n_params + 1]:.6f} vs {want['loss']:.6f}{flag}")
        m.close()
print("bad cases:", bad)
