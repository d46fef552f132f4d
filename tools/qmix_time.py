"""Time of one QMIX update at the headline batch shape (1024 episodes x 25 steps, 2 agents, obs 15): python tools/qmix_time.py"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codebase_b200.dqn import model as M  # noqa: E402
from codebase_b200.lbf import TrajStore  # noqa: E402


def main():
    N, D, A, T, B, CAP = 2, 15, 6, 25, 1024, 4096
    cfg = types.SimpleNamespace(optimizer="Adam", lr=3e-4, gamma=0.99, grad_clip=1.0, double_q=True, target_update_interval_or_tau=200, standardise_returns=False)
    sp = lambda **kw: types.SimpleNamespace(shape=kw.get("shape"), n=kw.get("n"))
    out = {}
    for name, cls, extra in (("vdn", M.VDNetwork, {}), ("qmix", M.QMixNetwork, dict(mixing=dict(embed_dim=64, hypernet_layers=2, hypernet_embed=32)))):
        args = [[sp(shape=(D,))] * N, [sp(n=A)] * N, cfg, [128, 128], False, False, True]
        m = cls(*args, *( [extra["mixing"]] if extra else []), "cuda", max_batch=B, max_episode_length=T)
        ts = TrajStore(CAP, N, T, D, m.device)
        ts.obs.copy_(torch.randn_like(ts.obs)); ts.act.copy_(torch.randint(0, A, ts.act.shape)); ts.rew.copy_(torch.rand_like(ts.rew).mean(1, keepdim=True).expand_as(ts.rew))
        ts.filled.fill_(1)
        m.update_n(ts, B, CAP, 1, 0, 20)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.update_n(ts, B, CAP, 1, 20, 200); e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 200 * 1e3
        print(f"{name}: {out[name]:.1f} us per update (batch {B} x T {T})", flush=True)


if __name__ == "__main__":
    main()
