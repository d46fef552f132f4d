"""Per-segment gradient difference between the tensor-core pipeline and the fused FP32 kernel (debug aid, GPU)."""
import ctypes as C, types, sys
import numpy as np, torch
sys.path.insert(0, ".")
from tests.test_tc_backward_gpu import _space, _opt, _store, A
from oracle import learner_ref as lr
from codebase_b200.dqn import model as M
from codebase_b200.lbf import TrajStore

def run(mixer, N, D, T, B, sharing):
    rng = np.random.default_rng(B * 7 + T)
    hp = lr.DqnHP(mixer=mixer)
    cfg = types.SimpleNamespace(optimizer="Adam", lr=hp.lr, gamma=hp.gamma, grad_clip=hp.grad_clip, double_q=True, target_update_interval_or_tau=200, standardise_returns=False)
    m = (M.VDNetwork if mixer else M.QNetwork)([_space(shape=(D,))] * N, [_space(n=A)] * N, cfg, [128, 128], sharing, False, True, "cuda", max_batch=B, max_episode_length=T)
    m.theta.add_(0.02 * torch.randn_like(m.theta)); m.hard_update(); m.theta.add_(0.01 * torch.randn_like(m.theta)); m.params_changed()
    s = _store(rng, 300, N, T, D, bool(mixer))
    idx = rng.integers(0, 300, size=B).astype(np.int32)
    ts = TrajStore(300, N, T, D, m.device)
    for k in ("obs", "act", "rew", "done", "filled"):
        getattr(ts, k).copy_(torch.as_tensor(s[k]))
    idx_d = torch.tensor(idx, device="cuda")
    n = m.n_params
    g = {}
    for tc in (0, 1, 1):
        _opt(b"tensor_core_backward", tc)
        m.update_grads(ts, idx_d); torch.cuda.synchronize()
        x = m.grad.cpu().numpy(); g.setdefault(tc, []).append(x[:n] / x[n + 1])
    P = n // (len(set(m.agent_net)) if hasattr(m, "agent_net") else N)
    segs = [("w1", D * 128), ("b1", 128), ("w2", 128 * 128), ("b2", 128), ("w3", A * 128), ("b3", A)]
    sc = max(1.0, np.abs(g[0][0]).max())
    print(f"case mixer={mixer} N={N} D={D} T={T} B={B}: P={P} scale={sc:.3g} rerun-diff={np.abs(g[1][0]-g[1][1]).max():.3g}")
    for net in range(n // P):
        o = net * P
        for name, sz in segs:
            d = np.abs(g[0][0][o:o + sz] - g[1][0][o:o + sz])
            print(f"  net{net} {name:3s} max|d|/scale={d.max()/sc:.3g} at {int(d.argmax())} ref={np.abs(g[0][0][o:o+sz]).max():.3g}")
            o += sz

for case in [(0, 2, 15, 25, 64, False), (0, 2, 15, 25, 1024, False), (1, 2, 15, 25, 257, False)]:
    run(*case)
