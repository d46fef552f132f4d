"""Profiling aid (GPU): a few IDQN updates at the bench shape.  With a library built with MARL_NVCC_DEFINES=-DMARL_TC_TIMESTAMPS the
tensor-core training kernels print the cycle offsets of their phase boundaries (CTA 5); see tc_common.cuh."""
import ctypes as C, sys, types
import numpy as np, torch
sys.path.insert(0, ".")
from codebase_b200.dqn import model as M
from codebase_b200.lbf import TrajStore
sp = lambda **k: types.SimpleNamespace(**k)
N, D, T, B, A = 2, 15, 25, 1024, 6
cfg = sp(optimizer="Adam", lr=3e-4, gamma=0.99, grad_clip=1.0, double_q=True, target_update_interval_or_tau=200, standardise_returns=False)
m = M.QNetwork([sp(shape=(D,), n=None)] * N, [sp(shape=None, n=A)] * N, cfg, [128, 128], False, False, True, "cuda", max_batch=B, max_episode_length=T)
rng = np.random.default_rng(0)
ts = TrajStore(2000, N, T, D, m.device)
ts.obs.copy_(torch.as_tensor(rng.integers(-1, 5, size=tuple(ts.obs.shape)).astype(np.float32)))
ts.act.copy_(torch.as_tensor(rng.integers(0, A, size=tuple(ts.act.shape)).astype(np.int32)))
ts.filled.fill_(1)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    idx = torch.tensor(rng.integers(0, 2000, size=B).astype(np.int32), device="cuda")
    m.update_grads(ts, idx); m.update_apply(); torch.cuda.synchronize()
