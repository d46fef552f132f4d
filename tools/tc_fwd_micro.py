#!/usr/bin/env python
"""Micro-benchmark (GPU): the forward-only tensor-core kernel on the bench's target-network shape (53 248 rows = 2 nets x 26 624 rows, 148 CTAs x
~2.8 tiles), one-tile-at-a-time kernel vs two-accumulator kernel, CUDA events over 200 launches.  With MARL_B200_SO pointing at a
-DMARL_TC_TIMESTAMPS build the kernels also print their phase stamps (CTA 5)."""
import ctypes as C
import json
import sys
import types

import numpy as np
import torch

sys.path.insert(0, ".")
from codebase_b200 import _native as nat  # noqa: E402
from codebase_b200.dqn.model import QNetwork  # noqa: E402

sp = lambda **k: types.SimpleNamespace(**{"shape": None, "n": None, **k})
cfg = types.SimpleNamespace(optimizer="Adam", lr=3e-4, gamma=0.99, grad_clip=1.0, double_q=True, target_update_interval_or_tau=200, standardise_returns=False)
m = QNetwork([sp(shape=(15,))] * 2, [sp(n=6)] * 2, cfg, [128, 128], False, False, True, "cuda", max_batch=8, max_episode_length=25)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 26624
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
obs = torch.tensor(np.random.default_rng(0).integers(-1, 8, size=(E, 2, 15)).astype(np.float32), device="cuda")
out = torch.empty(E, 2, 6, device="cuda")
res = {}
for pp in (0, 1):
    nat.check(nat.lib().marl_set_option(b"tensor_core_pingpong", C.c_int32(pp)), "opt")
    for _ in range(5):
        m.q_values(obs, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m.q_values(obs, out=out)
    e1.record(); torch.cuda.synchronize()
    res["pingpong" if pp else "one_tile"] = 1e3 * e0.elapsed_time(e1) / reps
print(json.dumps({"rows": 2 * E, "us_per_launch": res}))
