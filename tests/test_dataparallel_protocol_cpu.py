"""CPU, world_size 2 over gloo: the data-parallel protocol of SURVEY §8e as bench.py / the drivers run it on GPUs --
rank r owns global env ids [r*E, (r+1)*E); per update ONE all_reduce(SUM) over [un-normalised gradient sums | loss
numerator | filled count]; every rank divides by the global filled count, clips by the global norm, applies the same Adam
step.  Checked against the single-process result on the concatenated batch (oracle arithmetic; no GPU, no product kernels)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lbf_c
from oracle import learner_ref as lr

N, D, A, T = 2, 15, 6, 25


def _batch(rng, B):
    L = rng.integers(2, T + 1, size=B)
    filled = (np.arange(T)[:, None] < L[None, :]).astype(np.float32)
    dones = np.zeros((T + 1, B), np.float32)
    dones[L, np.arange(B)] = 1
    return dict(obss=torch.tensor(rng.integers(-1, 8, (N, T + 1, B, D)), dtype=torch.float32), actions=torch.tensor(rng.integers(0, A, (N, T, B))),
                rewards=torch.tensor(rng.random((N, T, B)), dtype=torch.float32), dones=torch.tensor(dones), filled=torch.tensor(filled))


def _unnormalised(theta, theta_tgt, batch, hp):
    """What marl_dqn_update_grads leaves in `grad`: sum-gradients, loss numerator, filled count."""
    th = theta.clone().requires_grad_(True)
    fs = batch["filled"].sum()
    loss = lr.dqn_loss(th, theta_tgt, [0, 1], D, A, batch, hp) * fs  # undo the masked mean -> plain sums
    (g,) = torch.autograd.grad(loss, th)
    return torch.cat([g, loss.detach().reshape(1), fs.reshape(1)])


def _worker(rank, world, port, theta0, halves, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    hp = lr.DqnHP()
    st = lr.DqnState(theta0.clone(), theta0.clone(), [0, 1], D, A)
    buf = _unnormalised(st.theta, st.theta_tgt, halves[rank], hp)
    dist.all_reduce(buf)                      # the single exchange of the update
    g = buf[:-2] / buf[-1]
    coef, _ = lr.clip_coef(g, hp.grad_clip)
    st.updates += 1
    lr.adam_step(st.theta, st.m, st.v, g * coef, st.updates, hp.lr)
    out[rank] = (st.theta.clone(), float(buf[-2] / buf[-1]))
    # env sharding: this rank's shard of the global env id space reproduces the unsharded boards
    E = 8
    shard = lbf_c.OracleVecEnv(lbf_c.make_cfg(), E, seed=11, env_gid0=rank * E)
    out[f"obs{rank}"] = torch.tensor(shard.reset())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_update_equals_single_process_update():
    rng = np.random.default_rng(0)
    theta0 = lr.init_flat(2, D, A)
    halves = [_batch(rng, 12), _batch(rng, 20)]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, theta0, halves, out), nprocs=2, join=True)
    full = {k: torch.cat([halves[0][k], halves[1][k]], dim=-1 if k in ("dones", "filled") else 2) for k in halves[0]}
    st = lr.DqnState(theta0.clone(), theta0.clone(), [0, 1], D, A)
    want = lr.dqn_update(st, full, lr.DqnHP())
    for r in (0, 1):
        theta_r, loss_r = out[r]
        assert abs(loss_r - want["loss"]) <= 1e-5 * max(1.0, abs(want["loss"]))
        d = (theta_r - st.theta).abs()
        assert float(d.quantile(0.999)) < 1e-5
    assert torch.equal(out[0][0], out[1][0])  # replicated parameters stay bit-identical across ranks
    whole = lbf_c.OracleVecEnv(lbf_c.make_cfg(), 16, seed=11, env_gid0=0).reset()
    assert np.array_equal(np.concatenate([out["obs0"].numpy(), out["obs1"].numpy()]), whole)
