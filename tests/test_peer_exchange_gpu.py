"""GPU (needs 2 devices; skipped on a one-GPU box): the gradient exchange over NVLink peer memory inside the fused reduce + Adam kernel
(marl_dqn_peer_handle / marl_dqn_peer_attach) against the two-call form with an explicit all-reduce between update_grads and
update_apply.  Two ranks, different replay data per rank, identical initial parameters: after three updates the parameters must be
bit-identical across ranks and equal to the all-reduce path up to the rounding of the local partial sums (the fused kernel and
grad_reduce_kernel add the per-CTA partials in different fixed orders)."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
A, N, D, T, B = 6, 2, 15, 25, 96


def _space(shape=None, n=None):
    return types.SimpleNamespace(shape=shape, n=n)


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from codebase_b200.dqn import model as M
    from codebase_b200.lbf import TrajStore

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    cfg = types.SimpleNamespace(optimizer="Adam", lr=3e-4, gamma=0.99, grad_clip=1.0, double_q=True, target_update_interval_or_tau=2, standardise_returns=False)

    def make():
        m = M.QNetwork([_space(shape=(D,))] * N, [_space(n=A)] * N, cfg, [128, 128], False, False, True, f"cuda:{rank}", max_batch=B, max_episode_length=T)
        rng0 = np.random.default_rng(1234)  # the same parameters on every rank
        m.theta.copy_(torch.as_tensor(0.05 * rng0.standard_normal(m.theta.numel()), dtype=torch.float32).view_as(m.theta))
        m.params_changed(); m.hard_update()
        return m

    rng = np.random.default_rng(100 + rank)  # different data per rank
    ts = TrajStore(200, N, T, D, torch.device(f"cuda:{rank}"))
    ts.obs.copy_(torch.as_tensor(rng.integers(-1, 9, size=tuple(ts.obs.shape)).astype(np.float32)))
    ts.act.copy_(torch.as_tensor(rng.integers(0, A, size=tuple(ts.act.shape)).astype(np.int32)))
    ts.rew.copy_(torch.as_tensor((rng.random(tuple(ts.rew.shape)) < 0.3).astype(np.float32)))
    ts.filled.fill_(1)
    idx = [torch.tensor(rng.integers(0, 200, size=B).astype(np.int32), device=f"cuda:{rank}") for _ in range(3)]

    peer = make()
    peer.attach_peers()
    for k in range(3):
        peer.update_from_store(ts, idx[k])
    torch.cuda.synchronize()
    ref = make()
    for k in range(3):
        ref.update_grads(ts, idx[k])
        g = ref.grad.cpu()
        dist.all_reduce(g)
        ref.grad.copy_(g)
        ref.update_apply()
    torch.cuda.synchronize()
    assert not peer.peer_timed_out(), "the in-kernel exchange gave up waiting for a peer"
    th = peer.theta.cpu()
    gathered = [torch.empty_like(th) for _ in range(world)]
    dist.all_gather(gathered, th)
    out.put((rank, bool(all(torch.equal(gathered[0], t) for t in gathered)), float((th - ref.theta.cpu()).abs().max()),
             float((peer.theta_tgt.cpu() - ref.theta_tgt.cpu()).abs().max()), float(th.abs().max())))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_peer_exchange_matches_allreduce():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29617, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    hung = [p for p in procs if p.is_alive()]
    for p in hung:
        p.kill()
    assert not hung, "a rank did not finish (hang in the peer exchange?)"
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = [out.get(timeout=10) for _ in range(2)]
    for rank, same, d_theta, d_tgt, scale in res:
        assert same, f"rank {rank}: parameters differ across ranks"
        assert d_theta <= 1e-6 and d_tgt <= 1e-6, (rank, d_theta, d_tgt, scale)
