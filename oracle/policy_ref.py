"""CPU restatement of the action-selection / sampling streams used by the kernels.  TEST INFRASTRUCTURE ONLY.

 * epsilon-greedy: marlbase/dqn/model.py:105-115 -- ONE uniform per env step decides whether the joint action is
   random; random actions are independent uniforms per agent (action_space.sample()), greedy = first argmax.
   The reference draws from Python's unseeded `random`, so the stream itself is ours: Philox4x32-10,
   key (seed_lo, seed_hi ^ TAG), counter (env_gid, episode, t, block).
 * categorical: marlbase/ac/model.py:150-152 `Categorical(logits).sample()` (torch global RNG in the reference);
   here inverse CDF over exp(logit - max) in float32 with a Philox uniform.
 * replay sampling: marlbase/dqn/train.py:95 `np.random.randint(0, len, B)` (with replacement) -> Philox stream.
"""
from __future__ import annotations

import numpy as np

TAG_ACT, TAG_CAT, TAG_SAMPLE = 0x41435430, 0x43415430, 0x53414D50
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)


def philox_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 over numpy arrays (broadcast); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(x, np.uint64) & _LO for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0 & _LO, p1 & _LO, n2 & _LO, p0 & _LO
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in (c0, c1, c2, c3)]


def _u01(u):
    return (u >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def _bounded(u, n):
    return ((u.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int32)


def _keys(seed, tag):
    return seed & 0xFFFFFFFF, ((seed >> 32) & 0xFFFFFFFF) ^ tag


def eps_greedy(values, epsilon, seed, gid, episode, t):
    """values f32[E,N,A]; gid/episode/t integer arrays [E].  Returns int32[E,N]."""
    E, N, A = values.shape
    k0, k1 = _keys(seed, TAG_ACT)
    b0 = philox_np(gid, episode, t, 0, k0, k1)
    explore = np.float32(epsilon) > _u01(b0[0])
    greedy = values.argmax(-1).astype(np.int32)  # first max
    rnd = np.zeros((E, N), np.int32)
    for blk in range((N + 3) // 4):
        words = philox_np(gid, episode, t, 1 + blk, k0, k1)
        for w in range(4):
            i = 4 * blk + w
            if i < N:
                rnd[:, i] = _bounded(words[w], A)
    return np.where(explore[:, None], rnd, greedy)


def categorical_uniforms(seed, gid, episode, t, n_agents):
    k0, k1 = _keys(seed, TAG_CAT)
    E = len(gid)
    u = np.zeros((E, n_agents), np.float32)
    for blk in range((n_agents + 3) // 4):
        words = philox_np(gid, episode, t, blk, k0, k1)
        for w in range(4):
            i = 4 * blk + w
            if i < n_agents:
                u[:, i] = _u01(words[w])
    return u


def categorical(logits, seed, gid, episode, t):
    """Inverse-CDF sample.  Returns (actions int32[E,N], margin f32[E,N]) where margin is the distance of the
    threshold to the nearest CDF edge relative to the total (tests skip exactness when it is ~1 ulp)."""
    E, N, A = logits.shape
    u = categorical_uniforms(seed, gid, episode, t, N)
    m = logits.max(-1, keepdims=True)
    ex = np.exp((logits - m).astype(np.float32)).astype(np.float32)
    cum = np.zeros((E, N), np.float32)
    cdf = np.zeros((E, N, A), np.float32)
    for k in range(A):  # sequential float32 accumulation like the kernel
        cum = (cum + ex[..., k]).astype(np.float32)
        cdf[..., k] = cum
    thresh = (u * cdf[..., -1]).astype(np.float32)
    act = (thresh[..., None] >= cdf).sum(-1).clip(max=A - 1).astype(np.int32)
    margin = np.abs(cdf - thresh[..., None]).min(-1) / cdf[..., -1]
    return act, margin


def replay_sample(seed, update_idx, batch_size, n_valid):
    k0, k1 = _keys(seed, TAG_SAMPLE)
    blocks = np.arange((batch_size + 3) // 4)
    words = philox_np(update_idx & 0xFFFFFFFF, (update_idx >> 32) & 0xFFFFFFFF, blocks, 0, k0, k1)
    flat = np.stack(words, 1).reshape(-1)[:batch_size]
    return _bounded(flat, n_valid)
