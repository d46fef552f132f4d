"""CPU: the LBF oracle (oracle/lbf_oracle.c, oracle/lbf_ref.py) against hand-computed known answers, the
published Philox4x32-10 vectors, and each other.  No GPU, no product code."""
import numpy as np
import pytest

from oracle import lbf_c, lbf_ref, policy_ref
from tests.lbf_kats import KATS, expected, materialise

# Random123 kat_vectors for philox4x32-10
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_known_answers(ctr, key, want):
    assert lbf_c.philox(ctr, key) == want
    assert tuple(lbf_ref.philox4x32_10(ctr, key)) == want
    got = policy_ref.philox_np(*[np.array([c]) for c in ctr], *key)
    assert tuple(int(g[0]) for g in got) == want


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_kat_c_oracle(kat):
    cfgkw, field, players, step = materialise(kat)
    env = lbf_c.OracleVecEnv(lbf_c.make_cfg(**cfgkw), 1, seed=0)
    env.set_state(field[None], players[None], np.array([step]))
    obs, rew, done, trunc, fret, flen = env.step(np.array([kat["actions"]]), autoreset=False)
    pa, want_rew, want_raw = expected(kat, cfgkw)
    assert np.array_equal(env.players[0], pa)
    want_field = field.copy().reshape(cfgkw["rows"], cfgkw["cols"])
    for r, c in kat["removed"]:
        want_field[r, c] = 0
    assert np.array_equal(env.field[0], want_field.reshape(-1))
    assert np.array_equal(rew[0], want_rew)
    assert np.array_equal(env.ep_return[0], want_raw)
    assert bool(done[0]) == kat["done"] and bool(trunc[0]) == kat["trunc"]
    if "obs_after" in kat:
        assert np.array_equal(obs[0], np.array(kat["obs_after"], np.float32))
    if kat["done"] or kat["trunc"]:
        assert flen[0] == 1 and np.array_equal(fret[0], want_raw) and env.active[0] == 0


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_kat_python_restatement(kat):
    cfgkw, field, players, step = materialise(kat)
    cfg = lbf_ref.LBFConfig(**cfgkw)
    w = lbf_ref.WrappedForaging(cfg, seed=0)
    w.env.load(field, players, step, int(field.astype(np.int32).sum()))
    obs, rew, done, trunc, info = w.step(kat["actions"])
    pa, want_rew, _ = expected(kat, cfgkw)
    _, got_players = w.env.export()
    assert np.array_equal(got_players, pa)
    assert np.array_equal(np.array(rew, np.float64).astype(np.float32), want_rew)
    assert done == kat["done"] and trunc == kat["trunc"]
    if "obs_after" in kat:
        assert np.array_equal(np.stack(obs), np.array(kat["obs_after"], np.float32))


CONFIGS = [
    dict(),
    dict(rows=15, cols=15, n_agents=4, max_num_food=5, sight=15, cooperative_reward=1),
    dict(rows=10, cols=10, n_agents=3, max_num_food=4, sight=2, penalty=0.1, force_coop=1),
    dict(rows=5, cols=5, n_agents=5, max_num_food=2, sight=5, normalize_reward=0),
    dict(observe_id=1, standardise_rewards=1),                                                        # ObserveID + StandardiseReward (wrappers.py:75-141)
    dict(rows=6, cols=6, n_agents=4, max_num_food=3, sight=6, upstream_reset=1),                      # upstream's reset details on a crowded board
    dict(rows=10, cols=10, n_agents=3, max_num_food=4, sight=2, penalty=0.1, standardise_rewards=1, cooperative_reward=1, observe_id=1),
]


@pytest.mark.parametrize("cfgkw", CONFIGS, ids=["8x8-2p-3f", "15x15-4p-5f-coopreward", "10x10-3p-4f-2s-coop-pen", "5x5-5p-2f-raw", "8x8-2p-3f-id-stdrew", "6x6-4p-3f-upstream-reset", "10x10-3p-4f-2s-pen-id-stdrew-coopreward"])
def test_c_oracle_matches_python_restatement_on_random_rollouts(cfgkw):
    rng = np.random.default_rng(7)
    ccfg, pcfg, E = lbf_c.make_cfg(**cfgkw), lbf_ref.LBFConfig(**cfgkw), 24
    venv = lbf_c.OracleVecEnv(ccfg, E, seed=99, env_gid0=5)
    obs = venv.reset()
    penvs = [lbf_ref.WrappedForaging(pcfg, 99, 5 + e) for e in range(E)]
    for e, p in enumerate(penvs):
        assert np.array_equal(np.stack(p.reset()[0]), obs[e])
    for _ in range(60):
        acts = rng.integers(0, 6, size=(E, ccfg.n_agents))
        acts[rng.random(acts.shape) < 0.3] = 5
        obs, rew, done, trunc, fret, flen = venv.step(acts, autoreset=True)
        for e, p in enumerate(penvs):
            o, r, d, tr, info = p.step(acts[e])
            assert d == done[e] and tr == trunc[e]
            assert np.array_equal(np.array(r, dtype=np.float32), rew[e])
            if d or tr:
                assert np.array_equal(info["episode_returns"], fret[e]) and info["episode_length"] == flen[e]
                o, _ = p.reset()
            assert np.array_equal(np.stack(o), obs[e])


def test_reset_is_a_pure_function_of_seed_env_and_episode():
    cfg = lbf_c.make_cfg()
    a = lbf_c.OracleVecEnv(cfg, 8, seed=3, env_gid0=0)
    b = lbf_c.OracleVecEnv(cfg, 4, seed=3, env_gid0=4)  # a shard of the same global env ids
    oa, ob = a.reset(), b.reset()
    assert np.array_equal(oa[4:], ob)
    assert not np.array_equal(oa[0], oa[1])
    # spawn invariants: foods on interior cells, none adjacent (8-neighbourhood), levels within bounds
    f = a.field.reshape(8, 8, 8)
    for e in range(8):
        rr, cc = np.nonzero(f[e])
        assert len(rr) == 3 and rr.min() >= 1 and rr.max() <= 6 and cc.min() >= 1 and cc.max() <= 6
        lv = sorted(a.players[e, :, 2])
        assert f[e].max() <= sum(lv[:3]) and a.food_spawned[e] == f[e].sum()
        for i in range(3):
            for j in range(i + 1, 3):
                assert max(abs(rr[i] - rr[j]), abs(cc[i] - cc[j])) > 1


def test_upstream_reset_flag_changes_the_draw_stream_and_blocks_stale_cells():
    """marl_lbf_cfg.upstream_reset: the permutation draws shift the spawn stream (boards differ from the default from the first reset on), and from the
    second reset on no player is placed on a cell that a not yet re-placed player still occupies from the previous episode."""
    kw = dict(rows=5, cols=5, n_agents=5, max_num_food=1, sight=5)
    a = lbf_c.OracleVecEnv(lbf_c.make_cfg(**kw), 256, 5)
    b = lbf_c.OracleVecEnv(lbf_c.make_cfg(upstream_reset=1, **kw), 256, 5)
    assert not np.array_equal(a.reset(), b.reset())
    hits_default = hits_upstream = 0
    for env, name in ((a, "default"), (b, "upstream")):
        for _ in range(6):
            before = env.players.copy()
            env.reset()
            after = env.players
            for e in range(env.E):
                for i in range(env.N):       # player i was placed while players j > i still stood on their previous cells
                    for j in range(i + 1, env.N):
                        if after[e, i, 0] == before[e, j, 0] and after[e, i, 1] == before[e, j, 1]:
                            if name == "default":
                                hits_default += 1
                            else:
                                hits_upstream += 1
    assert hits_upstream == 0 and hits_default > 0, (hits_default, hits_upstream)
