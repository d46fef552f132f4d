"""ctypes binding of oracle/liblbf_oracle.so (the plain-C LBF restatement).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleCfg(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32), ("n_agents", C.c_int32), ("max_num_food", C.c_int32),
        ("sight", C.c_int32), ("min_player_level", C.c_int32), ("max_player_level", C.c_int32),
        ("min_food_level", C.c_int32), ("max_food_level", C.c_int32), ("max_episode_steps", C.c_int32),
        ("time_limit", C.c_int32), ("force_coop", C.c_int32), ("normalize_reward", C.c_int32),
        ("cooperative_reward", C.c_int32), ("penalty", C.c_double), ("observe_id", C.c_int32), ("standardise_rewards", C.c_int32), ("upstream_reset", C.c_int32),
    ]


class _State(C.Structure):
    _fields_ = [
        ("field", C.c_void_p), ("players", C.c_void_p), ("step", C.c_void_p), ("food_spawned", C.c_void_p),
        ("ep_return", C.c_void_p), ("ep_len", C.c_void_p), ("episode_idx", C.c_void_p), ("active", C.c_void_p),
        ("stdr", C.c_void_p), ("stdr_n", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liblbf_oracle.so")
    src = os.path.join(_HERE, "lbf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liblbf_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.lbf_oracle_obs_dim.restype = C.c_int
    return _LIB


def make_cfg(**kw) -> OracleCfg:
    d = dict(rows=8, cols=8, n_agents=2, max_num_food=3, sight=8, min_player_level=1, max_player_level=2,
             min_food_level=1, max_food_level=0, max_episode_steps=50, time_limit=25, force_coop=0,
             normalize_reward=1, cooperative_reward=0, penalty=0.0, observe_id=0, standardise_rewards=0, upstream_reset=0)
    d.update(kw)
    return OracleCfg(**d)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().lbf_oracle_philox(c, k, o)
    return tuple(o)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleVecEnv:
    """E envs held in numpy arrays with the same layout as the device state of marl_lbf."""

    def __init__(self, cfg: OracleCfg, n_envs: int, seed: int, env_gid0: int = 0):
        self.cfg, self.E, self.seed, self.gid0 = cfg, n_envs, seed, env_gid0
        N, RC = cfg.n_agents, cfg.rows * cfg.cols
        self.N, self.D = N, 3 * cfg.max_num_food + 3 * N + (N if cfg.observe_id else 0)
        self.field = np.zeros((n_envs, RC), np.int8)
        self.players = np.zeros((n_envs, N, 4), np.int8)
        self.step_count = np.zeros(n_envs, np.int32)
        self.food_spawned = np.zeros(n_envs, np.int32)
        self.ep_return = np.zeros((n_envs, N), np.float32)
        self.ep_len = np.zeros(n_envs, np.int32)
        self.episode_idx = np.zeros(n_envs, np.uint32)
        self.active = np.zeros(n_envs, np.uint8)
        self.stdr = np.zeros((n_envs, 2 * N + 1), np.float32)
        self.stdr_n = np.zeros(n_envs, np.int32)
        self._st = _State(_p(self.field), _p(self.players), _p(self.step_count), _p(self.food_spawned),
                          _p(self.ep_return), _p(self.ep_len), _p(self.episode_idx), _p(self.active), _p(self.stdr), _p(self.stdr_n))

    def reset(self, mask=None):
        obs = np.zeros((self.E, self.N, self.D), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().lbf_oracle_reset(C.byref(self.cfg), C.c_int32(self.E), C.c_uint64(self.seed), C.c_uint32(self.gid0),
                               C.byref(self._st), _p(m), _p(obs))
        return obs

    def step(self, actions, autoreset=False):
        a = np.ascontiguousarray(actions, np.int32).reshape(self.E, self.N)
        obs = np.zeros((self.E, self.N, self.D), np.float32)
        rew = np.zeros((self.E, self.N), np.float32)
        done = np.zeros(self.E, np.uint8)
        trunc = np.zeros(self.E, np.uint8)
        fret = np.zeros((self.E, self.N), np.float32)
        flen = np.zeros(self.E, np.int32)
        lib().lbf_oracle_step(C.byref(self.cfg), C.c_int32(self.E), C.c_uint64(self.seed), C.c_uint32(self.gid0),
                              C.byref(self._st), _p(a), _p(obs), _p(rew), _p(done), _p(trunc), _p(fret), _p(flen),
                              C.c_int32(int(autoreset)))
        return obs, rew, done, trunc, fret, flen

    def set_state(self, field, players, step, food_spawned=None):
        self.field[:] = np.asarray(field, np.int8).reshape(self.field.shape)
        self.players[:] = np.asarray(players, np.int8).reshape(self.players.shape)
        self.step_count[:] = np.asarray(step, np.int32)
        self.food_spawned[:] = self.field.astype(np.int32).sum(1) if food_spawned is None else np.asarray(food_spawned, np.int32)
        self.ep_return[:] = 0
        self.ep_len[:] = 0
        self.active[:] = 1
