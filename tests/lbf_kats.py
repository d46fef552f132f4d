"""Hand-computed known-answer cases for the LBF transition.  Each case:
cfg overrides, field {(r, c): level}, players [(r, c, level)], step before, actions ->
expected players after, removed food cells, raw per-agent rewards (python floats), done, truncated."""
from fractions import Fraction as Fr

KATS = [
    dict(name="head_on_collision_nobody_moves", field={(6, 6): 1}, players=[(2, 2, 1), (2, 4, 1)], actions=[4, 3],
         players_after=[(2, 2, 1), (2, 4, 1)], removed=[], rewards=[0, 0], done=False, trunc=False),
    dict(name="swap_succeeds", field={(6, 6): 1}, players=[(2, 2, 1), (2, 3, 2)], actions=[4, 3],
         players_after=[(2, 3, 1), (2, 2, 2)], removed=[], rewards=[0, 0], done=False, trunc=False),
    dict(name="move_into_stationary_player_fails", field={(6, 6): 1}, players=[(2, 2, 1), (2, 3, 1)], actions=[4, 0],
         players_after=[(2, 2, 1), (2, 3, 1)], removed=[], rewards=[0, 0], done=False, trunc=False),
    dict(name="follower_enters_cell_of_blocked_leader_upstream_quirk", cfg=dict(n_agents=3), field={(6, 6): 1},
         players=[(2, 2, 1), (2, 3, 1), (2, 5, 1)], actions=[4, 4, 3],
         players_after=[(2, 3, 1), (2, 3, 1), (2, 5, 1)], removed=[], rewards=[0, 0, 0], done=False, trunc=False),
    dict(name="cooperative_load", field={(3, 3): 3, (6, 6): 2}, players=[(2, 3, 1), (3, 2, 2)], actions=[5, 5],
         players_after=[(2, 3, 1), (3, 2, 2)], removed=[(3, 3)], rewards=[Fr(3, 15), Fr(6, 15)], done=False, trunc=False,
         obs_after=[[6, 6, 2, -1, -1, 0, -1, -1, 0, 2, 3, 1, 3, 2, 2], [6, 6, 2, -1, -1, 0, -1, -1, 0, 3, 2, 2, 2, 3, 1]]),
    dict(name="cooperative_load_with_cooperative_reward_wrapper", cfg=dict(cooperative_reward=1), field={(3, 3): 3, (6, 6): 2},
         players=[(2, 3, 1), (3, 2, 2)], actions=[5, 5], players_after=[(2, 3, 1), (3, 2, 2)], removed=[(3, 3)],
         rewards=[Fr(3, 15), Fr(6, 15)], done=False, trunc=False),
    dict(name="failed_load_pays_penalty", cfg=dict(penalty=0.1), field={(3, 3): 3}, players=[(2, 3, 1), (6, 6, 2)], actions=[5, 0],
         players_after=[(2, 3, 1), (6, 6, 2)], removed=[], rewards=[-0.1, 0], done=False, trunc=False),
    dict(name="two_adjacent_foods_north_first", field={(2, 3): 1, (3, 4): 1}, players=[(3, 3, 1), (7, 7, 1)], actions=[5, 0],
         players_after=[(3, 3, 1), (7, 7, 1)], removed=[(2, 3)], rewards=[Fr(1, 2), 0], done=False, trunc=False),
    dict(name="last_food_collected_terminates", field={(3, 3): 1}, players=[(3, 2, 1), (7, 7, 2)], actions=[5, 0],
         players_after=[(3, 2, 1), (7, 7, 2)], removed=[(3, 3)], rewards=[1, 0], done=True, trunc=False),
    dict(name="time_limit_truncates", field={(3, 3): 1}, players=[(0, 0, 1), (7, 7, 2)], actions=[0, 0], step=24,
         players_after=[(0, 0, 1), (7, 7, 2)], removed=[], rewards=[0, 0], done=False, trunc=True),
    dict(name="max_episode_steps_terminates", cfg=dict(time_limit=0), field={(3, 3): 1}, players=[(0, 0, 1), (7, 7, 2)], actions=[0, 0], step=49,
         players_after=[(0, 0, 1), (7, 7, 2)], removed=[], rewards=[0, 0], done=True, trunc=False),
    dict(name="move_into_food_is_invalid", field={(3, 3): 1}, players=[(3, 2, 1), (7, 7, 2)], actions=[4, 0],
         players_after=[(3, 2, 1), (7, 7, 2)], removed=[], rewards=[0, 0], done=False, trunc=False),
    dict(name="out_of_bounds_moves_are_invalid", field={(3, 3): 1}, players=[(0, 0, 1), (7, 7, 2)], actions=[1, 4],
         players_after=[(0, 0, 1), (7, 7, 2)], removed=[], rewards=[0, 0], done=False, trunc=False),
    dict(name="load_without_adjacent_food_is_noop", field={(3, 3): 1}, players=[(5, 5, 1), (7, 7, 2)], actions=[5, 5],
         players_after=[(5, 5, 1), (7, 7, 2)], removed=[], rewards=[0, 0], done=False, trunc=False),
    dict(name="diagonal_is_not_adjacent", field={(3, 3): 1}, players=[(2, 2, 1), (7, 7, 2)], actions=[5, 0],
         players_after=[(2, 2, 1), (7, 7, 2)], removed=[], rewards=[0, 0], done=False, trunc=False),
    dict(name="three_players_two_load_one_idle", cfg=dict(n_agents=3), field={(3, 3): 2, (6, 6): 1},
         players=[(2, 3, 1), (3, 2, 1), (5, 5, 2)], actions=[5, 5, 0], players_after=[(2, 3, 1), (3, 2, 1), (5, 5, 2)], removed=[(3, 3)],
         rewards=[Fr(2, 6), Fr(2, 6), 0], done=False, trunc=False),
    dict(name="adjacent_non_loader_does_not_help", field={(3, 3): 2, (6, 6): 1}, players=[(2, 3, 1), (3, 2, 1)], actions=[5, 0],
         players_after=[(2, 3, 1), (3, 2, 1)], removed=[], rewards=[0, 0], done=False, trunc=False),
    dict(name="unnormalised_reward", cfg=dict(normalize_reward=0), field={(3, 3): 2, (6, 6): 1}, players=[(2, 3, 2), (0, 0, 1)], actions=[5, 0],
         players_after=[(2, 3, 2), (0, 0, 1)], removed=[(3, 3)], rewards=[4, 0], done=False, trunc=False),
    dict(name="partial_sight_window_and_edge_quirk", cfg=dict(rows=10, cols=10, sight=2, max_num_food=2), field={(4, 6): 1, (1, 1): 2},
         players=[(5, 5, 1), (0, 1, 2)], actions=[0, 0], players_after=[(5, 5, 1), (0, 1, 2)], removed=[], rewards=[0, 0], done=False, trunc=False,
         # agent0 window rows 3..7, cols 3..7: food (4,6)->(1,3); agent1 at (0,1) is not visible.
         # agent1 window rows 0..2, cols 0..3: food (1,1)->(1,1); agent0 transforms to (5,5-1+1)=(5,5): 5 > 4 -> hidden.
         obs_after=[[1, 3, 1, -1, -1, 0, 2, 2, 1, -1, -1, 0], [1, 1, 2, -1, -1, 0, 0, 1, 2, -1, -1, 0]]),
    dict(name="partial_sight_sees_player_outside_window_near_edge", cfg=dict(rows=10, cols=10, sight=2, max_num_food=1), field={(8, 8): 1},
         players=[(0, 1, 1), (3, 4, 2)], actions=[0, 0], players_after=[(0, 1, 1), (3, 4, 2)], removed=[], rewards=[0, 0], done=False, trunc=False,
         # agent0 at (0,1): transform = pos - centre + min(sight, centre) -> agent1 (3,4) -> (3, 4): max 4 <= 2*sight -> "seen" although
         # row 3 is outside its field window (upstream checks only the transformed coordinates).
         # agent1 at (3,4): agent0 -> (0-3+2, 1-4+2) = (-1,-1) -> hidden.
         obs_after=[[-1, -1, 0, 0, 1, 1, 3, 4, 2], [-1, -1, 0, 2, 2, 2, -1, -1, 0]]),
]


def materialise(kat):
    """-> (cfg kwargs, field int8[R*C], players int8[N,4], step)"""
    import numpy as np

    cfg = dict(rows=8, cols=8, n_agents=2, max_num_food=3, sight=8, time_limit=25)
    cfg.update(kat.get("cfg", {}))
    if "sight" not in kat.get("cfg", {}):
        cfg["sight"] = cfg["rows"]
    R, C, N = cfg["rows"], cfg["cols"], cfg["n_agents"]
    field = np.zeros((R, C), np.int8)
    for (r, c), lvl in kat["field"].items():
        field[r, c] = lvl
    players = np.zeros((N, 4), np.int8)
    for i, (r, c, lvl) in enumerate(kat["players"]):
        players[i, :3] = (r, c, lvl)
    return cfg, field.reshape(-1), players, kat.get("step", 0)


def expected(kat, cfg):
    import numpy as np

    raw = [float(x) for x in kat["rewards"]]
    if cfg.get("cooperative_reward"):
        rew = [sum(raw)] * len(raw)
    else:
        rew = raw
    pa = np.zeros((len(kat["players_after"]), 4), np.int8)
    for i, (r, c, lvl) in enumerate(kat["players_after"]):
        pa[i, :3] = (r, c, lvl)
    return pa, np.array(rew, np.float64).astype(np.float32), np.array(raw, np.float64).astype(np.float32)
