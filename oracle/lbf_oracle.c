/*
 * lbf_oracle.c -- plain-C CPU restatement of lbforaging's ForagingEnv (reset / step / observation) and
 * of the marlbase wrapper stack applied on top of it.  See lbf_oracle.h for provenance and the
 * "parity unpinned" statement.  TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Algorithm sources (restated, not copied):
 *   - lbforaging (uoe-agents/lb-foraging, `-v3` ids): ForagingEnv.step / reset / _make_gym_obs,
 *     spawn_players / spawn_food, _is_valid_action, adjacent_food(_location), adjacent_players.
 *     Entered by the reference at marlbase/dqn/train.py:203,217 and marlbase/ac/train.py:30,79-81.
 *   - gymnasium<1.0 TimeLimit: truncated = elapsed_steps >= time_limit  (marlbase/utils/envs.py:95-96).
 *   - RecordEpisodeStatistics: float32 per-agent return accumulation (marlbase/utils/wrappers.py:31-45).
 *   - CooperativeReward: N * [sum(reward)] in python floats           (marlbase/utils/wrappers.py:106-108).
 *
 * Random numbers: upstream draws from gymnasium's PCG64; that stream is not reproducible on a GPU, so
 * spawns use Philox4x32-10 keyed by (seed, env id, episode index) -- the same counter scheme the CUDA
 * kernel implements independently in codebase_b200/csrc.  Documented deviations from upstream reset:
 * the level-bound permutations (no-ops for scalar bounds) draw nothing, and stale previous-episode
 * positions of not-yet-placed players do not block a cell.
 */
#include "lbf_oracle.h"
#include <string.h>
#include <stdlib.h>

#define MAX_AGENTS 32
#define TAG_RESET 0x52455345u /* "RESE" */

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

void lbf_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Sequential draw stream for one (env, episode) reset. */
typedef struct { uint32_t key[2]; uint32_t gid, ep; uint32_t n; uint32_t buf[4]; } draw_stream;

static void ds_init(draw_stream* d, uint64_t seed, uint32_t gid, uint32_t ep) {
  d->key[0] = (uint32_t)seed; d->key[1] = (uint32_t)(seed >> 32) ^ TAG_RESET;
  d->gid = gid; d->ep = ep; d->n = 0;
}
static uint32_t ds_next(draw_stream* d) {
  if ((d->n & 3u) == 0) {
    uint32_t ctr[4] = {d->gid, d->ep, d->n >> 2, 0};
    lbf_oracle_philox(ctr, d->key, d->buf);
  }
  return d->buf[d->n++ & 3u];
}
/* integer in [lo, hi) -- multiply-shift, the same mapping the kernel uses */
static int ds_randint(draw_stream* d, int lo, int hi) { return lo + (int)mulhi32(ds_next(d), (uint32_t)(hi - lo)); }

int lbf_oracle_obs_dim(const lbf_oracle_cfg* c) { return 3 * c->max_num_food + 3 * c->n_agents + (c->observe_id ? c->n_agents : 0); }

#define F(r_, c_) field[(r_) * C + (c_)]

static int is_empty(const lbf_oracle_cfg* cfg, const int8_t* field, const int8_t* players, int placed, int r, int cc) {
  int C = cfg->cols;
  if (F(r, cc) != 0) return 0;
  /* upstream_reset: upstream's _is_empty_location looks at EVERY player that has a position -- reset() never clears positions, so players not yet
   * re-placed still block their previous episode's cell (level byte 0 = no position yet: first reset) */
  const int n = cfg->upstream_reset ? cfg->n_agents : placed;
  for (int j = 0; j < n; ++j)
    if (players[4 * j] == r && players[4 * j + 1] == cc && (!cfg->upstream_reset || players[4 * j + 2] != 0)) return 0;
  return 1;
}

void lbf_oracle_reset_one(const lbf_oracle_cfg* cfg, uint64_t seed, uint32_t env_gid, uint32_t episode_idx,
                          int8_t* field, int8_t* players, int32_t* food_spawned) {
  const int R = cfg->rows, C = cfg->cols, N = cfg->n_agents;
  draw_stream ds; ds_init(&ds, seed, env_gid, episode_idx);
  memset(field, 0, (size_t)R * C);
  if (!cfg->upstream_reset) memset(players, 0, (size_t)N * 4);
  else for (int k = N - 1; k >= 1; --k) (void)ds_randint(&ds, 0, k + 1); /* spawn_players: np_random.permutation over the level bounds */
  /* spawn_players: uniform empty cell (<=1000 attempts), then level ~ U[min, max] */
  for (int i = 0; i < N; ++i) {
    int placed = 0;
    for (int attempts = 0; attempts < 1000 && !placed; ++attempts) {
      int r = ds_randint(&ds, 0, R), cc = ds_randint(&ds, 0, C);
      if (is_empty(cfg, field, players, i, r, cc)) {
        players[4 * i] = (int8_t)r; players[4 * i + 1] = (int8_t)cc;
        players[4 * i + 2] = (int8_t)ds_randint(&ds, cfg->min_player_level, cfg->max_player_level + 1);
        placed = 1;
      }
    }
    if (!placed) { /* never reached for sane sizes; deterministic fallback: first empty cell, min level */
      for (int p = 0; p < R * C && !placed; ++p)
        if (is_empty(cfg, field, players, i, p / C, p % C)) {
          players[4 * i] = (int8_t)(p / C); players[4 * i + 1] = (int8_t)(p % C);
          players[4 * i + 2] = (int8_t)cfg->min_player_level; placed = 1;
        }
    }
  }
  /* max food level: sum of the three lowest player levels unless configured */
  int max_lvl = cfg->max_food_level;
  if (max_lvl <= 0) {
    int lv[MAX_AGENTS];
    for (int i = 0; i < N; ++i) lv[i] = players[4 * i + 2];
    for (int i = 1; i < N; ++i) { int v = lv[i], j = i - 1; while (j >= 0 && lv[j] > v) { lv[j + 1] = lv[j]; --j; } lv[j + 1] = v; }
    max_lvl = 0;
    for (int i = 0; i < N && i < 3; ++i) max_lvl += lv[i];
  }
  int min_lvl = cfg->force_coop ? max_lvl : cfg->min_food_level;
  if (cfg->upstream_reset) for (int k = cfg->max_num_food - 1; k >= 1; --k) (void)ds_randint(&ds, 0, k + 1); /* spawn_food: permutation over the food level bounds */
  /* spawn_food: interior cells, nothing in the 3x3 neighbourhood, nothing within 2 along row/col, empty */
  int count = 0;
  for (int attempts = 0; count < cfg->max_num_food && attempts < 1000; ++attempts) {
    int r = ds_randint(&ds, 1, R - 1), cc = ds_randint(&ds, 1, C - 1);
    int sum3 = 0, cross = 0;
    for (int rr = (r - 1 < 0 ? 0 : r - 1); rr < (r + 2 > R ? R : r + 2); ++rr)
      for (int c2 = (cc - 1 < 0 ? 0 : cc - 1); c2 < (cc + 2 > C ? C : cc + 2); ++c2) sum3 += F(rr, c2);
    for (int rr = (r - 2 < 0 ? 0 : r - 2); rr < (r + 3 > R ? R : r + 3); ++rr) cross += F(rr, cc);
    for (int c2 = (cc - 2 < 0 ? 0 : cc - 2); c2 < (cc + 3 > C ? C : cc + 3); ++c2) cross += F(r, c2);
    if (sum3 > 0 || cross > 0 || !is_empty(cfg, field, players, N, r, cc)) continue;
    F(r, cc) = (int8_t)(min_lvl == max_lvl ? min_lvl : ds_randint(&ds, min_lvl, max_lvl + 1));
    ++count;
  }
  int s = 0;
  for (int p = 0; p < R * C; ++p) s += field[p];
  *food_spawned = s;
}

static int adjacent_food_sum(const int8_t* field, int R, int C, int r, int cc) {
  int up = r - 1 < 0 ? 0 : r - 1, dn = r + 1 > R - 1 ? R - 1 : r + 1;
  int lf = cc - 1 < 0 ? 0 : cc - 1, rt = cc + 1 > C - 1 ? C - 1 : cc + 1;
  return F(up, cc) + F(dn, cc) + F(r, lf) + F(r, rt);
}

/* upstream's adjacent_food_location, including its `row > 1` / `col > 1` guards */
static int adjacent_food_loc(const int8_t* field, int R, int C, int r, int cc, int* fr, int* fc) {
  if (r > 1 && F(r - 1, cc) > 0) { *fr = r - 1; *fc = cc; return 1; }
  if (r < R - 1 && F(r + 1, cc) > 0) { *fr = r + 1; *fc = cc; return 1; }
  if (cc > 1 && F(r, cc - 1) > 0) { *fr = r; *fc = cc - 1; return 1; }
  if (cc < C - 1 && F(r, cc + 1) > 0) { *fr = r; *fc = cc + 1; return 1; }
  return 0;
}

void lbf_oracle_step_one(const lbf_oracle_cfg* cfg, int8_t* field, int8_t* players, int32_t* step,
                         int32_t food_spawned, const int32_t* actions, double* rewards_raw,
                         int32_t* done, int32_t* truncated) {
  const int R = cfg->rows, C = cfg->cols, N = cfg->n_agents;
  int act[MAX_AGENTS], tr[MAX_AGENTS], tc[MAX_AGENTS], pending[MAX_AGENTS];
  *step += 1;
  for (int i = 0; i < N; ++i) rewards_raw[i] = 0.0;
  /* invalid actions become NONE; validity is judged on the pre-step field (other players not checked) */
  for (int i = 0; i < N; ++i) {
    int r = players[4 * i], cc = players[4 * i + 1], a = actions[i], ok;
    switch (a) {
      case 0: ok = 1; break;
      case 1: ok = r > 0 && F(r - 1, cc) == 0; break;
      case 2: ok = r < R - 1 && F(r + 1, cc) == 0; break;
      case 3: ok = cc > 0 && F(r, cc - 1) == 0; break;
      case 4: ok = cc < C - 1 && F(r, cc + 1) == 0; break;
      case 5: ok = adjacent_food_sum(field, R, C, r, cc) > 0; break;
      default: ok = 0;
    }
    act[i] = ok ? a : 0;
    tr[i] = r + (act[i] == 2) - (act[i] == 1);
    tc[i] = cc + (act[i] == 4) - (act[i] == 3);
    pending[i] = act[i] == 5;
  }
  /* a cell proposed by more than one player is entered by nobody; sole proposers move */
  int moved_r[MAX_AGENTS], moved_c[MAX_AGENTS];
  for (int i = 0; i < N; ++i) {
    int cnt = 0;
    for (int j = 0; j < N; ++j) cnt += (tr[j] == tr[i] && tc[j] == tc[i]);
    moved_r[i] = cnt == 1 ? tr[i] : players[4 * i];
    moved_c[i] = cnt == 1 ? tc[i] : players[4 * i + 1];
  }
  for (int i = 0; i < N; ++i) { players[4 * i] = (int8_t)moved_r[i]; players[4 * i + 1] = (int8_t)moved_c[i]; }
  /* loading, ascending agent index (upstream pops a python set; order fixed here) */
  for (int i = 0; i < N; ++i) {
    if (!pending[i]) continue;
    int fr, fc;
    if (!adjacent_food_loc(field, R, C, players[4 * i], players[4 * i + 1], &fr, &fc)) { pending[i] = 0; continue; }
    int food = F(fr, fc), lvl_sum = 0, adj[MAX_AGENTS];
    for (int j = 0; j < N; ++j) {
      int rj = players[4 * j], cj = players[4 * j + 1];
      int near = (abs(rj - fr) == 1 && cj == fc) || (abs(cj - fc) == 1 && rj == fr);
      adj[j] = near && (pending[j] || j == i);
      if (adj[j]) lvl_sum += players[4 * j + 2];
    }
    for (int j = 0; j < N; ++j) if (adj[j]) pending[j] = 0;
    if (lvl_sum < food) {
      for (int j = 0; j < N; ++j) if (adj[j]) rewards_raw[j] -= cfg->penalty;
      continue;
    }
    for (int j = 0; j < N; ++j) if (adj[j]) {
      double rw = (double)(players[4 * j + 2] * food);
      if (cfg->normalize_reward) rw = rw / (double)(lvl_sum * food_spawned);
      rewards_raw[j] = rw;
    }
    F(fr, fc) = 0;
  }
  int left = 0;
  for (int p = 0; p < R * C; ++p) left += field[p];
  *done = (left == 0) || (cfg->max_episode_steps <= *step);
  *truncated = cfg->time_limit > 0 && *step >= cfg->time_limit;
}

void lbf_oracle_obs_one(const lbf_oracle_cfg* cfg, const int8_t* field, const int8_t* players, int agent, float* out) {
  if (cfg->observe_id) { /* ObserveID.observation (wrappers.py:96-103): np.eye(n_agents) concatenated in front */
    for (int j = 0; j < cfg->n_agents; ++j) out[j] = j == agent ? 1.f : 0.f;
    out += cfg->n_agents;
  }
  const int R = cfg->rows, C = cfg->cols, N = cfg->n_agents, S = cfg->sight, NF = cfg->max_num_food;
  const int pr = players[4 * agent], pc = players[4 * agent + 1];
  const int r0 = pr - S < 0 ? 0 : pr - S, r1 = pr + S + 1 > R ? R : pr + S + 1;
  const int c0 = pc - S < 0 ? 0 : pc - S, c1 = pc + S + 1 > C ? C : pc + S + 1;
  for (int i = 0; i < NF + N; ++i) { out[3 * i] = -1.f; out[3 * i + 1] = -1.f; out[3 * i + 2] = 0.f; }
  int k = 0;
  for (int r = r0; r < r1; ++r)
    for (int cc = c0; cc < c1; ++cc)
      if (F(r, cc) != 0 && k < NF) { out[3 * k] = (float)(r - r0); out[3 * k + 1] = (float)(cc - c0); out[3 * k + 2] = (float)F(r, cc); ++k; }
  /* players: self first, then the others in index order; visible iff both transformed coords in [0, 2*sight] */
  const int orow = pr - (S < pr ? S : pr), ocol = pc - (S < pc ? S : pc);
  int slot = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int j = 0; j < N; ++j) {
      if ((pass == 0) != (j == agent)) continue;
      int y = players[4 * j] - orow, x = players[4 * j + 1] - ocol;
      int lo = y < x ? y : x, hi = y > x ? y : x;
      if (lo < 0 || hi > 2 * S) continue;
      float* o = out + 3 * NF + 3 * slot++;
      o[0] = (float)y; o[1] = (float)x; o[2] = (float)players[4 * j + 2];
    }
}

static void write_obs(const lbf_oracle_cfg* c, const lbf_oracle_state* s, int e, float* obs_out) {
  if (!obs_out) return;
  const int N = c->n_agents, D = lbf_oracle_obs_dim(c), RC = c->rows * c->cols;
  for (int i = 0; i < N; ++i)
    lbf_oracle_obs_one(c, s->field + (size_t)e * RC, s->players + (size_t)e * N * 4, i, obs_out + ((size_t)e * N + i) * D);
}

static void reset_env(const lbf_oracle_cfg* c, uint64_t seed, uint32_t gid, lbf_oracle_state* s, int e) {
  const int N = c->n_agents, RC = c->rows * c->cols;
  lbf_oracle_reset_one(c, seed, gid, s->episode_idx[e], s->field + (size_t)e * RC, s->players + (size_t)e * N * 4, &s->food_spawned[e]);
  s->episode_idx[e] += 1;
  s->step[e] = 0; s->ep_len[e] = 0; s->active[e] = 1;
  for (int i = 0; i < N; ++i) s->ep_return[(size_t)e * N + i] = 0.f;
}

void lbf_oracle_reset(const lbf_oracle_cfg* c, int32_t n_envs, uint64_t seed, uint32_t env_gid0,
                      lbf_oracle_state* s, const uint8_t* reset_mask, float* obs_out) {
  for (int e = 0; e < n_envs; ++e) {
    if (reset_mask && !reset_mask[e]) { write_obs(c, s, e, obs_out); continue; }
    reset_env(c, seed, env_gid0 + (uint32_t)e, s, e);
    write_obs(c, s, e, obs_out);
  }
}

void lbf_oracle_step(const lbf_oracle_cfg* c, int32_t n_envs, uint64_t seed, uint32_t env_gid0,
                     lbf_oracle_state* s, const int32_t* actions, float* obs_out, float* rew_out,
                     uint8_t* done_out, uint8_t* trunc_out, float* final_ret_out, int32_t* final_len_out,
                     int32_t autoreset) {
  const int N = c->n_agents, RC = c->rows * c->cols;
  for (int e = 0; e < n_envs; ++e) {
    if (!s->active[e]) { /* frozen after its episode ended (episode-synchronous collection) */
      for (int i = 0; i < N; ++i) rew_out[(size_t)e * N + i] = 0.f;
      done_out[e] = 1; trunc_out[e] = 0;
      write_obs(c, s, e, obs_out);
      continue;
    }
    double raw[MAX_AGENTS]; int32_t done, trunc;
    lbf_oracle_step_one(c, s->field + (size_t)e * RC, s->players + (size_t)e * N * 4, &s->step[e], s->food_spawned[e],
                        actions + (size_t)e * N, raw, &done, &trunc);
    /* RecordEpisodeStatistics sits inside CooperativeReward (envs.py:97-109): it sees raw rewards, as float32 */
    for (int i = 0; i < N; ++i) s->ep_return[(size_t)e * N + i] += (float)raw[i];
    s->ep_len[e] += 1;
    if (c->standardise_rewards) { /* StandardiseReward.reward (wrappers.py:119-141) with numpy's types: f32 state, f64 where the reward list enters */
      float* st = s->stdr + (size_t)e * (2 * N + 1);
      const float sumw = st[2 * N], temp_sumw = sumw + 1.0f;
      const int n = s->stdr_n[e] + 1;
      for (int i = 0; i < N; ++i) {
        const double q = raw[i] - (double)st[i];
        const double r = q / (double)temp_sumw;
        st[i] = (float)((double)st[i] + r);
        const double qr = q * r;
        st[N + i] = (float)((double)st[N + i] + qr * (double)sumw);
        if (n > 1) {
          const float num = st[N + i] * (float)n, den = temp_sumw * (float)(n - 1);
          const float var = num / den;
          raw[i] = (raw[i] - (double)st[i]) / (double)(sqrtf(var) + 1e-6f);
        }
      }
      st[2 * N] = temp_sumw; s->stdr_n[e] = n;
    }
    double tot = 0.0;
    for (int i = 0; i < N; ++i) tot += raw[i];
    for (int i = 0; i < N; ++i) rew_out[(size_t)e * N + i] = (float)(c->cooperative_reward ? tot : raw[i]);
    done_out[e] = (uint8_t)done; trunc_out[e] = (uint8_t)trunc;
    if (done || trunc) {
      if (final_ret_out) for (int i = 0; i < N; ++i) final_ret_out[(size_t)e * N + i] = s->ep_return[(size_t)e * N + i];
      if (final_len_out) final_len_out[e] = s->ep_len[e];
      if (autoreset) reset_env(c, seed, env_gid0 + (uint32_t)e, s, e);
      else s->active[e] = 0;
    }
    write_obs(c, s, e, obs_out);
  }
}
