/*
 * gw_oracle.c -- CPU restatement (plain C) of the reference's grid-world step path.
 *
 * TEST INFRASTRUCTURE ONLY.  Linked/loaded only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, as the checker or the timed CPU baseline.
 * The product library (libgridworld_b200.so) does not contain, link or call any of this.
 *
 * It keeps the reference's literal structure -- a growing list of cells per agent, floor/ceil
 * indices recomputed per pass, the five-way elif chain, the fix-point loop -- so that it is an
 * independent formulation from the CUDA kernels (which use packed trajectories and crash masks).
 * Parity pin: tests/test_c_oracle.py checks it against the golden vectors recorded from the
 * reference (the .npz files under tests/golden/) and against the Python restatement (oracle/gridworld_oracle.py).
 *
 * It shares the POD structs of include/gridworld_b200.h (gw_config / gww_config, gw_io with HOST pointers) so
 * that the same inputs can be fed to both sides.  The device-RNG mode (Philox4x32-10 counters,
 * DESIGN.md "RNG") is restated here too, so full-size rollouts can be compared bit for bit.
 *
 * Reference lines (relative to the reference root):
 *   update_world      custom/grid_world.py:424-563  (UpdateGWorld), :233-405 (collisions), :190-209 (revert)
 *   fear_one_actor    custom/Responsibility.py:135-210, :20-54; close list custom/ma_customenv.py:456-464
 *   ma step / reset   custom/ma_customenv.py:169-215, :217-334, :338-452, :467-506
 *   single step/reset custom/customenv.py:78-183, :186-356
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gridworld_b200.h"

#define MAXSTEPS 4 /* GWorld.MaxSteps, grid_world.py:24 */

/* Two builds of this one source (oracle/Makefile): libgw_oracle.so restates the path on the packed layout's config
 * (gw_config: W = 16, H <= 16, <= 4 agents), libgw_oracle_wide.so (-DGWO_WIDE) on the general layout's (gww_config: up to
 * 64 x 64, up to 16 agents).  The algorithm text below is the same; only the array bounds, the cell index of the
 * policy / MdR maps and the cell encoding of the restricted paths differ. */
#ifdef GWO_WIDE
typedef gww_config gwo_config;
#define OR_A GWW_MAX_AGENTS
#define OR_CELLS GWW_MAX_CELLS
#define CELL_IDX(cfg, r, c) ((r) * (cfg)->width + (c))
#define BLK_R(v) ((int)((v) >> 8))
#define BLK_C(v) ((int)((v) & 255))
#else
typedef gw_config gwo_config;
#define OR_A GW_MAX_AGENTS
#define OR_CELLS (GW_MAX_H * GW_W)
#define CELL_IDX(cfg, r, c) ((r) * GW_W + (c))
#define BLK_R(v) ((int)((v) >> 4))
#define BLK_C(v) ((int)((v) & 15))
#endif

typedef struct { int r, c; } cell_t;

static const int MOVE_DR[9] = {0, -1, 1, 0, 0, -1, 1, 0, 0};   /* custom_agent.py:140-150 */
static const int MOVE_DC[9] = {0, 0, 0, -1, 1, 0, 0, -1, 1};
static const int MOVE_LEN[9] = {1, 1, 1, 1, 1, 2, 2, 2, 2};

typedef struct env_state {
  cell_t loc[OR_A];
  int apple_present[GW_MAX_LEARNERS];
  int term[GW_MAX_LEARNERS];
  int trunc;
  int pd_valid[GW_MAX_LEARNERS];
  int pd[GW_MAX_LEARNERS];
  int steps;
  uint32_t tick;
  int ep_ret[GW_MAX_LEARNERS];
} env_state;

typedef struct gwo_handle {
  gwo_config cfg;
  env_state* env;
  uint32_t thr[GW_MAX_POLICIES][2][8];
  cell_t active[OR_CELLS];
  int n_active;
  double resp_lut[10][10];
  int reset_done;
  char err[256];
  /* statistics */
  uint64_t env_steps, episodes, len_sum, crashes, apples, unresolved, fear_nz;
  double return_sum, fear_sum;
  pthread_mutex_t stat_mu;
} gwo_handle;

static int cfg_active(const gwo_config* c, int r, int col) {
  return r >= 0 && r < c->height && col >= 0 && col < c->width && ((c->map_rows[r] >> col) & 1);
}

/* ------------------------------------------------------------------ UpdateGWorld */
typedef struct {
  cell_t loc[OR_A];
  int crash[OR_A], restr[OR_A];
  int caught[2][2];
  int unresolved;
} update_out;

static int cell_eq(cell_t a, cell_t b) { return a.r == b.r && a.c == b.c; }

/* `[old, new] in self.RestrictedPaths` (grid_world.py:498) with the tuple-typed paths the reference intends (:654-665):
 * a linear search of the list, as there */
static int path_restricted(const gwo_config* cfg, cell_t a, cell_t b) {
  for (int k = 0; k < cfg->n_blocked; ++k)
    if (BLK_R(cfg->blocked_from[k]) == a.r && BLK_C(cfg->blocked_from[k]) == a.c && BLK_R(cfg->blocked_to[k]) == b.r &&
        BLK_C(cfg->blocked_to[k]) == b.c)
      return 1;
  return 0;
}

static void update_world(const gwo_config* cfg, int n, const cell_t* loc0, const int* act, const cell_t* apples,
                         const int* apple_on, int n_eaters, update_out* out) {
  cell_t path[OR_A][MAXSTEPS + 1];
  int plen[OR_A];
  int crash[OR_A] = {0}, restr[OR_A] = {0};
  cell_t cur[OR_A];
  memset(out, 0, sizeof(*out));
  for (int i = 0; i < n; ++i) { path[i][0] = loc0[i]; plen[i] = 1; }                 /* :437-439 */
  for (int step = 0; step < MAXSTEPS; ++step) {                                     /* :458 */
    for (int i = 0; i < n; ++i) cur[i] = loc0[i];                                   /* :460 */
    for (int i = 0; i < n; ++i) {                                                   /* :462-518 */
      int dr = 0, dc = 0;
      if (step < MOVE_LEN[act[i]] && !crash[i]) { dr = MOVE_DR[act[i]]; dc = MOVE_DC[act[i]]; }
      cell_t old = path[i][step];
      cell_t nw = {old.r + dr, old.c + dc};
      cell_t cl = nw;                                                               /* np.clip :486-487 */
      if (cl.r < 0) cl.r = 0;
      if (cl.r > cfg->height - 1) cl.r = cfg->height - 1;
      if (cl.c < 0) cl.c = 0;
      if (cl.c > cfg->width - 1) cl.c = cfg->width - 1;
      if (!cell_eq(cl, nw)) { restr[i] = 1; nw = cl; }
      if (cfg_active(cfg, nw.r, nw.c)) {                                            /* :496 */
        if (!path_restricted(cfg, old, nw)) path[i][plen[i]++] = nw;                /* :498-499 */
        else { path[i][plen[i]++] = old; restr[i] = 1; }                            /* :504-507: AgentPath in RestrictedPaths */
      } else { path[i][plen[i]++] = old; restr[i] = 1; }                            /* :512-515 */
    }
    int count = n, loops = 0;                                                       /* :247-248 */
    while (count > 0 && loops < 2 * n) {                                            /* :250 */
      ++loops;
      count = 0;
      for (int ii = 0; ii < n - 1; ++ii) {                                          /* :255 */
        const int qi = (step + 1) * MOVE_LEN[act[ii]];
        const int fi = qi / MAXSTEPS, ci = (qi + MAXSTEPS - 1) / MAXSTEPS;
        cur[ii] = path[ii][fi];                                                     /* :259 */
        for (int jj = ii + 1; jj < n; ++jj) {
          const int qj = (step + 1) * MOVE_LEN[act[jj]];
          const int fj = qj / MAXSTEPS, cj = (qj + MAXSTEPS - 1) / MAXSTEPS;
          cur[jj] = path[jj][fj];                                                   /* :264 */
          const cell_t ai = path[ii][fi], bi = path[ii][ci], aj = path[jj][fj], bj = path[jj][cj];
          int hit = 0;
          if (cell_eq(ai, aj) || cell_eq(bi, bj)) hit = 1;                          /* :276-278 */
          else if (cell_eq(ai, bj) && cell_eq(bi, aj)) hit = 1;                     /* :291-294 */
          else if (cell_eq(ai, bj)) {                                               /* :307-326 */
            const int overhang = ((4 * ci - qi) + (qj - 4 * fj)) <= 4;
            const int same = (bi.r - ai.r == bj.r - aj.r) && (bi.c - ai.c == bj.c - aj.c);
            hit = !(overhang && same);
          } else if (cell_eq(bi, aj)) {                                             /* :339-357 */
            const int overhang = ((4 * cj - qj) + (qi - 4 * fi)) <= 4;
            const int same = (bi.r - ai.r == bj.r - aj.r) && (bi.c - ai.c == bj.c - aj.c);
            hit = !(overhang && same);
          } else if ((cell_eq(ai, loc0[jj]) && cell_eq(loc0[ii], aj)) || (cell_eq(bi, loc0[jj]) && cell_eq(loc0[ii], bj)) ||
                     (cell_eq(ai, loc0[jj]) && cell_eq(loc0[ii], bj)) || (cell_eq(bi, loc0[jj]) && cell_eq(loc0[ii], aj)))
            hit = 1;                                                                /* :371-378 */
          if (hit) { ++count; crash[ii] = 1; crash[jj] = 1; }                       /* :407-412 */
        }
      }
      for (int i = 0; i < n; ++i)                                                   /* revert :200-208 */
        if (crash[i]) {
          const int q = (step + 1) * MOVE_LEN[act[i]];
          for (int k = q / MAXSTEPS; k < plen[i]; ++k) path[i][k] = loc0[i];
          cur[i] = loc0[i];
        }
      if (loops >= 2 * n && count > 0) out->unresolved = 1;                         /* :400-402 */
    }
    if (apples)                                                                     /* :531-540 */
      for (int e = 0; e < n_eaters; ++e)
        for (int k = 0; k < 2; ++k)
          if (apple_on[k] && cur[e].r == apples[k].r && cur[e].c == apples[k].c) out->caught[e][k] += 1;
  }
  for (int i = 0; i < n; ++i) { out->loc[i] = cur[i]; out->crash[i] = crash[i]; out->restr[i] = restr[i]; }   /* :552 */
}

/* ------------------------------------------------------------------ FeAR */
static int manhattan(cell_t a, cell_t b) { return abs(a.r - b.r) + abs(a.c - b.c); }

/* CountValidMovesOfAffected_tuple, Responsibility.py:20-54; agents outside `in_list` Stay (:43) */
static int count_valid(const gwo_config* cfg, int n, const cell_t* loc, const int* list_act, const int* in_list,
                       int affected) {
  int count = 0;
  for (int a = 0; a < GW_N_ACTIONS; ++a) {
    int act[OR_A];
    for (int i = 0; i < n; ++i) act[i] = in_list[i] ? list_act[i] : 0;
    if (in_list[affected]) act[affected] = a;                     /* SwapActionIDs4Agents, grid_world.py:709-726 */
    update_out o;
    update_world(cfg, n, loc, act, NULL, NULL, 0, &o);
    if (!o.crash[affected] && !o.restr[affected]) ++count;        /* :46 */
  }
  return count;
}

static void fear_one_actor(const gwo_handle* h, int n, const cell_t* loc, const int* act, const int* in_list,
                           const int* mdr, int actor, double* resp, int* n_mdr, int* n_act) {
  for (int j = 0; j < OR_A; ++j) { resp[j] = 0.0; n_mdr[j] = 0; n_act[j] = 0; }
  for (int jj = 0; jj < n; ++jj) {                                /* Responsibility.py:163-198 */
    if (jj == actor) continue;
    int la[OR_A];
    for (int i = 0; i < n; ++i) la[i] = act[i];
    la[actor] = mdr[actor];
    n_mdr[jj] = count_valid(&h->cfg, n, loc, la, in_list, jj);
    la[actor] = act[actor];
    n_act[jj] = count_valid(&h->cfg, n, loc, la, in_list, jj);
    double r = ((double)n_mdr[jj] - (double)n_act[jj]) / ((double)n_mdr[jj] + 0.000001);
    resp[jj] = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
  }
}

/* np.sum over the n x n Resp matrix (a single non-zero row) in numpy's pairwise order (numpy/_core/src/umath/
 * loops_utils.h.src, pairwise_sum): fewer than 8 elements are added left to right; up to 128 go through 8 interleaved
 * accumulators r[k] += a[i+k] combined as ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)), then the tail; longer arrays are split
 * at n/2 rounded down to a multiple of 8 and the halves' sums added (n = 12..16 agents: 144..256 elements). */
static double np_pairwise(const double* a, int len) {
  if (len < 8) {
    double res = 0.0;
    for (int i = 0; i < len; ++i) res += a[i];
    return res;
  }
  if (len <= 128) {
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < len - (len % 8); i += 8)
      for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < len; ++i) res += a[i];
    return res;
  }
  int half = len / 2;
  half -= half % 8;
  return np_pairwise(a, half) + np_pairwise(a + half, len - half);
}

static double np_sum_matrix(int n, int actor, const double* row) {
  double a[OR_A * OR_A];
  const int len = n * n;
  for (int i = 0; i < len; ++i) a[i] = 0.0;
  for (int j = 0; j < n; ++j) a[actor * n + j] = row[j];
  return np_pairwise(a, len);
}

/* ------------------------------------------------------------------ RNG (device-RNG mode restated) */
static void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1,
                   n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

static void policy_thresholds(const float sw[3], const float dw_in[4], int perturbed, uint32_t thr[8]) {
  double dw[4], p[9], tot = 0, cdf = 0;
  for (int d = 0; d < 4; ++d) dw[d] = perturbed ? 1.0 : (double)dw_in[d];
  p[0] = sw[0];
  for (int d = 0; d < 4; ++d) { p[1 + d] = (double)sw[1] * dw[d]; p[5 + d] = (double)sw[2] * dw[d]; }
  for (int k = 0; k < 9; ++k) tot += p[k];
  for (int k = 0; k < 8; ++k) {
    cdf += p[k] / tot;
    double t = floor(cdf * 2147483648.0 + 0.5);
    if (t > 2147483648.0) t = 2147483648.0;
    if (t < 0) t = 0;
    thr[k] = (uint32_t)t;
  }
  int last = 8;
  while (last > 0 && p[last] == 0.0) --last;
  for (int k = last; k < 8; ++k) thr[k] = 0x80000000u;
}

static void spawn(const gwo_handle* h, int64_t e, uint32_t tick, const int8_t* spawn_in, cell_t* loc) {
  const gwo_config* c = &h->cfg;
  if (spawn_in) {
    for (int i = 0; i < c->n_agents; ++i) {
      loc[i].r = spawn_in[(e * c->n_agents + i) * 2];
      loc[i].c = spawn_in[(e * c->n_agents + i) * 2 + 1];
    }
    return;
  }
  const uint64_t gid = (uint64_t)(c->env_id_base + e);
  uint32_t w[4] = {0, 0, 0, 0};
  int chosen[OR_A] = {0};
  for (int k = 0; k < c->n_agents; ++k) {
    if ((k & 3) == 0) {                                /* draw k: Philox call 0x100 + k/4, word k%4 */
      w[0] = (uint32_t)gid; w[1] = (uint32_t)(gid >> 32); w[2] = tick; w[3] = 0x100u + (uint32_t)(k >> 2);
      philox4x32(w, (uint32_t)c->seed, (uint32_t)(c->seed >> 32));
    }
    int d = (int)(((uint64_t)w[k & 3] * (uint32_t)(h->n_active - k)) >> 32);
    int pos = 0;
    for (int t = 0; t < k; ++t)
      if (d >= chosen[t]) { ++d; pos = t + 1; }
    for (int t = k; t > pos; --t) chosen[t] = chosen[t - 1];
    chosen[pos] = d;
  }
  for (int i = 0; i < c->n_agents; ++i) loc[i] = h->active[chosen[i]];
}

/* ------------------------------------------------------------------ outputs */
static uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u); /* round to nearest even (all values here are exact anyway) */
  return (uint16_t)(u >> 16);
}

static void write_obs(const gwo_handle* h, void* base, int64_t e, const env_state* s, int fresh) {
  const gwo_config* c = &h->cfg;
  const int len = c->height * c->width;
  for (int k = 0; k < c->n_learners; ++k) {
    float obs[OR_CELLS];
    for (int r = 0; r < c->height; ++r)
      for (int col = 0; col < c->width; ++col) obs[r * c->width + col] = cfg_active(c, r, col) ? 0.0f : -1.0f;
    for (int i = 0; i < c->n_agents; ++i)                       /* WorldState: grid_world.py:230 / AddAgent :140 */
      obs[s->loc[i].r * c->width + s->loc[i].c] = fresh ? 0.5f : (float)(i + 1);
    if (c->env_kind == GW_ENV_MULTI) {
      if (s->apple_present[k]) obs[c->apple_row[k] * c->width + c->apple_col[k]] += 9.0f;   /* ma :306-312 */
      if (!fresh) {                                             /* :314-321 */
        for (int i = 0; i < len; ++i) {
          for (int id = 1; id <= 4; ++id)
            if (id != k + 1 && obs[i] == (float)id) { obs[i] = 5.0f; break; }
        }
        for (int i = 0; i < len; ++i)
          if (obs[i] == (float)(k + 1)) obs[i] = 1.0f;
      }
    } else {
      if (s->apple_present[0]) obs[c->apple_row[0] * c->width + c->apple_col[0]] += 9.0f;   /* customenv :161-163 */
    }
    if (c->obs_dtype == GW_OBS_F32) {
      memcpy((float*)base + (e * c->n_learners + k) * len, obs, sizeof(float) * len);
    } else {
      uint16_t* dst = (uint16_t*)base + (e * c->n_learners + k) * len;
      for (int i = 0; i < len; ++i) dst[i] = f32_to_bf16(obs[i]);
    }
  }
}

static void write_masks(const gwo_handle* h, int8_t* dst, int64_t e, const env_state* s) {
  const gwo_config* c = &h->cfg;
  if (!dst) return;
  for (int k = 0; k < c->n_learners; ++k) {                     /* ma_customenv.py:467-506: target cell only */
    int8_t* m = dst + (e * c->n_learners + k) * GW_N_ACTIONS;
    m[0] = 1;
    for (int a = 1; a < GW_N_ACTIONS; ++a)
      m[a] = (int8_t)cfg_active(c, s->loc[k].r + MOVE_DR[a] * MOVE_LEN[a], s->loc[k].c + MOVE_DC[a] * MOVE_LEN[a]);
  }
}

static void fresh_env(const gwo_handle* h, env_state* s) {
  const gwo_config* c = &h->cfg;
  for (int k = 0; k < GW_MAX_LEARNERS; ++k) {
    s->apple_present[k] = k < c->n_learners && c->apple_row[k] >= 0;
    s->term[k] = 0; s->pd_valid[k] = 0; s->pd[k] = 0; s->ep_ret[k] = 0;
  }
  s->trunc = 0; s->steps = 0;
  if (c->env_kind == GW_ENV_SINGLE) {                           /* customenv.py:349-352 */
    cell_t apple = {c->apple_row[0], c->apple_col[0]};
    s->pd_valid[0] = 1;
    s->pd[0] = manhattan(s->loc[0], apple);
  }
}

/* ------------------------------------------------------------------ public API */
#define EXPORT __attribute__((visibility("default")))

EXPORT int gwo_create(const gwo_config* cfg, gwo_handle** out) {
  if (!cfg || !out || cfg->struct_size != (int32_t)sizeof(gwo_config)) return GW_EINVAL;
  gwo_handle* h = (gwo_handle*)calloc(1, sizeof(gwo_handle));
  if (!h) return GW_ENOMEM;
  h->cfg = *cfg;
  h->env = (env_state*)calloc((size_t)cfg->num_envs, sizeof(env_state));
  if (!h->env) { free(h); return GW_ENOMEM; }
  for (int p = 0; p < cfg->n_policies; ++p) {
    policy_thresholds(cfg->step_weights[p], cfg->dir_weights[p], 0, h->thr[p][0]);
    policy_thresholds(cfg->step_weights[p], cfg->dir_weights[p], 1, h->thr[p][1]);
  }
  for (int r = 0; r < cfg->height; ++r)
    for (int c = 0; c < cfg->width; ++c)
      if (cfg_active(cfg, r, c)) { h->active[h->n_active].r = r; h->active[h->n_active].c = c; ++h->n_active; }
  for (int m = 0; m < 10; ++m)
    for (int a = 0; a < 10; ++a) {
      double r = ((double)m - (double)a) / ((double)m + 0.000001);
      h->resp_lut[m][a] = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
    }
  pthread_mutex_init(&h->stat_mu, NULL);
  *out = h;
  return GW_OK;
}

EXPORT int gwo_destroy(gwo_handle* h) {
  if (!h) return GW_OK;
  pthread_mutex_destroy(&h->stat_mu);
  free(h->env);
  free(h);
  return GW_OK;
}

static void reset_range(gwo_handle* h, const uint8_t* mask, const gw_io* io, int64_t lo, int64_t hi) {
  const gwo_config* c = &h->cfg;
  for (int64_t e = lo; e < hi; ++e) {
    if (mask && !mask[e]) continue;
    env_state* s = &h->env[e];
    spawn(h, e, s->tick, io->spawn, s->loc);
    fresh_env(h, s);
    s->tick += 1;
    if (io->obs) write_obs(h, io->obs, e, s, 1);
    write_masks(h, io->action_mask, e, s);
    if (io->positions)
      for (int i = 0; i < c->n_agents; ++i) {
        io->positions[(e * c->n_agents + i) * 2] = (int8_t)s->loc[i].r;
        io->positions[(e * c->n_agents + i) * 2 + 1] = (int8_t)s->loc[i].c;
      }
  }
}

EXPORT int gwo_reset(gwo_handle* h, const uint8_t* mask, const gw_io* io) {
  if (!h || !io) return GW_EINVAL;
  reset_range(h, mask, io, 0, h->cfg.num_envs);
  h->reset_done = 1;
  return GW_OK;
}

typedef struct { uint64_t episodes, len_sum, crashes, apples, unresolved, fear_nz; double return_sum, fear_sum; } stat_acc;

static void step_range(gwo_handle* h, const gw_io* io, int64_t lo, int64_t hi, stat_acc* acc) {
  const gwo_config* c = &h->cfg;
  const int n = c->n_agents, nl = c->n_learners;
  double pthr = c->perturb_prob * 4294967296.0;
  const uint32_t perturb_thr = pthr >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)pthr;
  for (int64_t e = lo; e < hi; ++e) {
    env_state* s = &h->env[e];
    int act[OR_A], mdr[OR_A];
    /* setup_step, ma_customenv.py:432-452 */
    for (int i = 0; i < n; ++i) {
      mdr[i] = c->mdr_map[CELL_IDX(c, s->loc[i].r, s->loc[i].c)];
      if (i < nl) act[i] = io->learner_actions[e * nl + i];
      else if (io->npc_actions) act[i] = io->npc_actions[e * n + i];
      else {
        /* RNG spec: NPC number m = i - n_learners uses Philox call m/2 (counter = global env id, tick,
         * call index; key = seed), words 2(m%2) [perturbation] and 2(m%2)+1 [action draw] */
        const uint64_t gid = (uint64_t)(c->env_id_base + e);
        const int m = i - nl;
        uint32_t w[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), s->tick, (uint32_t)(m >> 1)};
        philox4x32(w, (uint32_t)c->seed, (uint32_t)(c->seed >> 32));
        const int pert = w[2 * (m & 1)] < perturb_thr;
        const uint32_t* thr = h->thr[c->policy_map[CELL_IDX(c, s->loc[i].r, s->loc[i].c)]][pert];
        const uint32_t u = w[2 * (m & 1) + 1] >> 1;
        int a = 0;
        for (int k = 0; k < 8; ++k) a += (u >= thr[k]);
        act[i] = a;
      }
      if (act[i] < 0) act[i] = 0;
      if (act[i] > 8) act[i] = 8;
    }
    /* FeAR, ma_customenv.py:245-252 / customenv.py:113-120 */
    double fear[GW_MAX_LEARNERS] = {0.0, 0.0};
    if (c->fear) {
      for (int x = 0; x < nl; ++x) {
        int in_list[OR_A], n_close = 0;
        for (int k = 0; k < n; ++k) {                                  /* close_agents :456-464 */
          in_list[k] = (k == x) || manhattan(s->loc[x], s->loc[k]) <= c->fear_radius;
          n_close += in_list[k];
        }
        if (c->env_kind == GW_ENV_SINGLE && n_close <= 1) { fear[x] = 0.0; continue; }   /* customenv.py:117-118 */
        double resp[OR_A];
        int nm[OR_A], na[OR_A];
        fear_one_actor(h, n, s->loc, act, in_list, mdr, x, resp, nm, na);
        fear[x] = np_sum_matrix(n, x, resp);
      }
    }
    /* UpdateGWorld :254 */
    cell_t apples[2] = {{c->apple_row[0], c->apple_col[0]}, {c->apple_row[1], c->apple_col[1]}};
    int apple_on[2] = {s->apple_present[0], s->apple_present[1]};
    update_out u;
    update_world(c, n, s->loc, act, apples, apple_on, nl, &u);
    for (int i = 0; i < n; ++i) s->loc[i] = u.loc[i];
    double reward[GW_MAX_LEARNERS] = {0.0, 0.0};
    int term_now[GW_MAX_LEARNERS] = {0, 0}, trunc_now = 0, apples_rewarded = 0, crash_count = 0, shaped[2] = {0, 0};
    if (c->env_kind == GW_ENV_MULTI) {
      int ri[GW_MAX_LEARNERS] = {0, 0};
      /* apples_caught list order: sub-step, eater, apple; only the own apple counts (:258-275).  The list order
       * cannot change the totals, so process learners in index order. */
      for (int k = 0; k < nl; ++k)
        if (u.caught[k][k] > 0 && s->apple_present[k]) {
          s->apple_present[k] = 0;
          ri[k] += 20;
          ++apples_rewarded;
          int any = 0;
          for (int q = 0; q < nl; ++q) any |= s->apple_present[q];
          if (!any) {
            for (int q = 0; q < nl; ++q) ri[q] += 20;
            s->trunc = 1;
          }
        }
      int dist_valid[2] = {0, 0}, dist[2] = {0, 0};
      for (int k = 0; k < nl; ++k) {                                   /* :278-300 */
        if (u.crash[k]) { ri[k] -= 10; ++crash_count; s->trunc = 1; s->term[k] = 1; }
        if (s->apple_present[k]) {
          cell_t ap = {c->apple_row[k], c->apple_col[k]};
          dist_valid[k] = 1;
          dist[k] = manhattan(s->loc[k], ap);
        }
        if (s->pd_valid[k] && dist_valid[k] && s->pd[k] > dist[k]) { ri[k] += 1; shaped[k] = 1; }
      }
      for (int k = 0; k < nl; ++k) {
        s->pd_valid[k] = dist_valid[k]; s->pd[k] = dist[k];            /* :302 */
        reward[k] = (double)ri[k];
        term_now[k] = s->term[k];
      }
      trunc_now = s->trunc;
    } else {                                                           /* customenv.py:126-158 */
      double rew = 0.0;
      cell_t ap = {c->apple_row[0], c->apple_col[0]};
      const int d = manhattan(s->loc[0], ap);
      if (u.crash[0]) { rew -= 10.0; term_now[0] = 1; crash_count = 1; }
      if (s->apple_present[0] && u.caught[0][0] == 1) { s->apple_present[0] = 0; rew += 20.0; trunc_now = 1; apples_rewarded = 1; }
      if (d < s->pd[0]) { rew += 0.1; shaped[0] = 1; }
      reward[0] = rew;
      s->pd[0] = d; s->pd_valid[0] = 1;
    }
    s->steps += 1;
    if (s->steps > 0xFFF) s->steps = 0xFFF;
    const int over = c->env_kind == GW_ENV_MULTI ? trunc_now : (term_now[0] || trunc_now);
    const int ended = over || (c->max_steps > 0 && s->steps >= c->max_steps);
    for (int k = 0; k < nl; ++k) {
      const int64_t o = e * nl + k;
      if (io->reward) io->reward[o] = (float)reward[k];
      if (io->fear) io->fear[o] = fear[k];
      if (io->shaped_reward) io->shaped_reward[o] = (float)(c->fear_weight * fear[k] + reward[k]);   /* maddpg/agent.py:130 */
      if (io->terminated) io->terminated[o] = (uint8_t)term_now[k];
      if (io->truncated) io->truncated[o] = (uint8_t)(trunc_now ? 1 : 0);
    }
    if (io->positions)
      for (int i = 0; i < n; ++i) {
        io->positions[(e * n + i) * 2] = (int8_t)s->loc[i].r;
        io->positions[(e * n + i) * 2 + 1] = (int8_t)s->loc[i].c;
      }
    if (io->ended) io->ended[e] = (uint8_t)ended;
    if (io->info) {
      uint32_t bits = 0;
      for (int i = 0; i < n && i < 4; ++i) bits |= ((uint32_t)u.crash[i] << i) | ((uint32_t)u.restr[i] << (4 + i));
#ifdef GWO_WIDE
      for (int i = 0; i < n; ++i) bits |= (uint32_t)u.crash[i] << (16 + i);
#endif
      bits |= (uint32_t)crash_count << 8 | (uint32_t)apples_rewarded << 10 | (uint32_t)ended << 12 |
              (uint32_t)u.unresolved << 13 | (uint32_t)shaped[0] << 14 | (uint32_t)shaped[1] << 15;
      io->info[e] = bits;
    }
    const double unit = c->env_kind == GW_ENV_MULTI ? 1.0 : 10.0;
    s->ep_ret[0] += (int)lrint(reward[0] * unit);
    s->ep_ret[1] += (int)lrint(reward[1] * unit);
    if (ended) {
      acc->episodes += 1; acc->len_sum += (uint64_t)s->steps;
      acc->return_sum += (double)(s->ep_ret[0] + s->ep_ret[1]) / unit;
    }
    acc->crashes += (uint64_t)crash_count; acc->apples += (uint64_t)apples_rewarded; acc->unresolved += (uint64_t)u.unresolved;
    if (c->fear) {
      const int nz = (fear[0] != 0.0) + (fear[1] != 0.0);
      if (nz) { acc->fear_nz += (uint64_t)nz; acc->fear_sum += fear[0] + fear[1]; }
    }
    if (ended && c->auto_reset) {
      if (io->final_obs) write_obs(h, io->final_obs, e, s, 0);
      spawn(h, e, s->tick, io->spawn, s->loc);
      fresh_env(h, s);
      if (io->obs) write_obs(h, io->obs, e, s, 1);
    } else if (io->obs) {
      write_obs(h, io->obs, e, s, 0);
    }
    write_masks(h, io->action_mask, e, s);
    s->tick += 1;
  }
}

typedef struct { gwo_handle* h; const gw_io* io; int64_t lo, hi; stat_acc acc; } job_t;

static void* step_job(void* arg) {
  job_t* j = (job_t*)arg;
  step_range(j->h, j->io, j->lo, j->hi, &j->acc);
  return NULL;
}

static void merge(gwo_handle* h, const stat_acc* a) {
  h->episodes += a->episodes; h->len_sum += a->len_sum; h->crashes += a->crashes; h->apples += a->apples;
  h->unresolved += a->unresolved; h->fear_nz += a->fear_nz; h->return_sum += a->return_sum; h->fear_sum += a->fear_sum;
}

/* n_threads <= 1: in the calling thread */
EXPORT int gwo_step(gwo_handle* h, const gw_io* io, int n_threads) {
  if (!h || !io || !io->learner_actions) return GW_EINVAL;
  if (!h->reset_done) return GW_ESTATE;
  const int64_t E = h->cfg.num_envs;
  if (n_threads > E) n_threads = (int)E;
  if (n_threads <= 1) {
    stat_acc acc;
    memset(&acc, 0, sizeof(acc));
    step_range(h, io, 0, E, &acc);
    merge(h, &acc);
  } else {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job_t* jobs = (job_t*)calloc((size_t)n_threads, sizeof(job_t));
    for (int t = 0; t < n_threads; ++t) {
      jobs[t].h = h; jobs[t].io = io;
      jobs[t].lo = E * t / n_threads; jobs[t].hi = E * (t + 1) / n_threads;
      pthread_create(&th[t], NULL, step_job, &jobs[t]);
    }
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); merge(h, &jobs[t].acc); }
    free(th); free(jobs);
  }
  h->env_steps += (uint64_t)E;
  return GW_OK;
}

EXPORT int gwo_get_stats(gwo_handle* h, gw_stats* out) {
  if (!h || !out) return GW_EINVAL;
  memset(out, 0, sizeof(*out));
  out->env_steps = h->env_steps; out->agent_steps = h->env_steps * (uint64_t)h->cfg.n_learners;
  out->episodes = h->episodes; out->episode_len_sum = h->len_sum; out->crashes = h->crashes; out->apples = h->apples;
  out->unresolved = h->unresolved; out->fear_nonzero = h->fear_nz; out->return_sum = h->return_sum; out->fear_sum = h->fear_sum;
  return GW_OK;
}

#ifdef GWO_WIDE
/* gww_env_state (include/gridworld_b200.h), 64 bytes per env */
EXPORT int gwo_get_state(gwo_handle* h, uint32_t* dst_words) {
  if (!h || !dst_words) return GW_EINVAL;
  gww_env_state* dst = (gww_env_state*)dst_words;
  for (int64_t e = 0; e < h->cfg.num_envs; ++e) {
    const env_state* s = &h->env[e];
    gww_env_state o;
    memset(&o, 0, sizeof(o));
    for (int i = 0; i < OR_A; ++i)
      o.cell[i] = i < h->cfg.n_agents ? (uint16_t)((s->loc[i].r << 8) | s->loc[i].c) : (uint16_t)0xFFFFu;
    o.flags = (uint32_t)s->apple_present[0] | (uint32_t)s->apple_present[1] << 1;
    if (h->cfg.env_kind == GW_ENV_MULTI) o.flags |= (uint32_t)s->term[0] << 2 | (uint32_t)s->term[1] << 3 | (uint32_t)s->trunc << 4;
    o.flags |= (uint32_t)s->pd_valid[0] << 5 | (uint32_t)s->pd_valid[1] << 6;
    o.tick = s->tick;
    o.episode_return[0] = s->ep_ret[0]; o.episode_return[1] = s->ep_ret[1];
    o.prev_distance[0] = (uint16_t)s->pd[0]; o.prev_distance[1] = (uint16_t)s->pd[1];
    o.steps = (uint32_t)s->steps;
    dst[e] = o;
  }
  return GW_OK;
}
#else
/* the same 16-byte packed state as gw_get_state (csrc/gw_kernels.cu "meta word") */
EXPORT int gwo_get_state(gwo_handle* h, uint32_t* dst) {
  if (!h || !dst) return GW_EINVAL;
  for (int64_t e = 0; e < h->cfg.num_envs; ++e) {
    const env_state* s = &h->env[e];
    uint32_t cells = 0;
    for (int i = 0; i < h->cfg.n_agents; ++i) cells |= (uint32_t)((s->loc[i].r << 4) | s->loc[i].c) << (8 * i);
    uint32_t meta = (uint32_t)s->apple_present[0] | (uint32_t)s->apple_present[1] << 1;
    if (h->cfg.env_kind == GW_ENV_MULTI) meta |= (uint32_t)s->term[0] << 2 | (uint32_t)s->term[1] << 3 | (uint32_t)s->trunc << 4;
    meta |= (uint32_t)s->pd_valid[0] << 5 | (uint32_t)s->pd_valid[1] << 6 | (uint32_t)s->pd[0] << 7 | (uint32_t)s->pd[1] << 12 |
            (uint32_t)s->steps << 17;
    dst[e * 4 + 0] = cells; dst[e * 4 + 1] = meta; dst[e * 4 + 2] = s->tick;
    dst[e * 4 + 3] = ((uint32_t)s->ep_ret[0] & 0xFFFFu) | ((uint32_t)s->ep_ret[1] << 16);
  }
  return GW_OK;
}

#endif

EXPORT int gwo_update_world(gwo_handle* h, int64_t C, const int8_t* n_per, const int8_t* pos, const int8_t* act,
                            const int8_t* apples, int8_t* new_pos, uint8_t* crash, uint8_t* restr, int8_t* caught) {
  if (!h) return GW_EINVAL;
  for (int64_t c = 0; c < C; ++c) {
    const int n = n_per ? n_per[c] : h->cfg.n_agents;
    cell_t loc[OR_A], ap[2] = {{-1, -1}, {-1, -1}};
    int a[OR_A] = {0}, on[2] = {0, 0};
    for (int i = 0; i < n; ++i) { loc[i].r = pos[(c * OR_A + i) * 2]; loc[i].c = pos[(c * OR_A + i) * 2 + 1]; a[i] = act[c * OR_A + i]; }
    if (apples)
      for (int k = 0; k < 2; ++k)
        if (apples[(c * 2 + k) * 2] >= 0) { on[k] = 1; ap[k].r = apples[(c * 2 + k) * 2]; ap[k].c = apples[(c * 2 + k) * 2 + 1]; }
    update_out u;
    update_world(&h->cfg, n, loc, a, apples ? ap : NULL, on, n < 2 ? n : 2, &u);
    for (int i = 0; i < OR_A; ++i) {
      new_pos[(c * OR_A + i) * 2] = i < n ? (int8_t)u.loc[i].r : -1;
      new_pos[(c * OR_A + i) * 2 + 1] = i < n ? (int8_t)u.loc[i].c : -1;
      crash[c * OR_A + i] = i < n ? (uint8_t)u.crash[i] : 0;
      restr[c * OR_A + i] = i < n ? (uint8_t)u.restr[i] : 0;
    }
    if (caught)
      for (int e = 0; e < 2; ++e)
        for (int k = 0; k < 2; ++k) caught[c * 4 + e * 2 + k] = (int8_t)u.caught[e][k];
  }
  return GW_OK;
}

EXPORT int gwo_fear_one_actor(gwo_handle* h, int64_t C, const int8_t* n_per, const int8_t* pos, const int8_t* act,
                              const int8_t* mdr, const int8_t* actor, const uint8_t* in_list, double* resp,
                              int8_t* n_mdr, int8_t* n_act, double* fear_sum) {
  if (!h) return GW_EINVAL;
  for (int64_t c = 0; c < C; ++c) {
    const int n = n_per ? n_per[c] : h->cfg.n_agents;
    cell_t loc[OR_A];
    int a[OR_A] = {0}, m[OR_A] = {0}, il[OR_A] = {0};
    for (int i = 0; i < n; ++i) {
      loc[i].r = pos[(c * OR_A + i) * 2]; loc[i].c = pos[(c * OR_A + i) * 2 + 1];
      a[i] = act[c * OR_A + i]; m[i] = mdr[c * OR_A + i];
      il[i] = in_list ? (in_list[c * OR_A + i] != 0) : 1;
    }
    il[actor[c]] = 1;
    double r[OR_A];
    int nm[OR_A], na[OR_A];
    fear_one_actor(h, n, loc, a, il, m, actor[c], r, nm, na);
    for (int i = 0; i < OR_A; ++i) {
      resp[c * OR_A + i] = i < n ? r[i] : 0.0;
      if (n_mdr) n_mdr[c * OR_A + i] = (int8_t)(i < n ? nm[i] : 0);
      if (n_act) n_act[c * OR_A + i] = (int8_t)(i < n ? na[i] : 0);
    }
    if (fear_sum) fear_sum[c] = np_sum_matrix(n, actor[c], r);
  }
  return GW_OK;
}

/* Responsibility.FeAR (all actors), custom/Responsibility.py:57-132.  An actor outside the list is never swapped. */
EXPORT int gwo_fear_matrix(gwo_handle* h, int64_t C, const int8_t* n_per, const int8_t* pos, const int8_t* act,
                           const int8_t* mdr, const uint8_t* in_list, double* resp, int8_t* n_mdr, int8_t* n_act) {
  if (!h) return GW_EINVAL;
  for (int64_t c = 0; c < C; ++c) {
    const int n = n_per ? n_per[c] : h->cfg.n_agents;
    cell_t loc[OR_A];
    int a[OR_A] = {0}, il[OR_A] = {0};
    for (int i = 0; i < n; ++i) {
      loc[i].r = pos[(c * OR_A + i) * 2]; loc[i].c = pos[(c * OR_A + i) * 2 + 1];
      a[i] = act[c * OR_A + i];
      il[i] = in_list ? (in_list[c * OR_A + i] != 0) : 1;
    }
    for (int i = 0; i < OR_A * OR_A; ++i) { resp[c * OR_A * OR_A + i] = 0.0; n_mdr[c * OR_A * OR_A + i] = 0; n_act[c * OR_A * OR_A + i] = 0; }
    for (int ii = 0; ii < n; ++ii)
      for (int jj = 0; jj < n; ++jj) {
        if (ii == jj) continue;
        int la[OR_A];
        for (int i = 0; i < n; ++i) la[i] = a[i];
        if (il[ii]) la[ii] = mdr[c * OR_A + ii];
        const int m = count_valid(&h->cfg, n, loc, la, il, jj);
        const int v = count_valid(&h->cfg, n, loc, a, il, jj);
        double r = ((double)m - (double)v) / ((double)m + 0.000001);
        resp[(c * OR_A + ii) * OR_A + jj] = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
        n_mdr[(c * OR_A + ii) * OR_A + jj] = (int8_t)m;
        n_act[(c * OR_A + ii) * OR_A + jj] = (int8_t)v;
      }
  }
  return GW_OK;
}

/* Responsibility.FeAL, custom/Responsibility.py:213-303 */
EXPORT int gwo_feal(gwo_handle* h, int64_t C, const int8_t* n_per, const int8_t* pos, const int8_t* act,
                    const int8_t* mdr, const uint8_t* in_list, double* feal, int8_t* n_mdr, int8_t* n_act) {
  if (!h) return GW_EINVAL;
  for (int64_t c = 0; c < C; ++c) {
    const int n = n_per ? n_per[c] : h->cfg.n_agents;
    cell_t loc[OR_A];
    int a[OR_A] = {0}, il[OR_A] = {0};
    for (int i = 0; i < n; ++i) {
      loc[i].r = pos[(c * OR_A + i) * 2]; loc[i].c = pos[(c * OR_A + i) * 2 + 1];
      a[i] = act[c * OR_A + i];
      il[i] = in_list ? (in_list[c * OR_A + i] != 0) : 1;
    }
    for (int ii = 0; ii < OR_A; ++ii) {
      feal[c * OR_A + ii] = 0.0; n_mdr[c * OR_A + ii] = 0; n_act[c * OR_A + ii] = 0;
      if (ii >= n) continue;
      int la[OR_A];
      for (int i = 0; i < n; ++i) la[i] = (i == ii) ? a[i] : mdr[c * OR_A + i];
      const int m = count_valid(&h->cfg, n, loc, la, il, ii);
      const int v = count_valid(&h->cfg, n, loc, a, il, ii);
      double r = (double)v / ((double)m + 0.000001);
      feal[c * OR_A + ii] = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
      n_mdr[c * OR_A + ii] = (int8_t)m;
      n_act[c * OR_A + ii] = (int8_t)v;
    }
  }
  return GW_OK;
}
