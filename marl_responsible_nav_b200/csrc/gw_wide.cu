// gw_wide.cu -- the GENERAL state layout (gww_* in include/gridworld_b200.h): grids up to 64 x 64, up to
// 16 world agents, restricted paths per cell and direction.
//
// The packed layout (gw_kernels.cu) folds the whole collision logic into a table over the 41 cell offsets within
// Manhattan distance 4 and the 13 effective trajectories of FOUR agents on a 16-wide grid.  None of that carries over to
// an arbitrary agent count and map, so this translation unit evaluates the reference's rules on the paths themselves:
//   world_update   <- GWorld.UpdateGWorld                custom/grid_world.py:424-563
//   pair test      <- collision_checks_and_resolution    :233-405 (five rules, first match wins, fix-point <= 2N passes)
//   revert         <- revertStepsWithCollisions          :190-209
//   count_valid    <- CountValidMovesOfAffected_tuple    custom/Responsibility.py:20-54
//   FeAR / FeAL    <- Responsibility.py:57-132, :135-210, :213-303
//   env step       <- custom/ma_customenv.py:217-334 (multi), custom/customenv.py:78-183 (single)
// Mapping: one CTA = 128 threads on tiles of 128 envs (32 envs while the batch is small, so that 4 096 envs spread over 128
// SMs instead of 32).  P1 thread-per-env: state, NPC draws (Philox, the packed layout's spec), MdRs, close lists.
// P2 the counterfactual simulations of FeAR as work items (env, actor, affected, variant) dealt to all threads of the CTA
// (2 x (n-1) x 2 per env, nine world updates each).  P3 thread-per-env: responsibility sums in numpy's pairwise order, the
// real world update, rewards / flags / statistics, re-spawn.  P4 warp-per-row: observation rows written cell by cell with
// unit stride (an observation is H*W values: up to 16 KB per learner, far more than the packed layout's 640 B, so it is
// the bytes of this phase that bound the kernel once FeAR is off).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/gridworld_b200.h"
#include "gw_device.cuh"
#include "gw_internal.h"

namespace gww {

constexpr int NA = GWW_MAX_AGENTS;
constexpr int THREADS = 128;               // threads per CTA; a CTA works on tiles of 32 envs (small batches: more CTAs) or 128
constexpr int MAXSTEPS = 4;                // GWorld.MaxSteps, grid_world.py:24
constexpr int STAT_SLOTS = 64;

enum { ST_EPISODES, ST_LEN, ST_CRASH, ST_APPLES, ST_UNRESOLVED, ST_FEAR_NZ, ST_RETURN_UNITS, ST_FEAR_BITS, ST_TASKS, ST_N };

struct Tab {                               // device-global, read-only, built by gww_create
  int H, W, n, nl, kind, fear, fear_radius, max_steps, auto_reset, obs_bf16, n_active, pad_;
  int apple_r[2], apple_c[2];
  uint32_t perturb_thr, pad2_;
  double fear_weight;
  unsigned long long seed;
  long long env_id_base;
  unsigned long long map_rows[GWW_MAX_DIM];
  uint32_t thr[GW_MAX_POLICIES][2][8];
  double resp_lut[10][10];
  uint8_t policy_map[GWW_MAX_CELLS];
  uint8_t mdr_map[GWW_MAX_CELLS];
  uint8_t blocked_dir[GWW_MAX_CELLS];      // bit d: the one-cell move Up / Down / Left / Right out of this cell is a restricted path
  uint16_t active[GWW_MAX_CELLS];          // row-major list of active cells, (row << 8) | col
};

struct EnvState {                          // = gww_env_state
  uint16_t cell[NA];
  uint32_t flags, tick;
  int32_t ep_ret[2];
  uint16_t pd[2];
  uint32_t steps;
  uint32_t reserved[2];
};
static_assert(sizeof(EnvState) == 64 && sizeof(gww_env_state) == 64, "state layout");

struct Params {
  const Tab* T;
  EnvState* state;
  unsigned long long* stats;
  long long E;
  gw_io io;
  const uint8_t* reset_mask;
  const uint8_t* tmpl;                     // the constant part of one env's observation block [L, H*W] in the obs dtype
  int stage_bytes;                         // bytes of one env's observation block [L, H*W] when it can leave as ONE bulk copy (a multiple of 16), else 0
  int literal_fear;                        // GWW_FEAR_LITERAL=1: nine world updates per count instead of count_valid_fast (tests)
};

__device__ __constant__ int8_t MOVE_DR[9] = {0, -1, 1, 0, 0, -1, 1, 0, 0};     // custom_agent.py:140-150
__device__ __constant__ int8_t MOVE_DC[9] = {0, 0, 0, -1, 1, 0, 0, -1, 1};
__device__ __constant__ int8_t MOVE_LEN[9] = {1, 1, 1, 1, 1, 2, 2, 2, 2};

__device__ __forceinline__ bool is_active(const unsigned long long* rows, int H, int W, int r, int c) {
  return r >= 0 && r < H && c >= 0 && c < W && ((rows[r] >> c) & 1ull);
}
__device__ __forceinline__ int manhattan(uint16_t a, uint16_t b) {
  return abs((int)(a >> 8) - (int)(b >> 8)) + abs((int)(a & 255) - (int)(b & 255));
}

struct UpdateOut {
  uint16_t loc[NA];
  uint32_t crash, restr;                   // bit i = agent i
  uint8_t caught[2][2];
  bool unresolved;
};

// GWorld.UpdateGWorld with explicit actions (grid_world.py:424-563).  `rows` = the map in shared memory.
__device__ void world_update(const Tab* __restrict__ T, const unsigned long long* rows, int n, const uint16_t* loc0,
                             const uint8_t* act, const uint16_t* apples, int apple_on, int n_eaters, UpdateOut& out) {
  const int H = T->H, W = T->W;
  uint16_t path[NA][MAXSTEPS + 1];
  uint16_t cur[NA];
  uint32_t crash = 0, restr = 0;
  out.caught[0][0] = out.caught[0][1] = out.caught[1][0] = out.caught[1][1] = 0;
  out.unresolved = false;
  for (int i = 0; i < n; ++i) path[i][0] = loc0[i];                                    // :437-439
  // pairs further apart than Manhattan distance 4 cannot share a cell within one step: no rule of the pair test can match
  uint16_t near[NA];
  for (int i = 0; i < n; ++i) {
    uint32_t m = 0;
    for (int k = i + 1; k < n; ++k) m |= (manhattan(loc0[i], loc0[k]) <= 4 ? 1u : 0u) << k;
    near[i] = (uint16_t)m;
  }
  for (int step = 0; step < MAXSTEPS; ++step) {                                       // :458
    for (int i = 0; i < n; ++i) cur[i] = loc0[i];                                     // :460
    for (int i = 0; i < n; ++i) {                                                     // :462-518
      const int a = act[i];
      const bool moving = step < MOVE_LEN[a] && !((crash >> i) & 1u) && a != 0;
      const uint16_t old = path[i][step];
      uint16_t nw = old;
      if (moving) {
        const int r = (old >> 8) + MOVE_DR[a], c = (old & 255) + MOVE_DC[a];
        if (r < 0 || r >= H || c < 0 || c >= W) restr |= 1u << i;                     // np.clip :486-487: back onto the old cell
        else if (!((rows[r] >> c) & 1ull)) restr |= 1u << i;                          // inactive target :512-515
        else if ((__ldg(&T->blocked_dir[(old >> 8) * W + (old & 255)]) >> ((a - 1) & 3)) & 1) restr |= 1u << i;   // :498, :504-507
        else nw = (uint16_t)((r << 8) | c);
      }
      path[i][step + 1] = nw;
    }
    int count = n, loops = 0;                                                         // :247-248
    while (count > 0 && loops < 2 * n) {                                              // :250
      ++loops;
      count = 0;
      uint32_t hit_mask = 0;
      for (int ii = 0; ii < n - 1; ++ii) {                                            // :255
        if (!near[ii]) continue;
        const int qi = (step + 1) * MOVE_LEN[act[ii]];
        const int fi = qi >> 2, ci = (qi + 3) >> 2;
        const uint16_t ai = path[ii][fi], bi = path[ii][ci], pi = loc0[ii];
        for (uint32_t rest = near[ii]; rest; rest &= rest - 1u) {
          const int jj = __ffs(rest) - 1;
          const int qj = (step + 1) * MOVE_LEN[act[jj]];
          const int fj = qj >> 2, cj = (qj + 3) >> 2;
          const uint16_t aj = path[jj][fj], bj = path[jj][cj], pj = loc0[jj];
          bool hit = false;
          if (ai == aj || bi == bj) hit = true;                                       // :276-278
          else if (ai == bj && bi == aj) hit = true;                                  // :291-294
          else if (ai == bj) {                                                        // :307-326
            const bool overhang = ((4 * ci - qi) + (qj - 4 * fj)) <= 4;
            hit = !(overhang && (int)bi - (int)ai == (int)bj - (int)aj);              // same direction: both coordinates differ alike
          } else if (bi == aj) {                                                      // :339-357
            const bool overhang = ((4 * cj - qj) + (qi - 4 * fi)) <= 4;
            hit = !(overhang && (int)bi - (int)ai == (int)bj - (int)aj);
          } else if ((ai == pj && pi == aj) || (bi == pj && pi == bj) || (ai == pj && pi == bj) || (bi == pj && pi == aj))
            hit = true;                                                               // :371-378
          if (hit) { ++count; hit_mask |= (1u << ii) | (1u << jj); }                  // :407-412
        }
      }
      crash |= hit_mask;
      if (n >= 2)
        for (int i = 0; i < n; ++i) {
          const int q = (step + 1) * MOVE_LEN[act[i]];
          if ((crash >> i) & 1u) {                                                    // revert :200-208
            for (int k = q >> 2; k <= step + 1; ++k) path[i][k] = loc0[i];
            cur[i] = loc0[i];
          } else cur[i] = path[i][q >> 2];                                            // :259, :264
        }
      if (loops >= 2 * n && count > 0) out.unresolved = true;                         // :400-402
    }
    if (apples)                                                                       // :531-540
      for (int e = 0; e < n_eaters; ++e)
        for (int k = 0; k < 2; ++k)
          if (((apple_on >> k) & 1) && cur[e] == apples[k]) out.caught[e][k] += 1;
  }
  for (int i = 0; i < n; ++i) out.loc[i] = cur[i];                                    // :552
  out.crash = crash; out.restr = restr;
}
// (bi - ai == bj - aj on the packed 16-bit cells: the moves are single cells, so the difference of the packed values is
//  +-256 or +-1 or 0 and equal differences mean equal (drow, dcol); a row borrow cannot occur for unit moves inside the grid.)

// CountValidMovesOfAffected_tuple, Responsibility.py:20-54; agents outside the list Stay (:43), the affected agent's action is
// swapped only if it is in the list (SwapActionIDs4Agents, grid_world.py:709-726).  The literal form: nine world updates.
__device__ int count_valid(const Tab* T, const unsigned long long* rows, int n, const uint16_t* loc, const uint8_t* list_act,
                           uint32_t in_list, int affected) {
  int count = 0;
  uint8_t act[NA];
  for (int i = 0; i < n; ++i) act[i] = ((in_list >> i) & 1u) ? list_act[i] : (uint8_t)0;
  const bool swap = (in_list >> affected) & 1u;
  for (int a = 0; a < GW_N_ACTIONS; ++a) {
    if (swap) act[affected] = (uint8_t)a;
    UpdateOut o;
    world_update(T, rows, n, loc, act, nullptr, 0, 0, o);
    if (!((o.crash >> affected) & 1u) && !((o.restr >> affected) & 1u)) ++count;      // :46
  }
  return count;
}

// One move of grid_world.py:481-518 from `old` with action a: the new cell, or `old` with `restricted` set.
__device__ __forceinline__ uint16_t one_move(const Tab* T, const unsigned long long* rows, uint16_t old, int a, bool& restricted) {
  const int r = (old >> 8) + MOVE_DR[a], c = (old & 255) + MOVE_DC[a];
  if (r < 0 || r >= T->H || c < 0 || c >= T->W || !((rows[r] >> c) & 1ull) ||
      ((__ldg(&T->blocked_dir[(old >> 8) * T->W + (old & 255)]) >> ((a - 1) & 3)) & 1)) { restricted = true; return old; }
  return (uint16_t)((r << 8) | c);
}

// UpdateGWorld in trajectory form (what the env step runs; the literal form above stays for the operator-level calls and as
// the cross-check).  A crash reverts the path from floor(q/4) on and stops the agent (grid_world.py:200-208, :470), and
// floor(q/4) only grows, so at any sub-step an agent is either on its nominal trajectory -- start cell, cell after the
// first move, cell after the second -- or, once crashed, on its start cell with BOTH interpolation points there.  So three
// cells and a crashed bit per agent replace the path lists; a second move is attempted (and can count as restricted) only
// by an agent that did not crash in sub-step 0.  The fix-point cap of 2N passes never binds: two crashed agents stand on
// distinct start cells and cannot hit each other, so every pass with a hit crashes somebody new.
__device__ void world_update_fast(const Tab* __restrict__ T, const unsigned long long* rows, int n, const uint16_t* loc0,
                                  const uint8_t* act, const uint16_t* apples, int apple_on, int n_eaters, UpdateOut& out) {
  uint16_t p1[NA], p2[NA], near[NA];
  uint32_t r2 = 0, crash = 0, restr = 0;
  out.caught[0][0] = out.caught[0][1] = out.caught[1][0] = out.caught[1][1] = 0;
  out.unresolved = false;
  for (int i = 0; i < n; ++i) {
    const int a = act[i];
    bool ra = false, rb = false;
    p1[i] = a != 0 ? one_move(T, rows, loc0[i], a, ra) : loc0[i];
    p2[i] = MOVE_LEN[a] == 2 ? one_move(T, rows, p1[i], a, rb) : p1[i];
    restr |= (ra ? 1u : 0u) << i;
    r2 |= (rb ? 1u : 0u) << i;
    uint32_t m = 0;
    for (int k = i + 1; k < n; ++k) m |= (manhattan(loc0[i], loc0[k]) <= 4 ? 1u : 0u) << k;
    near[i] = (uint16_t)m;
  }
  for (int step = 0; step < MAXSTEPS; ++step) {
    if (step == 1) restr |= r2 & ~crash;
    int count = n;
    while (count > 0) {
      count = 0;
      uint32_t hit_mask = 0;
      for (int ii = 0; ii < n - 1; ++ii) {
        if (!near[ii]) continue;
        const int qi = (step + 1) * MOVE_LEN[act[ii]], fi = qi >> 2, ci = (qi + 3) >> 2;
        const bool xi = (crash >> ii) & 1u;
        const int pi = loc0[ii];
        const int ai = (xi || fi == 0) ? pi : (fi == 1 ? p1[ii] : p2[ii]), bi = xi ? pi : (ci == 1 ? p1[ii] : p2[ii]);
        for (uint32_t rest = near[ii]; rest; rest &= rest - 1u) {
          const int jj = __ffs(rest) - 1;
          const int qj = (step + 1) * MOVE_LEN[act[jj]], fj = qj >> 2, cj = (qj + 3) >> 2;
          const bool xj = (crash >> jj) & 1u;
          const int pj = loc0[jj];
          const int aj = (xj || fj == 0) ? pj : (fj == 1 ? p1[jj] : p2[jj]), bj = xj ? pj : (cj == 1 ? p1[jj] : p2[jj]);
          if (gw::pair_hit(ai, bi, pi, qi, fi, ci, aj, bj, pj, qj, fj, cj)) { ++count; hit_mask |= (1u << ii) | (1u << jj); }
        }
      }
      crash |= hit_mask;
    }
    if (apples)                                                                       // :531-540 (a lone agent's `cur` stays its start cell)
      for (int e = 0; e < n_eaters; ++e) {
        const int q = (step + 1) * MOVE_LEN[act[e]], f = q >> 2;
        const uint16_t cur = (n < 2 || ((crash >> e) & 1u) || f == 0) ? loc0[e] : (f == 1 ? p1[e] : p2[e]);
        for (int k = 0; k < 2; ++k)
          if (((apple_on >> k) & 1) && cur == apples[k]) out.caught[e][k] += 1;
      }
  }
  for (int i = 0; i < n; ++i)                                                         // :552: floor(4 len / 4) = len
    out.loc[i] = (n < 2 || ((crash >> i) & 1u)) ? loc0[i] : (MOVE_LEN[act[i]] == 1 ? p1[i] : p2[i]);
  out.crash = crash; out.restr = restr;
}

// Agents further apart than Manhattan distance 4 cannot meet within one step (every cell of an agent's path is within 2 of
// its start cell, and the pair rules compare cells only), so what happens to agent j depends only on the agents linked to
// it by a chain of such near pairs: its component.
__device__ __forceinline__ uint32_t near_component(int n, const uint16_t* loc, int j) {
  uint32_t comp = 1u << j, frontier = comp;
  while (frontier) {
    const int a = __ffs(frontier) - 1;
    frontier &= frontier - 1u;
    for (int k = 0; k < n; ++k)
      if (!((comp >> k) & 1u) && manhattan(loc[a], loc[k]) <= 4) { comp |= 1u << k; frontier |= 1u << k; }
  }
  return comp;
}

// The same count from ONE world update.  The nine re-simulations differ only in the affected agent j's action, and j
// influences the others only by colliding with them -- which makes the action invalid whatever follows.  So until j is hit
// the others evolve exactly as in the world WITHOUT j: one update of that world gives their crashed sets before (Cb) and
// after (Ca) the fix-point of every sub-step, and j's action a is valid iff none of its moves is restricted and at no
// sub-step s the pair test hits j against some k on k's nominal path (k not crashed before s: what pass 1 sees) or against
// k standing on its start cell (k crashed by the end of s: what the pass after k's crash sees).  Same pair rules, same
// order of the two roles (lower index first); checked against the literal form and against the C oracle in tests/.
// `comp`: j's component (near_component); the agents outside it are left out altogether.
__device__ int count_valid_fast(const Tab* T, const unsigned long long* rows, int n, const uint16_t* loc, const uint8_t* list_act,
                                uint32_t in_list, int j, uint32_t comp) {
  uint8_t act[NA];
  uint16_t nom[NA][MAXSTEPS + 1];
  for (int i = 0; i < n; ++i) act[i] = ((in_list >> i) & 1u) ? list_act[i] : (uint8_t)0;
  // the world without j: nominal paths (no collisions) and the crashed sets per sub-step
  uint32_t Cb[MAXSTEPS], Ca[MAXSTEPS];
  const uint32_t others = comp & ~(1u << j);
  for (uint32_t rest = others; rest; rest &= rest - 1u) {
    const int k = __ffs(rest) - 1;
    nom[k][0] = loc[k];
    for (int s = 0; s < MAXSTEPS; ++s) {
      bool r = false;
      nom[k][s + 1] = (act[k] != 0 && s < MOVE_LEN[act[k]]) ? one_move(T, rows, nom[k][s], act[k], r) : nom[k][s];
    }
  }
  uint32_t crash = 0;
  const int n_others = __popc(others);
  for (int s = 0; s < MAXSTEPS; ++s) {
    Cb[s] = crash;
    int count = n_others, loops = 0;
    while (count > 0 && loops < 2 * n) {              // the fix-point among the others (a crashed agent stands on its start cell;
      ++loops;                                        //  every pass with a hit crashes somebody new, so the cap never binds)
      count = 0;
      uint32_t hit_mask = 0;
      for (uint32_t ri = others; ri; ri &= ri - 1u) {
        const int ii = __ffs(ri) - 1;
        const int qi = (s + 1) * MOVE_LEN[act[ii]], fi = qi >> 2, ci = (qi + 3) >> 2;
        const bool xi = (crash >> ii) & 1u;
        const int ai = xi ? loc[ii] : nom[ii][fi], bi = xi ? loc[ii] : nom[ii][ci];
        for (uint32_t rj = ri & (ri - 1u); rj; rj &= rj - 1u) {
          const int jj = __ffs(rj) - 1;
          const int qj = (s + 1) * MOVE_LEN[act[jj]], fj = qj >> 2, cj = (qj + 3) >> 2;
          const bool xj = (crash >> jj) & 1u;
          const int aj = xj ? loc[jj] : nom[jj][fj], bj = xj ? loc[jj] : nom[jj][cj];
          if (gw::pair_hit(ai, bi, loc[ii], qi, fi, ci, aj, bj, loc[jj], qj, fj, cj)) { ++count; hit_mask |= (1u << ii) | (1u << jj); }
        }
      }
      crash |= hit_mask;
    }
    Ca[s] = crash;
  }
  // (a crashed agent's path: entries from floor(q/4) on are its start cell, so both interpolation points A and B are --
  //  a crash at sub-step s' reverts from f(s') and f is monotone in s; before its crash the path is the nominal one)
  const bool swap = (in_list >> j) & 1u;
  int count = 0, last = 0;
  for (int a = 0; a < GW_N_ACTIONS; ++a) {
    if (!swap && a > 0) { count += last; continue; }          // j not in the list: nine identical simulations (its list action: Stay)
    const int aj_act = swap ? a : (int)act[j];
    const int len = MOVE_LEN[aj_act];
    uint16_t pj[MAXSTEPS + 1];
    pj[0] = loc[j];
    bool bad = false;
    for (int s = 0; s < MAXSTEPS && !bad; ++s) {
      pj[s + 1] = (aj_act != 0 && s < len) ? one_move(T, rows, pj[s], aj_act, bad) : pj[s];
      if (bad) break;                                         // restricted: invalid whatever the others do (:46)
      const int qj = (s + 1) * len, fj = qj >> 2, cj = (qj + 3) >> 2;
      const int A = pj[fj], B = pj[cj], P = loc[j];
      for (uint32_t rk = others; rk && !bad; rk &= rk - 1u) {
        const int k = __ffs(rk) - 1;
        const int qk = (s + 1) * MOVE_LEN[act[k]], fk = qk >> 2, ck = (qk + 3) >> 2;
        if (!((Cb[s] >> k) & 1u)) {
          const int ak = nom[k][fk], bk = nom[k][ck];
          bad = j < k ? gw::pair_hit(A, B, P, qj, fj, cj, ak, bk, loc[k], qk, fk, ck) : gw::pair_hit(ak, bk, loc[k], qk, fk, ck, A, B, P, qj, fj, cj);
        }
        if (!bad && ((Ca[s] >> k) & 1u)) {
          const int pk = loc[k];
          bad = j < k ? gw::pair_hit(A, B, P, qj, fj, cj, pk, pk, pk, qk, fk, ck) : gw::pair_hit(pk, pk, pk, qk, fk, ck, A, B, P, qj, fj, cj);
        }
      }
    }
    last = bad ? 0 : 1;
    count += last;
  }
  return count;
}

// np.sum over the n x n matrix whose only non-zero row is `row` (row index x): numpy's pairwise summation
// (loops_utils.h.src pairwise_sum) over the n*n flat elements: fewer than 8 left to right; up to 128 through 8 interleaved
// accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail; longer ones split at n/2 rounded down to 8.
__device__ __forceinline__ double mat_elem(int i, int base, int n, const double* row) {
  const int k = i - base;
  return (k >= 0 && k < n) ? row[k] : 0.0;
}
__device__ double pairwise_block(int lo, int len, int base, int n, const double* row) {      // len <= 128
  if (len < 8) {
    double res = 0.0;
    for (int i = 0; i < len; ++i) res = __dadd_rn(res, mat_elem(lo + i, base, n, row));
    return res;
  }
  double r[8];
  for (int k = 0; k < 8; ++k) r[k] = mat_elem(lo + k, base, n, row);
  int i = 8;
  for (; i < len - (len % 8); i += 8)
    for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], mat_elem(lo + i + k, base, n, row));
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < len; ++i) res = __dadd_rn(res, mat_elem(lo + i, base, n, row));
  return res;
}
__device__ double np_sum_matrix(int n, int x, const double* row) {
  const int len = n * n, base = x * n;
  if (len <= 128) return pairwise_block(0, len, base, n, row);
  int half = len / 2;
  half -= half % 8;                                                                   // n <= 16: both halves <= 128
  return __dadd_rn(pairwise_block(0, half, base, n, row), pairwise_block(half, len - half, base, n, row));
}

__device__ __forceinline__ double resp_of(const Tab* T, int m, int a) { return T->resp_lut[m][a]; }

// sorted n-subset of the active cells: draw k uses Philox call 0x100 + k/4, word k%4 (DESIGN.md "RNG")
__device__ void spawn_cells(const Tab* T, long long e, uint32_t tick, const int8_t* spawn_in, uint16_t* loc) {
  const int n = T->n;
  if (spawn_in) {
    for (int i = 0; i < n; ++i)
      loc[i] = (uint16_t)(((int)spawn_in[(e * n + i) * 2] << 8) | (uint8_t)spawn_in[(e * n + i) * 2 + 1]);
    return;
  }
  const unsigned long long gid = (unsigned long long)(T->env_id_base + e);
  uint32_t w[4] = {0, 0, 0, 0};
  int chosen[NA];
  for (int k = 0; k < n; ++k) {
    if ((k & 3) == 0) {
      w[0] = (uint32_t)gid; w[1] = (uint32_t)(gid >> 32); w[2] = tick; w[3] = 0x100u + (uint32_t)(k >> 2);
      gw::philox4x32(w, (uint32_t)T->seed, (uint32_t)(T->seed >> 32));
    }
    int d = (int)(((unsigned long long)w[k & 3] * (uint32_t)(T->n_active - k)) >> 32);
    int pos = 0;
    for (int t = 0; t < k; ++t)
      if (d >= chosen[t]) { ++d; pos = t + 1; }
    for (int t = k; t > pos; --t) chosen[t] = chosen[t - 1];
    chosen[pos] = d;
  }
  for (int i = 0; i < n; ++i) loc[i] = __ldg(&T->active[chosen[i]]);
}

__device__ __forceinline__ void fresh_env(const Tab* T, EnvState& s) {
  s.flags = 0;
  for (int k = 0; k < 2; ++k) {
    if (k < T->nl && T->apple_r[k] >= 0) s.flags |= 1u << k;
    s.ep_ret[k] = 0; s.pd[k] = 0;
  }
  s.steps = 0;
  if (T->kind == GW_ENV_SINGLE) {                                                     // customenv.py:349-352
    s.flags |= 1u << 5;
    s.pd[0] = (uint16_t)manhattan(s.cell[0], (uint16_t)((T->apple_r[0] << 8) | T->apple_c[0]));
  }
}

__device__ __forceinline__ void write_masks(const Tab* T, const unsigned long long* rows, int8_t* dst, long long e, const uint16_t* cell) {
  if (!dst) return;
  for (int k = 0; k < T->nl; ++k) {                                                   // ma_customenv.py:467-506: target cell only
    int8_t* m = dst + (e * T->nl + k) * GW_N_ACTIONS;
    m[0] = 1;
    for (int a = 1; a < GW_N_ACTIONS; ++a)
      m[a] = (int8_t)is_active(rows, T->H, T->W, (cell[k] >> 8) + MOVE_DR[a] * MOVE_LEN[a], (cell[k] & 255) + MOVE_DC[a] * MOVE_LEN[a]);
  }
}

__device__ __forceinline__ void write_positions(int8_t* dst, long long e, int n, const uint16_t* cell) {
  if (!dst) return;
  for (int i = 0; i < n; ++i) {
    dst[(e * n + i) * 2] = (int8_t)(cell[i] >> 8);
    dst[(e * n + i) * 2 + 1] = (int8_t)(cell[i] & 255);
  }
}

// One observation row set (all learners) of env e, written by one warp with unit stride.
// WorldState (grid_world.py:230 / AddAgent :140), apple +9, id remap (ma_customenv.py:303-322; customenv.py:161-163).
__device__ void render_rows(const Tab* T, const unsigned long long* rows, void* base, long long e, const uint16_t* cell,
                            uint32_t apple_bits, bool fresh, int lane) {
  const int H = T->H, W = T->W, n = T->n, nl = T->nl, len = H * W;
  for (int k = 0; k < nl; ++k) {
    const int apple_k = T->kind == GW_ENV_MULTI ? k : 0;
    const int apple_cell = ((apple_bits >> apple_k) & 1u) ? (T->apple_r[apple_k] << 8 | T->apple_c[apple_k]) : -1;
    for (int idx = lane; idx < len; idx += 32) {
      const int r = idx / W, c = idx - r * W;
      const int packed = (r << 8) | c;
      float v = ((rows[r] >> c) & 1ull) ? 0.0f : -1.0f;
      for (int i = 0; i < n; ++i)
        if (cell[i] == packed) v = fresh ? 0.5f : (float)(i + 1);
      if (packed == apple_cell) v += 9.0f;
      if (T->kind == GW_ENV_MULTI && !fresh) {
        if (v == (float)(k + 1)) v = 1.0f;
        else if (v == 1.0f || v == 2.0f || v == 3.0f || v == 4.0f) v = 5.0f;
      }
      const long long o = (e * nl + k) * (long long)len + idx;
      if (T->obs_bf16) reinterpret_cast<__nv_bfloat16*>(base)[o] = __float2bfloat16(v);
      else reinterpret_cast<float*>(base)[o] = v;
    }
  }
}

// ---- observation blocks through shared memory: a warp keeps a copy of the constant part of one env's block (every learner's
// row: -1 on inactive cells, 0 on active ones) in a staging area, patches the <= n + 2 special cells per row, hands the whole
// block to the TMA engine (cp.async.bulk.global.shared::cta: whole lines, one instruction) and restores the patched cells
// once the engine has read it.  Agents and apples stand on active cells, so restoring means writing 0.
__device__ __forceinline__ void stage_put(uint8_t* stage, bool bf16, int idx, float v) {
  if (bf16) reinterpret_cast<__nv_bfloat16*>(stage)[idx] = __float2bfloat16(v);
  else reinterpret_cast<float*>(stage)[idx] = v;
}
// the constant part of one env's block (every learner's row: -1 on inactive cells, 0 on active ones), built by gww_create
__device__ __forceinline__ void stage_init(const uint8_t* __restrict__ tmpl, int bytes, uint8_t* stage, int lane) {
  for (int i = lane; i < bytes / 16; i += 32) reinterpret_cast<uint4*>(stage)[i] = __ldg(reinterpret_cast<const uint4*>(tmpl) + i);
}
__device__ __forceinline__ void stage_store(void* gdst, const uint8_t* stage, int bytes, int lane) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(stage)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
}
// What a warp needs of the tables for the staged rows, loaded once per kernel.
struct StageCtx {
  int W, n, nl, len, bytes;
  bool bf16, multi;
  int apple[2];                            // packed apple cell shown in learner k's row (the single env shows apple 0), -1 = none configured
};
__device__ __forceinline__ StageCtx stage_ctx(const Tab* T, int bytes) {
  StageCtx c;
  c.W = T->W; c.n = T->n; c.nl = T->nl; c.len = T->H * T->W; c.bytes = bytes;
  c.bf16 = T->obs_bf16 != 0; c.multi = T->kind == GW_ENV_MULTI;
  for (int k = 0; k < 2; ++k) {
    const int ak = c.multi ? k : 0;
    c.apple[k] = T->apple_r[ak] >= 0 ? (T->apple_r[ak] << 8 | T->apple_c[ak]) : -1;
  }
  return c;
}
// One env's block -> obs (or final_obs) through the warp's staging area.  Lane 16 k + i owns agent i in learner k's row
// (values of ma_customenv.py:303-322 / customenv.py:161-163, see render_rows); lane 16 k also owns k's apple when nobody stands
// on it.  The same lanes restore the constant part (0: agents and apples stand on active cells) once the engine has read the block.
__device__ __forceinline__ void render_staged(const StageCtx& c, uint8_t* stage, void* base, long long e, const uint16_t* cell,
                                              uint32_t apple_bits, bool fresh, int lane) {
  const int k = lane >> 4, i = lane & 15;
  const bool mine = k < c.nl && i < c.n;
  const int pc = mine ? (int)cell[i] : -2;
  const int apple_cell = (k < c.nl && ((apple_bits >> (c.multi ? k : 0)) & 1u)) ? c.apple[k] : -1;
  const bool on_apple = mine && pc == apple_cell;
  const uint32_t covered = __ballot_sync(0xFFFFFFFFu, on_apple);
  float v = fresh ? 0.5f : (float)(i + 1);
  if (on_apple) v += 9.0f;
  else if (c.multi && !fresh) v = (i == k) ? 1.0f : (i < 4 ? 5.0f : v);
  const int idx = mine ? k * c.len + (pc >> 8) * c.W + (pc & 255) : -1;
  const int aidx = (i == 0 && apple_cell >= 0 && !((covered >> (16 * k)) & 0xFFFFu)) ? k * c.len + (apple_cell >> 8) * c.W + (apple_cell & 255) : -1;
  if (idx >= 0) stage_put(stage, c.bf16, idx, v);
  if (aidx >= 0) stage_put(stage, c.bf16, aidx, 9.0f);
  stage_store(reinterpret_cast<uint8_t*>(base) + e * (long long)c.bytes, stage, c.bytes, lane);
  if (idx >= 0) stage_put(stage, c.bf16, idx, 0.0f);
  if (aidx >= 0) stage_put(stage, c.bf16, aidx, 0.0f);
  __syncwarp();
}

template <int TILE, bool FEAR = true>
struct Smem {
  static constexpr int FT = FEAR ? TILE : 1;   // the FeAR arrays shrink to nothing without FeAR (more CTAs per SM)
  unsigned long long rows[GWW_MAX_DIM];
  uint16_t cell[TILE][NA];                 // P1-P2: pre-step cells; P3 on: what `obs` shows (post-step, or the fresh spawn)
  uint16_t cell_final[TILE][NA];           // P4: the terminal observation's cells (envs that re-spawned)
  uint8_t act[TILE][NA];
  uint8_t mdr[FT][NA];
  uint16_t close[FT][2];                   // close list of learner x (bit i = agent i in the list)
  uint8_t cnt[FT][2][NA][2];               // valid-move counts [actor][affected][0 = MdR variant, 1 = action variant]
  uint16_t task_comp[FT * 2 * (NA - 1)];   // the queued pair's near component (actor and affected agent share it)
  uint16_t task[FT * 2 * (NA - 1)];        // FeAR work queue: (env in tile << 5) | (actor << 4) | affected -- only the pairs that need simulating
  int n_tasks;
  uint8_t render[TILE];                    // bit 0 obs fresh, bit 1 write final_obs, bits 2-3 apples shown in obs, 4-5 in final_obs, 7 live
  unsigned long long stat[ST_N];
};

template <int TILE, bool FEAR = true>
__host__ __device__ constexpr size_t smem_fixed() { return (sizeof(Smem<TILE, FEAR>) + 127) / 128 * 128; }

template <int TILE, bool FEAR>
__device__ __forceinline__ void load_rows(Smem<TILE, FEAR>& s, const Tab* T) {
  for (int i = threadIdx.x; i < GWW_MAX_DIM; i += blockDim.x) s.rows[i] = T->map_rows[i];
  if (threadIdx.x < ST_N) s.stat[threadIdx.x] = 0;
  if (threadIdx.x == 0) s.n_tasks = 0;
}

template <int TILE>
__global__ void __launch_bounds__(THREADS) gww_reset_kernel(Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<TILE, false>& s = *reinterpret_cast<Smem<TILE, false>*>(smem_raw);
  const Tab* T = p.T;
  load_rows(s, T);
  __syncthreads();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  uint8_t* const stage = p.stage_bytes ? smem_raw + smem_fixed<TILE, false>() + (size_t)warp * p.stage_bytes : nullptr;
  if (stage) stage_init(p.tmpl, p.stage_bytes, stage, lane);
  const StageCtx sc = stage_ctx(T, p.stage_bytes);
  __syncwarp();
  for (long long base = (long long)blockIdx.x * TILE; base < p.E; base += (long long)gridDim.x * TILE) {
    const long long e = base + t;
    if (t < TILE) s.render[t] = 0;
    if (t < TILE && e < p.E && (!p.reset_mask || p.reset_mask[e])) {
      EnvState st = p.state[e];
      spawn_cells(T, e, st.tick, p.io.spawn, st.cell);
      for (int i = T->n; i < NA; ++i) st.cell[i] = 0xFFFFu;
      fresh_env(T, st);
      st.tick += 1;
      p.state[e] = st;
      for (int i = 0; i < NA; ++i) s.cell[t][i] = st.cell[i];
      s.render[t] = (uint8_t)(0x80u | 1u | ((st.flags & 3u) << 2));
      write_masks(T, s.rows, p.io.action_mask, e, st.cell);
      write_positions(p.io.positions, e, T->n, st.cell);
    }
    __syncthreads();
    if (p.io.obs)
      for (int q = warp; q < TILE; q += THREADS / 32) {
        if (!(s.render[q] & 0x80u)) continue;
        if (stage) render_staged(sc, stage, p.io.obs, base + q, s.cell[q], (s.render[q] >> 2) & 3u, true, lane);
        else render_rows(T, s.rows, p.io.obs, base + q, s.cell[q], (s.render[q] >> 2) & 3u, true, lane);
      }
    __syncthreads();
  }
}

template <bool FEAR, int TILE>
__global__ void __launch_bounds__(THREADS) gww_step_kernel(Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<TILE, FEAR>& s = *reinterpret_cast<Smem<TILE, FEAR>*>(smem_raw);
  const Tab* T = p.T;
  load_rows(s, T);
  __syncthreads();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n = T->n, nl = T->nl;
  uint8_t* const stage = p.stage_bytes ? smem_raw + smem_fixed<TILE, FEAR>() + (size_t)warp * p.stage_bytes : nullptr;
  if (stage) stage_init(p.tmpl, p.stage_bytes, stage, lane);
  const StageCtx sc = stage_ctx(T, p.stage_bytes);
  __syncwarp();
  for (long long base = (long long)blockIdx.x * TILE; base < p.E; base += (long long)gridDim.x * TILE) {
    const long long e = base + t;
    const bool live = t < TILE && e < p.E;
    EnvState st;
    // ---- P1: actions of all agents (setup_step, ma_customenv.py:432-452), MdRs, close lists (:456-464)
    if (live) {
      st = p.state[e];
      const unsigned long long gid = (unsigned long long)(T->env_id_base + e);
      uint32_t w[4] = {0, 0, 0, 0};
      for (int i = 0; i < n; ++i) {
        const int ci = (st.cell[i] >> 8) * T->W + (st.cell[i] & 255);
        s.cell[t][i] = st.cell[i];
        if (FEAR) s.mdr[t][i] = __ldg(&T->mdr_map[ci]);
        int a;
        if (i < nl) a = p.io.learner_actions[e * nl + i];
        else if (p.io.npc_actions) a = p.io.npc_actions[e * n + i];
        else {
          const int m = i - nl;                        // NPC m: Philox call m/2, words 2(m%2) [perturbation], 2(m%2)+1 [action]
          if ((m & 1) == 0 || i == nl) {
            w[0] = (uint32_t)gid; w[1] = (uint32_t)(gid >> 32); w[2] = st.tick; w[3] = (uint32_t)(m >> 1);
            gw::philox4x32(w, (uint32_t)T->seed, (uint32_t)(T->seed >> 32));
          }
          const bool pert = w[2 * (m & 1)] < T->perturb_thr;
          const uint32_t* thr = T->thr[__ldg(&T->policy_map[ci])][pert ? 1 : 0];
          const uint32_t u = w[2 * (m & 1) + 1] >> 1;
          a = 0;
          for (int k = 0; k < 8; ++k) a += (u >= thr[k]);
        }
        a = a < 0 ? 0 : (a > 8 ? 8 : a);
        s.act[t][i] = (uint8_t)a;
      }
      if (FEAR) {
        for (int i = 0; i < 2 * NA * 2; i += 4) *reinterpret_cast<uint32_t*>(&s.cnt[t][0][0][0] + i) = 0u;
        for (int x = 0; x < nl; ++x) {
          uint32_t m = 0;
          for (int k = 0; k < n; ++k)
            if (k == x || manhattan(st.cell[x], st.cell[k]) <= T->fear_radius) m |= 1u << k;
          s.close[t][x] = (uint16_t)m;
          // which (actor, affected) pairs need simulating: none when the action IS the MdR (both variants alike: Resp = 0
          // exactly), none in the single env without a close agent (customenv.py:117-118), and only the agents linked to the
          // actor by a chain of near pairs (the others' counts cannot depend on the actor's action: Resp = 0 exactly)
          if (s.act[t][x] == s.mdr[t][x]) continue;
          if (T->kind == GW_ENV_SINGLE && __popc(m) <= 1) continue;
          const uint32_t comp = p.literal_fear ? 0xFFFFu : near_component(n, st.cell, x);
          for (int j = 0; j < n; ++j)
            if (j != x && ((comp >> j) & 1u)) {
              const int slot = atomicAdd(&s.n_tasks, 1);
              s.task[slot] = (uint16_t)((t << 5) | (x << 4) | j);
              s.task_comp[slot] = (uint16_t)comp;
            }
        }
      }
    }
    __syncthreads();
    // ---- P2: FeAR_4_one_actor's counterfactual counts (Responsibility.py:163-198), one work item = nine world updates
    if (FEAR) {
      const int n_items = 2 * s.n_tasks;
      for (int it = t; it < n_items; it += THREADS) {
        const int task = s.task[it >> 1], v = it & 1;
        const int q = task >> 5, x = (task >> 4) & 1, j = task & 15;
        const uint32_t in_list = s.close[q][x];
        uint8_t la[NA];
        for (int i = 0; i < n; ++i) la[i] = s.act[q][i];
        if (v == 0) la[x] = s.mdr[q][x];
        s.cnt[q][x][j][v] = (uint8_t)(p.literal_fear ? count_valid(T, s.rows, n, s.cell[q], la, in_list, j)
                                                      : count_valid_fast(T, s.rows, n, s.cell[q], la, in_list, j, s.task_comp[it >> 1]));
      }
      __syncthreads();
      if (t == 0) { s.stat[ST_TASKS] += (unsigned long long)s.n_tasks; s.n_tasks = 0; }
    }
    // ---- P3: responsibility sums, the real update, rewards / flags (ma_customenv.py:254-302, customenv.py:124-158)
    uint8_t render = 0;
    if (live) {
      double fear[2] = {0.0, 0.0};
      if (FEAR)
        for (int x = 0; x < nl; ++x) {
          if (s.act[t][x] == s.mdr[t][x]) continue;
          if (T->kind == GW_ENV_SINGLE && __popc((uint32_t)s.close[t][x]) <= 1) continue;
          double row[NA];
          for (int j = 0; j < n; ++j) row[j] = j == x ? 0.0 : resp_of(T, s.cnt[t][x][j][0], s.cnt[t][x][j][1]);
          fear[x] = np_sum_matrix(n, x, row);
        }
      uint16_t apples[2] = {(uint16_t)((T->apple_r[0] << 8) | (T->apple_c[0] & 255)), (uint16_t)((T->apple_r[1] << 8) | (T->apple_c[1] & 255))};
      UpdateOut u;
      if (p.literal_fear) world_update(T, s.rows, n, st.cell, s.act[t], apples, (int)(st.flags & 3u), nl, u);
      else world_update_fast(T, s.rows, n, st.cell, s.act[t], apples, (int)(st.flags & 3u), nl, u);
      for (int i = 0; i < n; ++i) st.cell[i] = u.loc[i];
      double reward[2] = {0.0, 0.0};
      int term_now[2] = {0, 0}, trunc_now = 0, apples_rewarded = 0, crash_count = 0, shaped[2] = {0, 0};
      if (T->kind == GW_ENV_MULTI) {
        int ri[2] = {0, 0};
        for (int k = 0; k < nl; ++k)                                                  // :258-275 (only the own apple counts)
          if (u.caught[k][k] > 0 && ((st.flags >> k) & 1u)) {
            st.flags &= ~(1u << k);
            ri[k] += 20;
            ++apples_rewarded;
            if (!(st.flags & 3u)) {
              for (int q = 0; q < nl; ++q) ri[q] += 20;
              st.flags |= 1u << 4;
            }
          }
        int dist_valid[2] = {0, 0}, dist[2] = {0, 0};
        for (int k = 0; k < nl; ++k) {                                                // :278-300
          if ((u.crash >> k) & 1u) { ri[k] -= 10; ++crash_count; st.flags |= (1u << 4) | (1u << (2 + k)); }
          if ((st.flags >> k) & 1u) { dist_valid[k] = 1; dist[k] = manhattan(st.cell[k], apples[k]); }
          if (((st.flags >> (5 + k)) & 1u) && dist_valid[k] && (int)st.pd[k] > dist[k]) { ri[k] += 1; shaped[k] = 1; }
        }
        for (int k = 0; k < nl; ++k) {
          st.flags = (st.flags & ~(1u << (5 + k))) | ((uint32_t)dist_valid[k] << (5 + k));   // :302
          st.pd[k] = (uint16_t)dist[k];
          reward[k] = (double)ri[k];
          term_now[k] = (st.flags >> (2 + k)) & 1u;
        }
        trunc_now = (st.flags >> 4) & 1u;
      } else {                                                                        // customenv.py:126-158
        double rew = 0.0;
        const int d = manhattan(st.cell[0], apples[0]);
        if (u.crash & 1u) { rew -= 10.0; term_now[0] = 1; crash_count = 1; }
        if ((st.flags & 1u) && u.caught[0][0] == 1) { st.flags &= ~1u; rew += 20.0; trunc_now = 1; apples_rewarded = 1; }
        if (d < (int)st.pd[0]) { rew = __dadd_rn(rew, 0.1); shaped[0] = 1; }
        reward[0] = rew;
        st.pd[0] = (uint16_t)d; st.flags |= 1u << 5;
      }
      st.steps += 1;
      if (st.steps > 0xFFFu) st.steps = 0xFFFu;
      const int over = T->kind == GW_ENV_MULTI ? trunc_now : (term_now[0] || trunc_now);
      const int ended = over || (T->max_steps > 0 && (int)st.steps >= T->max_steps);
      for (int k = 0; k < nl; ++k) {
        const long long o = e * nl + k;
        if (p.io.reward) p.io.reward[o] = (float)reward[k];
        if (p.io.fear) p.io.fear[o] = fear[k];
        if (p.io.shaped_reward) p.io.shaped_reward[o] = (float)__dadd_rn(__dmul_rn(T->fear_weight, fear[k]), reward[k]);   // maddpg/agent.py:130
        if (p.io.terminated) p.io.terminated[o] = (uint8_t)term_now[k];
        if (p.io.truncated) p.io.truncated[o] = (uint8_t)(trunc_now ? 1 : 0);
      }
      write_positions(p.io.positions, e, n, st.cell);
      if (p.io.ended) p.io.ended[e] = (uint8_t)ended;
      if (p.io.info) {
        uint32_t bits = (u.crash & 15u) | ((u.restr & 15u) << 4) | ((u.crash & 0xFFFFu) << 16);
        bits |= (uint32_t)crash_count << 8 | (uint32_t)apples_rewarded << 10 | (uint32_t)ended << 12 |
                (uint32_t)u.unresolved << 13 | (uint32_t)shaped[0] << 14 | (uint32_t)shaped[1] << 15;
        p.io.info[e] = bits;
      }
      const double unit = T->kind == GW_ENV_MULTI ? 1.0 : 10.0;
      st.ep_ret[0] += (int)llrint(reward[0] * unit);
      st.ep_ret[1] += (int)llrint(reward[1] * unit);
      if (ended) {
        atomicAdd(&s.stat[ST_EPISODES], 1ull);
        atomicAdd(&s.stat[ST_LEN], (unsigned long long)st.steps);
        atomicAdd(&s.stat[ST_RETURN_UNITS], (unsigned long long)(long long)(st.ep_ret[0] + st.ep_ret[1]));
      }
      if (crash_count) atomicAdd(&s.stat[ST_CRASH], (unsigned long long)crash_count);
      if (apples_rewarded) atomicAdd(&s.stat[ST_APPLES], (unsigned long long)apples_rewarded);
      if (u.unresolved) atomicAdd(&s.stat[ST_UNRESOLVED], 1ull);
      if (FEAR) {
        const int nz = (fear[0] != 0.0) + (fear[1] != 0.0);
        if (nz) {
          atomicAdd(&s.stat[ST_FEAR_NZ], (unsigned long long)nz);
          atomicAdd(reinterpret_cast<double*>(&p.stats[(blockIdx.x % STAT_SLOTS) * ST_N + ST_FEAR_BITS]), fear[0] + fear[1]);
        }
      }
      render = 0x80u;
      if (ended && T->auto_reset) {
        for (int i = 0; i < NA; ++i) s.cell_final[t][i] = st.cell[i];
        render |= 2u | ((st.flags & 3u) << 4);
        spawn_cells(T, e, st.tick, p.io.spawn, st.cell);
        fresh_env(T, st);
        render |= 1u;
      }
      render |= (st.flags & 3u) << 2;
      for (int i = 0; i < NA; ++i) s.cell[t][i] = st.cell[i];
      write_masks(T, s.rows, p.io.action_mask, e, st.cell);
      st.tick += 1;
      p.state[e] = st;
    }
    if (t < TILE) s.render[t] = render;
    __syncthreads();
    // ---- P4: observation rows, a warp per env
    for (int q = warp; q < TILE; q += THREADS / 32) {
      const uint8_t r = s.render[q];
      if (!(r & 0x80u)) continue;
      if (stage) {
        if ((r & 2u) && p.io.final_obs) render_staged(sc, stage, p.io.final_obs, base + q, s.cell_final[q], (r >> 4) & 3u, false, lane);
        if (p.io.obs) render_staged(sc, stage, p.io.obs, base + q, s.cell[q], (r >> 2) & 3u, (r & 1u) != 0, lane);
      } else {
        if ((r & 2u) && p.io.final_obs) render_rows(T, s.rows, p.io.final_obs, base + q, s.cell_final[q], (r >> 4) & 3u, false, lane);
        if (p.io.obs) render_rows(T, s.rows, p.io.obs, base + q, s.cell[q], (r >> 2) & 3u, (r & 1u) != 0, lane);
      }
    }
    __syncthreads();
  }
  if (t < ST_N && t != ST_FEAR_BITS && s.stat[t]) atomicAdd(&p.stats[(blockIdx.x % STAT_SLOTS) * ST_N + t], s.stat[t]);
}

// ------------------------------------------------------------------ operator level
__device__ __forceinline__ int load_case(long long c, int n, const int8_t* pos, const int8_t* act, uint16_t* loc, uint8_t* a) {
  for (int i = 0; i < n; ++i) {
    loc[i] = (uint16_t)(((int)pos[(c * NA + i) * 2] << 8) | (uint8_t)pos[(c * NA + i) * 2 + 1]);
    int v = act ? act[c * NA + i] : 0;
    a[i] = (uint8_t)(v < 0 ? 0 : (v > 8 ? 8 : v));
  }
  return n;
}
__device__ __forceinline__ uint32_t load_list(long long c, int n, const uint8_t* in_list) {
  uint32_t m = 0;
  for (int i = 0; i < n; ++i)
    if (!in_list || in_list[c * NA + i]) m |= 1u << i;
  return m;
}

__global__ void __launch_bounds__(128) gww_update_world_kernel(const Tab* T, long long C, const int8_t* n_per, const int8_t* pos,
                                                               const int8_t* act, const int8_t* apples, int8_t* new_pos,
                                                               uint8_t* crash, uint8_t* restr, int8_t* caught) {
  __shared__ unsigned long long rows[GWW_MAX_DIM];
  for (int i = threadIdx.x; i < GWW_MAX_DIM; i += blockDim.x) rows[i] = T->map_rows[i];
  __syncthreads();
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int n = n_per ? n_per[c] : T->n;
  uint16_t loc[NA], ap[2] = {0xFFFFu, 0xFFFFu};
  uint8_t a[NA];
  load_case(c, n, pos, act, loc, a);
  int on = 0;
  if (apples)
    for (int k = 0; k < 2; ++k)
      if (apples[(c * 2 + k) * 2] >= 0) { on |= 1 << k; ap[k] = (uint16_t)(((int)apples[(c * 2 + k) * 2] << 8) | (uint8_t)apples[(c * 2 + k) * 2 + 1]); }
  UpdateOut u;
  world_update(T, rows, n, loc, a, apples ? ap : nullptr, on, n < 2 ? n : 2, u);
  for (int i = 0; i < NA; ++i) {
    new_pos[(c * NA + i) * 2] = i < n ? (int8_t)(u.loc[i] >> 8) : (int8_t)-1;
    new_pos[(c * NA + i) * 2 + 1] = i < n ? (int8_t)(u.loc[i] & 255) : (int8_t)-1;
    crash[c * NA + i] = i < n ? (uint8_t)((u.crash >> i) & 1u) : (uint8_t)0;
    restr[c * NA + i] = i < n ? (uint8_t)((u.restr >> i) & 1u) : (uint8_t)0;
  }
  if (caught)
    for (int e = 0; e < 2; ++e)
      for (int k = 0; k < 2; ++k) caught[c * 4 + e * 2 + k] = (int8_t)u.caught[e][k];
}

// mode 0: FeAR_4_one_actor (thread per (case, affected)); mode 1: FeAR all actors (thread per (case, actor, affected));
// mode 2: FeAL (thread per (case, agent)).
__global__ void __launch_bounds__(128) gww_resp_kernel(const Tab* T, int mode, long long C, const int8_t* n_per, const int8_t* pos,
                                                       const int8_t* act, const int8_t* mdr, const int8_t* actor,
                                                       const uint8_t* in_list, double* resp, int8_t* n_mdr, int8_t* n_act) {
  __shared__ unsigned long long rows[GWW_MAX_DIM];
  for (int i = threadIdx.x; i < GWW_MAX_DIM; i += blockDim.x) rows[i] = T->map_rows[i];
  __syncthreads();
  const int per_case = mode == 1 ? NA * NA : NA;
  const long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= C * per_case) return;
  const long long c = gidx / per_case;
  const int sub = (int)(gidx - c * per_case);
  const int n = n_per ? n_per[c] : T->n;
  const int ii = mode == 1 ? sub / NA : (mode == 0 ? actor[c] : sub);
  const int jj = mode == 1 ? sub % NA : sub;
  double r = 0.0;
  int m = 0, v = 0;
  if (ii < n && jj < n && (mode == 2 || ii != jj)) {
    uint16_t loc[NA];
    uint8_t a[NA], la[NA];
    load_case(c, n, pos, act, loc, a);
    uint32_t il = load_list(c, n, in_list);
    if (mode == 0) il |= 1u << ii;                                                   // the actor's entry is forced on (the env's close list)
    for (int i = 0; i < n; ++i) la[i] = a[i];
    if (mode == 2) {                                                                  // FeAL :251-270: the others play their MdR
      for (int i = 0; i < n; ++i)
        if (i != ii) { const int q = mdr[c * NA + i]; la[i] = (uint8_t)(q < 0 ? 0 : (q > 8 ? 8 : q)); }
    } else if ((il >> ii) & 1u) {                                                     // an actor outside the list cannot be swapped
      const int q = mdr[c * NA + ii];
      la[ii] = (uint8_t)(q < 0 ? 0 : (q > 8 ? 8 : q));
    }
    m = count_valid(T, rows, n, loc, la, il, jj);
    v = count_valid(T, rows, n, loc, a, il, jj);
    if (mode == 2) {
      const double x = (double)v / ((double)m + 0.000001);                            // :287-290
      r = x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x);
    } else r = resp_of(T, m, v);
  }
  resp[gidx] = r;
  n_mdr[gidx] = (int8_t)m;
  n_act[gidx] = (int8_t)v;
}

__global__ void __launch_bounds__(128) gww_fear_sum_kernel(long long C, const int8_t* n_per, int n_default, const int8_t* actor,
                                                           const double* resp, double* fear_sum) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int n = n_per ? n_per[c] : n_default;
  double row[NA];
  for (int j = 0; j < NA; ++j) row[j] = j < n ? resp[c * NA + j] : 0.0;
  fear_sum[c] = np_sum_matrix(n, actor[c], row);
}

}  // namespace gww

// ------------------------------------------------------------------ host side

struct gww_handle {
  gww_config cfg;
  gww::Tab* d_tab = nullptr;
  gww::EnvState* d_state = nullptr;
  unsigned long long* d_stats = nullptr;
  uint8_t* d_tmpl = nullptr;
  bool reset_done = false;
  int sm_count = 148;
  uint64_t launches = 0, env_steps = 0;
  std::string err;
};

static std::string g_wide_create_error;

static int wfail(gww_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_wide_create_error = msg;
  return code;
}
static int wcuda(gww_handle* h, cudaError_t e, const char* what) {
  return wfail(h, GW_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define GWW_CUDA(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return wcuda(h, e_, #call); } while (0)

static int wide_validate(const gww_config* c, std::string& why) {
  auto bad = [&](const char* m) { why = m; return GW_EINVAL; };
  if (c->struct_size != (int32_t)sizeof(gww_config)) return bad("gww_config.struct_size mismatch");
  if (c->abi_version != GW_ABI_VERSION) return bad("gww_config.abi_version mismatch");
  if (c->height < 1 || c->height > GWW_MAX_DIM || c->width < 1 || c->width > GWW_MAX_DIM) return bad("height / width must be 1..64");
  if (c->n_agents < 1 || c->n_agents > GWW_MAX_AGENTS) return bad("n_agents must be 1..16");
  if (c->n_learners < 1 || c->n_learners > GW_MAX_LEARNERS || c->n_learners > c->n_agents) return bad("n_learners must be 1..2 and <= n_agents");
  if (c->env_kind != GW_ENV_MULTI && c->env_kind != GW_ENV_SINGLE) return bad("env_kind");
  if (c->env_kind == GW_ENV_SINGLE && c->n_learners != 1) return bad("the single-learner env has n_learners = 1");
  if (c->n_policies < 1 || c->n_policies > GW_MAX_POLICIES) return bad("n_policies must be 1..16");
  if (c->num_envs < 1) return bad("num_envs must be >= 1");
  if (c->obs_dtype != GW_OBS_F32 && c->obs_dtype != GW_OBS_BF16) return bad("obs_dtype");
  if (c->n_blocked < 0 || c->n_blocked > GWW_MAX_BLOCKED) return bad("n_blocked out of range");
  if (c->max_steps < 0 || c->fear_radius < 0) return bad("max_steps / fear_radius must be >= 0");
  if (!(c->perturb_prob >= 0.0 && c->perturb_prob <= 1.0)) return bad("perturb_prob must be in [0, 1]");
  int active = 0;
  for (int r = 0; r < c->height; ++r)
    for (int col = 0; col < c->width; ++col) active += (int)((c->map_rows[r] >> col) & 1ull);
  if (active < c->n_agents) return bad("fewer active cells than agents");
  for (int i = 0; i < c->height * c->width; ++i) {
    if (c->policy_map[i] >= c->n_policies) return bad("policy_map entry >= n_policies");
    if (c->mdr_map[i] >= GW_N_ACTIONS) return bad("mdr_map entry is not an action id");
  }
  for (int k = 0; k < c->n_learners; ++k) {
    const int r = c->apple_row[k], col = c->apple_col[k];
    if (c->env_kind == GW_ENV_SINGLE && r < 0) return bad("the single-learner env needs its apple");
    if (r >= 0 && (r >= c->height || col < 0 || col >= c->width || !((c->map_rows[r] >> col) & 1ull))) return bad("apple on an inactive cell");
  }
  return GW_OK;
}

static gww::Params wide_params(gww_handle* h, const gw_io* io) {
  gww::Params p;
  p.T = h->d_tab; p.state = h->d_state; p.stats = h->d_stats; p.E = h->cfg.num_envs; p.io = *io; p.reset_mask = nullptr;
  p.tmpl = h->d_tmpl;
  // the observation block of one env leaves as one bulk copy when its size and the destinations allow 16-byte granules
  const size_t block = (size_t)h->cfg.n_learners * h->cfg.height * h->cfg.width * (h->cfg.obs_dtype == GW_OBS_BF16 ? 2 : 4);
  const bool aligned = (reinterpret_cast<uintptr_t>(io->obs) % 16 == 0) && (reinterpret_cast<uintptr_t>(io->final_obs) % 16 == 0);
  const char* scalar = std::getenv("GWW_SCALAR_ROWS");
  p.stage_bytes = (block % 16 == 0 && aligned && !(scalar && scalar[0] == '1')) ? (int)block : 0;
  const char* lit = std::getenv("GWW_FEAR_LITERAL");
  p.literal_fear = (lit && lit[0] == '1') ? 1 : 0;
  return p;
}

extern "C" {

int gww_default_config(gww_config* cfg) {
  if (!cfg) return GW_EINVAL;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->struct_size = (int32_t)sizeof(gww_config);
  cfg->abi_version = GW_ABI_VERSION;
  cfg->n_learners = 2;
  cfg->env_kind = GW_ENV_MULTI;
  cfg->apple_row[0] = 9; cfg->apple_col[0] = 0;      // ma_customenv.py:422
  cfg->apple_row[1] = 5; cfg->apple_col[1] = 10;
  cfg->n_policies = 1;
  cfg->step_weights[0][0] = cfg->step_weights[0][1] = cfg->step_weights[0][2] = 1.0f;
  for (int d = 0; d < 4; ++d) cfg->dir_weights[0][d] = 1.0f;
  cfg->perturb_prob = 0.25;                          // ma_customenv.py:441
  cfg->fear = 1;
  cfg->fear_radius = 5;                              // :249
  cfg->max_steps = 150;                              // configs/custom.yaml:5
  cfg->auto_reset = 1;
  cfg->num_envs = 1;
  return GW_OK;
}

const char* gww_last_error(const gww_handle* h) { return h ? h->err.c_str() : g_wide_create_error.c_str(); }

int gww_create(const gww_config* cfg, gww_handle** out) {
  if (!cfg || !out) return wfail(nullptr, GW_EINVAL, "gww_create: null argument");
  *out = nullptr;
  std::string why;
  if (wide_validate(cfg, why) != GW_OK) return wfail(nullptr, GW_EINVAL, "gww_create: " + why);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return wfail(nullptr, GW_ENODEV, "gww_create: no CUDA device (this library has no CPU fallback)");
  }
  if (cfg->device < 0 || cfg->device >= n_dev) return wfail(nullptr, GW_EINVAL, "gww_create: bad device ordinal");
  gww_handle* h = new (std::nothrow) gww_handle();
  if (!h) return wfail(nullptr, GW_ENOMEM, "gww_create: host allocation failed");
  h->cfg = *cfg;
  gww::Tab* T = new (std::nothrow) gww::Tab();
  if (!T) { delete h; return wfail(nullptr, GW_ENOMEM, "gww_create: host allocation failed"); }
  std::memset(T, 0, sizeof(*T));
  T->H = cfg->height; T->W = cfg->width; T->n = cfg->n_agents; T->nl = cfg->n_learners; T->kind = cfg->env_kind;
  T->fear = cfg->fear; T->fear_radius = cfg->fear_radius; T->max_steps = cfg->max_steps; T->auto_reset = cfg->auto_reset;
  T->obs_bf16 = cfg->obs_dtype == GW_OBS_BF16;
  for (int k = 0; k < 2; ++k) { T->apple_r[k] = k < cfg->n_learners ? cfg->apple_row[k] : -1; T->apple_c[k] = k < cfg->n_learners ? cfg->apple_col[k] : -1; }
  const double pthr = cfg->perturb_prob * 4294967296.0;
  T->perturb_thr = pthr >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)pthr;
  T->fear_weight = cfg->fear_weight; T->seed = cfg->seed; T->env_id_base = cfg->env_id_base;
  for (int r = 0; r < cfg->height; ++r) {
    T->map_rows[r] = cfg->width == 64 ? cfg->map_rows[r] : (cfg->map_rows[r] & ((1ull << cfg->width) - 1ull));
    for (int c = 0; c < cfg->width; ++c)
      if ((T->map_rows[r] >> c) & 1ull) T->active[T->n_active++] = (uint16_t)((r << 8) | c);
  }
  for (int p = 0; p < cfg->n_policies; ++p) {
    gw_policy_thresholds(cfg->step_weights[p], cfg->dir_weights[p], false, T->thr[p][0]);
    gw_policy_thresholds(cfg->step_weights[p], cfg->dir_weights[p], true, T->thr[p][1]);
  }
  for (int m = 0; m < 10; ++m)
    for (int a = 0; a < 10; ++a) {
      const double r = ((double)m - (double)a) / ((double)m + 0.000001);
      T->resp_lut[m][a] = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
    }
  std::memcpy(T->policy_map, cfg->policy_map, sizeof(T->policy_map));
  std::memcpy(T->mdr_map, cfg->mdr_map, sizeof(T->mdr_map));
  // restricted paths (GWorld.__init__, grid_world.py:32-86): only moves between 4-neighbours inside the grid can match a path
  for (int k = 0; k < cfg->n_blocked; ++k) {
    const int fr = cfg->blocked_from[k] >> 8, fc = cfg->blocked_from[k] & 255, tr = cfg->blocked_to[k] >> 8, tc = cfg->blocked_to[k] & 255;
    if (fr >= cfg->height || tr >= cfg->height || fc >= cfg->width || tc >= cfg->width) continue;
    int d = -1;
    if (tc == fc && tr == fr - 1) d = 0; else if (tc == fc && tr == fr + 1) d = 1;
    else if (tr == fr && tc == fc - 1) d = 2; else if (tr == fr && tc == fc + 1) d = 3;
    if (d >= 0) T->blocked_dir[fr * cfg->width + fc] |= (uint8_t)(1u << d);
  }
  auto cleanup = [&](int code, const std::string& msg) { delete T; gww_destroy(h); return wfail(nullptr, code, msg); };
  cudaError_t e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return cleanup(GW_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
  if ((e = cudaMalloc(&h->d_tab, sizeof(gww::Tab))) != cudaSuccess ||
      (e = cudaMalloc(&h->d_state, sizeof(gww::EnvState) * (size_t)cfg->num_envs)) != cudaSuccess ||
      (e = cudaMalloc(&h->d_stats, sizeof(unsigned long long) * gww::STAT_SLOTS * gww::ST_N)) != cudaSuccess)
    return cleanup(e == cudaErrorMemoryAllocation ? GW_ENOMEM : GW_ECUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  if ((e = cudaMemcpy(h->d_tab, T, sizeof(gww::Tab), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemset(h->d_state, 0, sizeof(gww::EnvState) * (size_t)cfg->num_envs)) != cudaSuccess ||
      (e = cudaMemset(h->d_stats, 0, sizeof(unsigned long long) * gww::STAT_SLOTS * gww::ST_N)) != cudaSuccess)
    return cleanup(GW_ECUDA, std::string("initialising device tables: ") + cudaGetErrorString(e));
  {
    const int len = cfg->height * cfg->width, elt = cfg->obs_dtype == GW_OBS_BF16 ? 2 : 4;
    std::string tm((size_t)cfg->n_learners * len * elt + 16, '\0');
    for (int k = 0; k < cfg->n_learners; ++k)
      for (int i = 0; i < len; ++i) {
        const bool on = (T->map_rows[i / cfg->width] >> (i % cfg->width)) & 1ull;
        if (elt == 4) { const float v = on ? 0.0f : -1.0f; std::memcpy(&tm[((size_t)k * len + i) * 4], &v, 4); }
        else { const uint16_t v = on ? 0x0000u : 0xBF80u; std::memcpy(&tm[((size_t)k * len + i) * 2], &v, 2); }   // bf16 -1.0
      }
    if ((e = cudaMalloc(&h->d_tmpl, tm.size())) != cudaSuccess || (e = cudaMemcpy(h->d_tmpl, tm.data(), tm.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
      return cleanup(GW_ECUDA, std::string("observation template: ") + cudaGetErrorString(e));
  }
  delete T;
  const int max_dyn = (int)(gww::smem_fixed<128>() + (size_t)(gww::THREADS / 32) * GW_MAX_LEARNERS * GWW_MAX_CELLS * 4);   // 150 KB at 64 x 64 fp32
  cudaFuncSetAttribute(gww::gww_step_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
  cudaFuncSetAttribute(gww::gww_step_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
  cudaFuncSetAttribute(gww::gww_step_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
  cudaFuncSetAttribute(gww::gww_step_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
  // thread-per-env phases keep their paths in local memory: leave half of the SM's array to L1 (a hint; the driver raises the
  // shared-memory share when a block needs more)
  cudaFuncSetAttribute(gww::gww_step_kernel<false, 32>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
  cudaFuncSetAttribute(gww::gww_step_kernel<false, 128>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
  cudaFuncSetAttribute(gww::gww_reset_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
  cudaFuncSetAttribute(gww::gww_reset_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
  *out = h;
  return GW_OK;
}

int gww_destroy(gww_handle* h) {
  if (!h) return GW_OK;
  cudaSetDevice(h->cfg.device);
  if (h->d_tab) cudaFree(h->d_tab);
  if (h->d_state) cudaFree(h->d_state);
  if (h->d_stats) cudaFree(h->d_stats);
  if (h->d_tmpl) cudaFree(h->d_tmpl);
  delete h;
  return GW_OK;
}

static int wide_tile(const gww_handle* h) { return h->cfg.num_envs <= 16384 ? 32 : 128; }
static unsigned wide_blocks(const gww_handle* h) {
  const int tile = wide_tile(h);
  const long long tiles = (h->cfg.num_envs + tile - 1) / tile;
  const long long cap = (long long)h->sm_count * 8;
  return (unsigned)(tiles < cap ? tiles : cap);
}

int gww_reset(gww_handle* h, const uint8_t* reset_mask, const gw_io* io, void* stream) {
  if (!h) return GW_EINVAL;
  if (!io) return wfail(h, GW_EINVAL, "gww_reset: io is null");
  if (!h->reset_done && reset_mask) return wfail(h, GW_ESTATE, "gww_reset: the first reset must cover all envs (reset_mask = NULL)");
  GWW_CUDA(h, cudaSetDevice(h->cfg.device));
  gww::Params p = wide_params(h, io);
  p.reset_mask = reset_mask;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t stage = (size_t)(gww::THREADS / 32) * p.stage_bytes;
  if (wide_tile(h) == 32) gww::gww_reset_kernel<32><<<wide_blocks(h), gww::THREADS, gww::smem_fixed<32, false>() + stage, s>>>(p);
  else gww::gww_reset_kernel<128><<<wide_blocks(h), gww::THREADS, gww::smem_fixed<128, false>() + stage, s>>>(p);
  GWW_CUDA(h, cudaGetLastError());
  h->reset_done = true;
  h->launches += 1;
  return GW_OK;
}

int gww_step(gww_handle* h, const gw_io* io, void* stream) {
  if (!h) return GW_EINVAL;
  if (!io || !io->learner_actions) return wfail(h, GW_EINVAL, "gww_step: io->learner_actions is required");
  if (!h->reset_done) return wfail(h, GW_ESTATE, "gww_step: call gww_reset first");
  GWW_CUDA(h, cudaSetDevice(h->cfg.device));
  gww::Params p = wide_params(h, io);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = wide_blocks(h);
  const size_t stage = (size_t)(gww::THREADS / 32) * p.stage_bytes;
  if (wide_tile(h) == 32) {
    if (h->cfg.fear) gww::gww_step_kernel<true, 32><<<blocks, gww::THREADS, gww::smem_fixed<32>() + stage, s>>>(p);
    else gww::gww_step_kernel<false, 32><<<blocks, gww::THREADS, gww::smem_fixed<32, false>() + stage, s>>>(p);
  } else {
    if (h->cfg.fear) gww::gww_step_kernel<true, 128><<<blocks, gww::THREADS, gww::smem_fixed<128>() + stage, s>>>(p);
    else gww::gww_step_kernel<false, 128><<<blocks, gww::THREADS, gww::smem_fixed<128, false>() + stage, s>>>(p);
  }
  GWW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  h->env_steps += (uint64_t)h->cfg.num_envs;
  return GW_OK;
}

int gww_sync(gww_handle* h, void* stream) {
  if (!h) return GW_EINVAL;
  GWW_CUDA(h, cudaSetDevice(h->cfg.device));
  GWW_CUDA(h, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  GWW_CUDA(h, cudaGetLastError());
  return GW_OK;
}

size_t gww_state_bytes(const gww_handle* h) { return h ? sizeof(gww::EnvState) * (size_t)h->cfg.num_envs : 0; }

int gww_get_state(gww_handle* h, void* dst, int dst_is_device, void* stream) {
  if (!h || !dst) return wfail(h, GW_EINVAL, "gww_get_state: null argument");
  GWW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GWW_CUDA(h, cudaMemcpyAsync(dst, h->d_state, gww_state_bytes(h), dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  if (!dst_is_device) GWW_CUDA(h, cudaStreamSynchronize(s));
  return GW_OK;
}

int gww_set_state(gww_handle* h, const void* src, int src_is_device, void* stream) {
  if (!h || !src) return wfail(h, GW_EINVAL, "gww_set_state: null argument");
  GWW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GWW_CUDA(h, cudaMemcpyAsync(h->d_state, src, gww_state_bytes(h), src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  if (!src_is_device) GWW_CUDA(h, cudaStreamSynchronize(s));
  h->reset_done = true;
  return GW_OK;
}

int gww_get_stats(gww_handle* h, gw_stats* out, void* stream) {
  if (!h || !out) return wfail(h, GW_EINVAL, "gww_get_stats: null argument");
  GWW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long host[gww::STAT_SLOTS * gww::ST_N];
  GWW_CUDA(h, cudaMemcpyAsync(host, h->d_stats, sizeof(host), cudaMemcpyDeviceToHost, s));
  GWW_CUDA(h, cudaStreamSynchronize(s));
  std::memset(out, 0, sizeof(*out));
  long long ret_units = 0;
  double fear_sum = 0;
  for (int sl = 0; sl < gww::STAT_SLOTS; ++sl) {
    const unsigned long long* r = host + sl * gww::ST_N;
    out->episodes += r[gww::ST_EPISODES];
    out->episode_len_sum += r[gww::ST_LEN];
    out->crashes += r[gww::ST_CRASH];
    out->apples += r[gww::ST_APPLES];
    out->unresolved += r[gww::ST_UNRESOLVED];
    out->fear_nonzero += r[gww::ST_FEAR_NZ];
    out->fear_tasks += r[gww::ST_TASKS];
    ret_units += (long long)r[gww::ST_RETURN_UNITS];
    double f;
    std::memcpy(&f, &r[gww::ST_FEAR_BITS], 8);
    fear_sum += f;
  }
  out->env_steps = h->env_steps;
  out->agent_steps = h->env_steps * (uint64_t)h->cfg.n_learners;
  out->return_sum = (double)ret_units / (h->cfg.env_kind == GW_ENV_MULTI ? 1.0 : 10.0);
  out->fear_sum = fear_sum;
  return GW_OK;
}

int gww_reset_stats(gww_handle* h, void* stream) {
  if (!h) return GW_EINVAL;
  GWW_CUDA(h, cudaSetDevice(h->cfg.device));
  GWW_CUDA(h, cudaMemsetAsync(h->d_stats, 0, sizeof(unsigned long long) * gww::STAT_SLOTS * gww::ST_N, static_cast<cudaStream_t>(stream)));
  h->env_steps = 0;
  return GW_OK;
}

int gww_launch_count(const gww_handle* h, uint64_t* n) {
  if (!h || !n) return GW_EINVAL;
  *n = h->launches;
  return GW_OK;
}

int gww_update_world(gww_handle* h, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
                     const int8_t* apples, int8_t* new_positions, uint8_t* crash, uint8_t* restricted, int8_t* caught, void* stream) {
  if (!h) return GW_EINVAL;
  if (n_cases < 0 || !positions || !actions || !new_positions || !crash || !restricted) return wfail(h, GW_EINVAL, "gww_update_world: null argument");
  if (n_cases == 0) return GW_OK;
  GWW_CUDA(h, cudaSetDevice(h->cfg.device));
  gww::gww_update_world_kernel<<<(unsigned)((n_cases + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d_tab, n_cases, n_per, positions, actions, apples, new_positions, crash, restricted, caught);
  GWW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

static int wide_resp(gww_handle* h, int mode, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
                     const int8_t* mdr, const int8_t* actor, const uint8_t* in_list, double* resp, int8_t* n_mdr, int8_t* n_act,
                     void* stream, const char* what) {
  if (!h) return GW_EINVAL;
  if (n_cases < 0 || !positions || !actions || !mdr || !resp || !n_mdr || !n_act || (mode == 0 && !actor))
    return wfail(h, GW_EINVAL, std::string(what) + ": null argument");
  if (n_cases == 0) return GW_OK;
  GWW_CUDA(h, cudaSetDevice(h->cfg.device));
  const long long items = (long long)n_cases * (mode == 1 ? GWW_MAX_AGENTS * GWW_MAX_AGENTS : GWW_MAX_AGENTS);
  gww::gww_resp_kernel<<<(unsigned)((items + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d_tab, mode, n_cases, n_per, positions, actions, mdr, actor, in_list, resp, n_mdr, n_act);
  GWW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gww_fear_one_actor(gww_handle* h, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
                       const int8_t* mdr, const int8_t* actor, const uint8_t* in_list, double* resp, int8_t* n_mdr, int8_t* n_act,
                       double* fear_sum, void* stream) {
  if (int rc = wide_resp(h, 0, n_cases, n_per, positions, actions, mdr, actor, in_list, resp, n_mdr, n_act, stream, "gww_fear_one_actor")) return rc;
  if (fear_sum && n_cases > 0) {
    gww::gww_fear_sum_kernel<<<(unsigned)((n_cases + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        n_cases, n_per, h->cfg.n_agents, actor, resp, fear_sum);
    GWW_CUDA(h, cudaGetLastError());
    h->launches += 1;
  }
  return GW_OK;
}

int gww_fear_matrix(gww_handle* h, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
                    const int8_t* mdr, const uint8_t* in_list, double* resp, int8_t* n_mdr, int8_t* n_act, void* stream) {
  return wide_resp(h, 1, n_cases, n_per, positions, actions, mdr, nullptr, in_list, resp, n_mdr, n_act, stream, "gww_fear_matrix");
}

int gww_feal(gww_handle* h, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
             const int8_t* mdr, const uint8_t* in_list, double* feal, int8_t* n_mdr, int8_t* n_act, void* stream) {
  return wide_resp(h, 2, n_cases, n_per, positions, actions, mdr, nullptr, in_list, feal, n_mdr, n_act, stream, "gww_feal");
}

}  // extern "C"
