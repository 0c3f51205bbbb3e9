// Device-side building blocks of the batched grid world (sm_100a).
//
// Formulation (differs from the reference's on purpose; tests/ diff it against the literal
// restatement in oracle/):
//   * A cell is one byte (row << 4) | col (W = 16).
//   * The reference grows a list of cells per agent and sub-step (custom/grid_world.py:458-518).
//     Here an agent's nominal trajectory is three cells p0, p1, p2 read from a next-cell table, and
//     is classified into one of 13 "effective trajectories" (eff): 0 stationary, 1-4 one step in
//     direction d, 5-8 two steps, 9-12 two-step action whose second move is blocked.  Blocked and
//     Stay moves are the same trajectory, and so is a crashed agent: the revert of :200-208 puts
//     every later path entry back on p0, i.e. a crashed agent is a stationary one.
//   * The five-rule pair test (:276-390) only compares cells of the two agents, so it is translation
//     invariant: for a pair it depends on (p0_j - p0_i, eff_i, eff_j) and on nothing else.  gw_create
//     evaluates the literal rules (pair_hit below) once for every |delta|_1 <= 4 (41 offsets, "diamond"
//     index) and every eff pair into a 6.9 KB table of 4-bit masks (bit s = the pair collides at
//     sub-step s).  Agents further apart than 4 can never share or swap a cell within one step.
//   * The fix-point of :247-405 then runs on bit vectors: three 24-bit words hold the masks of the six
//     pairs for (both nominal / second crashed / first crashed); one pass is a handful of logic ops.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/gridworld_b200.h"

namespace gw {

constexpr int N_EFF = 13;
constexpr int N_DELTA = 41;                         // offsets with |dr| + |dc| <= 4 in row-major order: idx(-d) = 40 - idx(d)
constexpr int LUT_BYTES = N_DELTA * N_EFF * N_EFF;  // 6929

// Everything the world update and the counterfactual counting read, in one block that every CTA copies to shared
// memory (13 KB).  Built by gw_create.
struct SimTab {
  alignas(16) uint8_t lut[(LUT_BYTES + 15) / 16 * 16];      // [delta][eff_i][eff_j] -> 4-bit sub-step hit mask (symmetric: [d][a][b] == [40-d][b][a])
  alignas(16) uint8_t next[GW_MAX_H * GW_W * 4];            // [cell][dir Up,Down,Left,Right] -> cell after the move (same cell if blocked)
  alignas(16) uint16_t rowmask4[N_DELTA * N_EFF][4];        // [delta][eff_i][s] -> bit e: lut[delta][eff_i][e] has bit s (hit at sub-step s)
  alignas(16) uint16_t unres[GW_MAX_H * GW_W];              // [cell] -> bit a: action a is not restricted from this cell
  alignas(16) uint8_t diamond[96];                          // (dr + 4) * 9 + (dc + 4) -> delta index (0xFF outside the diamond)
  alignas(16) uint16_t reach[64];                           // [near6] -> 4 nibbles: agents linked to agent x by chains of near pairs
};

struct SmallTables {                    // per-step lookups, copied to shared memory by every CTA (2592 B)
  alignas(16) double resp_lut[10][10];                      // clip((m-a)/(m+1e-6),-1,1), Responsibility.py:194-198
  alignas(16) uint32_t policy_thr[GW_MAX_POLICIES][2][8];   // 31-bit cdf thresholds, [policy][perturbed][k]
  uint8_t mdr_map[GW_MAX_H * GW_W];
  uint8_t policy_map[GW_MAX_H * GW_W];
  uint8_t active_cell[GW_MAX_H * GW_W];         // row-major list of active cells (np.where order)
  uint16_t map_rows[GW_MAX_H];                  // bit c of map_rows[r] = cell (r, c) is active
};

struct Tables {                         // device-global, read-only, built by gw_create
  SimTab sim;                           // sim + small are contiguous here and in shared memory: one TMA copy per CTA
  alignas(16) SmallTables small;
  int32_t n_active;
  int32_t pad_;
};

// ---------------------------------------------------------------- Philox4x32-10
__host__ __device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[0];
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// ---------------------------------------------------------------- the literal pair test (LUT builder)
// custom/grid_world.py:276-390, first match wins.  q = (step+1)*len in quarter sub-steps,
// f = floor(q/4), c = ceil(q/4); A = path[f], B = path[c]; P = start cell.
__host__ __device__ __forceinline__ bool pair_hit(int Ai, int Bi, int Pi, int qi, int fi, int ci,
                                                  int Aj, int Bj, int Pj, int qj, int fj, int cj) {
  if (Ai == Aj || Bi == Bj) return true;                                   // :276-278
  if (Ai == Bj && Bi == Aj) return true;                                   // :291-294
  const bool same_dir = (Bi - Ai) == (Bj - Aj);                            // :314-318 (cell codes: unique per direction)
  if (Ai == Bj) return !(((4 * ci - qi) + (qj - 4 * fj) <= 4) && same_dir);   // :307-326
  if (Bi == Aj) return !(((4 * cj - qj) + (qi - 4 * fi) <= 4) && same_dir);   // :339-357
  return (Ai == Pj && Pi == Aj) || (Bi == Pj && Pi == Bj) ||               // :371-378
         (Ai == Pj && Pi == Bj) || (Bi == Pj && Pi == Aj);
}

// ---------------------------------------------------------------- geometry
__device__ __forceinline__ int manhattan(uint32_t a, uint32_t b) {
  return abs((int)(a >> 4) - (int)(b >> 4)) + abs((int)(a & 15) - (int)(b & 15));
}

// Trajectory of one agent for action a (custom/custom_agent.py:140-150 + grid_world.py:481-518):
// off-grid (clip) or inactive target => the agent stays and the move is "restricted".
struct Traj {
  uint32_t p1, p2, eff;
  uint32_t r1, r2;          // first / second move restricted
  uint32_t two;             // two-step action
};

__device__ __forceinline__ Traj make_traj(const uint8_t* __restrict__ s_next, uint32_t p0, uint32_t a) {
  Traj t;
  t.p1 = t.p2 = p0;
  t.eff = t.r1 = t.r2 = 0;
  t.two = a >= 5 ? 1u : 0u;
  if (a != 0) {
    const uint32_t d = (a - 1) & 3u;
    t.p1 = s_next[p0 * 4 + d];
    if (t.p1 == p0) {
      t.r1 = 1;                                   // a blocked first move of a two-step action is blocked twice
    } else if (!t.two) {
      t.eff = 1 + d;
    } else {
      t.p2 = s_next[t.p1 * 4 + d];
      t.r2 = (t.p2 == t.p1) ? 1u : 0u;
      t.eff = (t.r2 ? 9u : 5u) + d;
    }
    if (!t.two) t.p2 = t.p1;
  }
  return t;
}

// get_action_mask, custom/ma_customenv.py:467-506: only the TARGET cell is tested.
__device__ __forceinline__ uint32_t action_mask_bits(const uint16_t* rows, int H, uint32_t p) {
  uint32_t m = 1u;
  const int r = (int)(p >> 4), c = (int)(p & 15);
#pragma unroll
  for (int a = 1; a < GW_N_ACTIONS; ++a) {
    const int len = a >= 5 ? 2 : 1, d = (a - 1) & 3;
    const int tr = r + (d == 0 ? -len : d == 1 ? len : 0), tc = c + (d == 2 ? -len : d == 3 ? len : 0);
    if ((unsigned)tr < (unsigned)H && (unsigned)tc < (unsigned)GW_W && ((rows[tr] >> tc) & 1)) m |= 1u << a;
  }
  return m;
}

// ---------------------------------------------------------------- pair geometry of one env
// near6: bit p = pair p (01,02,03,12,13,23) within Manhattan distance 4; didx: delta index (0..40) per pair.
struct PairGeom {
  uint32_t near6;
  uint32_t didx_lo;     // pairs 0..3, 8 bits each
  uint32_t didx_hi;     // pairs 4..5
};

__device__ __forceinline__ PairGeom pair_geometry(const SimTab& T, int n, uint32_t cells) {
  PairGeom g;
  g.near6 = g.didx_lo = g.didx_hi = 0;
  int p = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = i + 1; j < 4; ++j, ++p) {
      if (j < n) {
        const uint32_t a = (cells >> (8 * i)) & 0xFFu, b = (cells >> (8 * j)) & 0xFFu;
        const int dr = (int)(b >> 4) - (int)(a >> 4), dc = (int)(b & 15) - (int)(a & 15);
        if (abs(dr) + abs(dc) <= 4) {
          g.near6 |= 1u << p;
          const uint32_t di = T.diamond[(dr + 4) * 9 + (dc + 4)];
          if (p < 4) g.didx_lo |= di << (8 * p); else g.didx_hi |= di << (8 * (p - 4));
        }
      }
    }
  }
  return g;
}

__device__ __forceinline__ uint32_t geom_didx(const PairGeom& g, int p) {
  return p < 4 ? (g.didx_lo >> (8 * p)) & 0xFFu : (g.didx_hi >> (8 * (p - 4))) & 0xFFu;
}

// agents connected to x through chains of near pairs (only they can influence x's crash outcome)
__host__ __device__ __forceinline__ uint32_t reach_mask_slow(uint32_t near6, int x) {
  uint32_t adj = 0;                                   // 4 nibbles: adjacency row per agent
  if (near6 & 1u) adj |= (2u << 0) | (1u << 4);       // (0,1)
  if (near6 & 2u) adj |= (4u << 0) | (1u << 8);       // (0,2)
  if (near6 & 4u) adj |= (8u << 0) | (1u << 12);      // (0,3)
  if (near6 & 8u) adj |= (4u << 4) | (2u << 8);       // (1,2)
  if (near6 & 16u) adj |= (8u << 4) | (2u << 12);     // (1,3)
  if (near6 & 32u) adj |= (8u << 8) | (4u << 12);     // (2,3)
  uint32_t r = 1u << x;
  for (int it = 0; it < 3; ++it) {
    uint32_t nr = r;
    for (int a = 0; a < 4; ++a)
      if ((r >> a) & 1u) nr |= (adj >> (4 * a)) & 0xFu;
    r = nr;
  }
  return r;
}
__device__ __forceinline__ uint32_t reach_mask(const SimTab& T, uint32_t near6, int x) {
  return ((uint32_t)T.reach[near6 & 63u] >> (4 * x)) & 0xFu;
}

// ---------------------------------------------------------------- collision fix-point on bit vectors
// Returns the crashed mask after each of the four sub-steps, 4 bits each (bits 12-15 = final AgentCrash).
// custom/grid_world.py:247-405: within a pass the paths are frozen (hits are evaluated on the crashed mask
// of the pass start), crashed agents revert after the pass, passes repeat while something was hit.  Every
// pass with a hit crashes at least one more agent, so the reference's 2N-pass cap can never bind.
// `stop`: agents whose crash ends the evaluation early (counterfactuals only ask about one agent); the
// returned word is then only valid in its final nibble for those agents.
__device__ __forceinline__ uint32_t agents_of_pairs(uint32_t hp) {
  return ((hp & 0x000111u) ? 1u : 0u) | ((hp & 0x011001u) ? 2u : 0u) | ((hp & 0x101010u) ? 4u : 0u) |
         ((hp & 0x110100u) ? 8u : 0u);
}

// effp: 4 x 4-bit effective trajectories.  pair_sel: which of the six pairs to look up.
__device__ __forceinline__ uint32_t nn_masks(const SimTab& T, const PairGeom& g, uint32_t effp, uint32_t pair_sel) {
  uint32_t NNw = 0;
  const uint32_t sel = g.near6 & pair_sel;
  int p = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j, ++p)
      if ((sel >> p) & 1u)
        NNw |= (uint32_t)T.lut[geom_didx(g, p) * (N_EFF * N_EFF) + ((effp >> (4 * i)) & 0xFu) * N_EFF +
                               ((effp >> (4 * j)) & 0xFu)] << (4 * p);
  return NNw;
}

// Hit-to-hit loop of the fix-point: NNw / NRw / RNw = 4-bit sub-step hit masks of the six pairs for (both on course /
// second agent crashed / first agent crashed), s = first sub-step with a hit, crashed = the agents it crashes.
__device__ __forceinline__ uint32_t fixpoint_loop(uint32_t NNw, uint32_t NRw, uint32_t RNw, int s, uint32_t crashed,
                                                  uint32_t stop) {
  uint32_t out = (crashed * 0x1111u) & (0xFFFFu << (4 * s));
  while (true) {
    const uint32_t Pi = ((crashed & 1u) ? 0x000FFFu : 0u) | ((crashed & 2u) ? 0x0FF000u : 0u) | ((crashed & 4u) ? 0xF00000u : 0u);
    const uint32_t Pj = ((crashed & 2u) ? 0x00000Fu : 0u) | ((crashed & 4u) ? 0x00F0F0u : 0u) | ((crashed & 8u) ? 0xFF0F00u : 0u);
    const uint32_t live = ((NNw & ~Pi & ~Pj) | (NRw & ~Pi & Pj) | (RNw & Pi & ~Pj)) & (((0xFu << s) & 0xFu) * 0x111111u);
    if (live == 0) break;
    s = __ffs((live | (live >> 4) | (live >> 8) | (live >> 12) | (live >> 16) | (live >> 20)) & 0xFu) - 1;
    crashed |= agents_of_pairs((live >> s) & 0x111111u);
    if (crashed & stop) return crashed << 12;
    out |= (crashed * 0x1111u) & (0xFFFFu << (4 * s));
  }
  return out;
}

// the whole fix-point from the three words (NNw != 0)
__device__ __forceinline__ uint32_t fixpoint_words(uint32_t NNw, uint32_t NRw, uint32_t RNw) {
  const uint32_t any = (NNw | (NNw >> 4) | (NNw >> 8) | (NNw >> 12) | (NNw >> 16) | (NNw >> 20)) & 0xFu;
  const int s = __ffs(any) - 1;
  return fixpoint_loop(NNw, NRw, RNw, s, agents_of_pairs((NNw >> s) & 0x111111u), 0u);
}

// The fix-point proper, given the both-on-course masks NNw != 0 of all near pairs.  The reference walks the sub-steps
// one by one and repeats passes within a sub-step (:247-405); the crashed set only changes at a sub-step with a hit, so
// this loop jumps from hit to hit: `live` = hits still possible under the current crashed set (4 bits per pair), the
// earliest one at or after the current sub-step crashes its agents, and the same sub-step is looked at again (= the
// next pass).  Every round adds a crashed agent, so there are at most four.
__device__ __forceinline__ uint32_t resolve(const SimTab& T, const PairGeom& g, uint32_t effp, uint32_t NNw, uint32_t stop) {
  const uint32_t any = (NNw | (NNw >> 4) | (NNw >> 8) | (NNw >> 12) | (NNw >> 16) | (NNw >> 20)) & 0xFu;
  int s = __ffs(any) - 1;
  uint32_t crashed = agents_of_pairs((NNw >> s) & 0x111111u);
  if (crashed & stop) return crashed << 12;
  uint32_t NRw = 0, RNw = 0;                          // second agent crashed (stationary) / first agent crashed
  {
    int p = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i + 1; j < 4; ++j, ++p)
        if ((g.near6 >> p) & 1u) {
          const uint32_t base = geom_didx(g, p) * (N_EFF * N_EFF);
          NRw |= (uint32_t)T.lut[base + ((effp >> (4 * i)) & 0xFu) * N_EFF] << (4 * p);
          RNw |= (uint32_t)T.lut[base + ((effp >> (4 * j)) & 0xFu)] << (4 * p);
        }
  }
  return fixpoint_loop(NNw, NRw, RNw, s, crashed, stop);
}

__device__ __forceinline__ uint32_t collide(const SimTab& T, const PairGeom& g, uint32_t effp, uint32_t stop = 0) {
  if (g.near6 == 0) return 0;
  const uint32_t NNw = nn_masks(T, g, effp, 0x3Fu);
  if (NNw == 0) return 0;                             // nobody collides while everyone is on course
  return resolve(T, g, effp, NNw, stop);
}

// cell of agent at the end of sub-step s (grid_world.py:259-264 floor index; crashed => start cell)
__device__ __forceinline__ uint32_t cell_at(uint32_t p0, const Traj& t, uint32_t crashed_s, int s) {
  if (crashed_s) return p0;
  if (t.two) return s == 0 ? p0 : (s == 3 ? t.p2 : t.p1);
  return s == 3 ? t.p1 : p0;
}

struct StepResult {
  uint32_t cells;       // 4 x 8-bit final cells
  uint32_t crash;       // bit i
  uint32_t restr;       // bit i
  uint32_t caught;      // 3-bit counters, field (eater*2 + apple)
  uint32_t effs;        // 4 x 4-bit effective trajectories of the actions played
};

// GWorld.UpdateGWorld with explicit actions for all n agents (custom/grid_world.py:424-563).
// OWN_APPLES: count only eater e on apple e (all the env wrappers ever look at, ma_customenv.py:264 / customenv.py:143).
template <bool OWN_APPLES = false>
__device__ __forceinline__ StepResult world_update(const SimTab& T, int n, uint32_t cells, uint32_t acts, const PairGeom& g,
                                                   uint32_t apple_cells, uint32_t apple_on, int n_eaters) {
  Traj t[4];
  StepResult r;
  r.effs = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t[i] = make_traj(T.next, (cells >> (8 * i)) & 0xFFu, i < n ? (acts >> (4 * i)) & 0xFu : 0u);
    r.effs |= t[i].eff << (4 * i);
  }
  const uint32_t cm = (n >= 2) ? collide(T, g, r.effs) : 0u;
  r.crash = (cm >> 12) & 0xFu;
  r.cells = 0;
  r.restr = 0;
  r.caught = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t p0 = (cells >> (8 * i)) & 0xFFu;
    // a lone agent never moves (the floor positions are only assigned inside the pair loops, :259/:264)
    const uint32_t fin = (n >= 2) ? cell_at(p0, t[i], (r.crash >> i) & 1u, 3) : p0;
    r.cells |= fin << (8 * i);
    // the second move is attempted only if the agent did not crash during sub-step 0 (:464)
    if (i < n && (t[i].r1 | (t[i].r2 & ~(cm >> i) & 1u))) r.restr |= 1u << i;
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {                                            // :531-540, every sub-step
    if (e < n_eaters) {
      const uint32_t p0 = (cells >> (8 * e)) & 0xFFu;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint32_t cur = (n >= 2) ? cell_at(p0, t[e], (cm >> (4 * s + e)) & 1u, s) : p0;
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if ((!OWN_APPLES || k == e) && ((apple_on >> k) & 1u) && cur == ((apple_cells >> (8 * k)) & 0xFFu))
            r.caught += 1u << (3 * (e * 2 + k));
      }
    }
  }
  return r;
}

// pairs (01,02,03,12,13,23) that involve agent j
__device__ __forceinline__ uint32_t pairs_of_agent(int j) {
  return j == 0 ? 0x07u : (j == 1 ? 0x19u : (j == 2 ? 0x2Au : 0x34u));
}

// CountValidMovesOfAffected (custom/Responsibility.py:20-54) for one action list: how many of the nine actions of
// the affected agent j leave it neither restricted nor crashed (:46).  eff_others: effective trajectories of
// everybody else, agents outside the list already 0 (Stay, :43).  in_list = false: j is not in the list, its action
// cannot be swapped in, so all nine simulations are the same one (j Stays) and the count is 0 or 9.
//
// No simulation per action.  An unrestricted action a is the effective trajectory a (restricted ones are invalid
// whatever happens), so the candidates are the bits of unres[cell_j].  Until j is hit for the first time the other
// agents evolve exactly as they would without j (j only influences them by colliding, and then it has crashed and the
// action is invalid anyway), so one fix-point over the three pairs that do not involve j gives every other agent k its
// crashed sets C_s after each sub-step, and j is hit at sub-step s iff for some near k
//     k was on course when the sub-step started (k not in C_{s-1}) and lut[d_jk][e_j][eff_k] has bit s, or
//     k stands crashed on its start cell by the end of it (k in C_s) and lut[d_jk][e_j][0] has bit s
// (every change of the crashed set is followed by another pass over all pairs, grid_world.py:247-405).  rowmask4 holds
// those bits for all trajectories of j at once, so the nine answers are a few 64-bit logic operations.
// Trajectories of j (16-bit field per sub-step) that get hit by the neighbour o: `on` = while o is on course, `st` = while o
// stands crashed on its start cell; cm = crashed sets of the other agents after each sub-step (0: nobody ever crashes).
__device__ __forceinline__ unsigned long long hits_on_j(unsigned long long on, unsigned long long st, uint32_t cm, int o) {
  if (cm == 0) return on;
  const uint32_t c = (cm >> o) & 0x1111u, cp = c << 4;                // bit 4s: o in C_s / in C_{s-1} (C_{-1} empty)
  unsigned long long m_on = 0, m_st = 0;
#pragma unroll
  for (int ss = 0; ss < 4; ++ss) {
    if (!((cp >> (4 * ss)) & 1u)) m_on |= 0xFFFFull << (16 * ss);
    if ((c >> (4 * ss)) & 1u) m_st |= 0xFFFFull << (16 * ss);
  }
  return (on & m_on) | (st & m_st);
}

// index of pair (i, k), i < k, in the order 01 02 03 12 13 23
__device__ __forceinline__ int pair_index(int i, int k) { return i == 0 ? k - 1 : (i == 1 ? k + 1 : 5); }

__device__ __forceinline__ uint32_t count_valid_moves(const SimTab& T, uint32_t cells, uint32_t eff_others, const PairGeom& g,
                                                      int j, bool in_list) {
  const uint32_t eo = eff_others;                                      // j's own nibble is never looked at
  const uint32_t cand = in_list ? (uint32_t)T.unres[(cells >> (8 * j)) & 0xFFu] : 1u;
  const unsigned long long dw = (unsigned long long)g.didx_lo | ((unsigned long long)g.didx_hi << 32);
  // the three others among themselves (walked as j+1, j+2, j+3 mod 4, so nothing here depends on a compile-time j)
  uint32_t NN = 0, cm_o = 0;
#pragma unroll
  for (int a = 1; a <= 2; ++a)
#pragma unroll
    for (int b = a + 1; b <= 3; ++b) {
      const int o1 = (j + a) & 3, o2 = (j + b) & 3, i = min(o1, o2), k = max(o1, o2), pp = pair_index(i, k);
      if ((g.near6 >> pp) & 1u)
        NN |= (uint32_t)T.lut[(uint32_t)((dw >> (8 * pp)) & 0xFFu) * (N_EFF * N_EFF) + ((eo >> (4 * i)) & 0xFu) * N_EFF + ((eo >> (4 * k)) & 0xFu)] << (4 * pp);
    }
  if (NN) {                                                            // rare: they do collide
    uint32_t NR = 0, RN = 0;
#pragma unroll
    for (int a = 1; a <= 2; ++a)
#pragma unroll
      for (int b = a + 1; b <= 3; ++b) {
        const int o1 = (j + a) & 3, o2 = (j + b) & 3, i = min(o1, o2), k = max(o1, o2), pp = pair_index(i, k);
        if ((g.near6 >> pp) & 1u) {
          const uint32_t base = (uint32_t)((dw >> (8 * pp)) & 0xFFu) * (N_EFF * N_EFF);
          NR |= (uint32_t)T.lut[base + ((eo >> (4 * i)) & 0xFu) * N_EFF] << (4 * pp);
          RN |= (uint32_t)T.lut[base + ((eo >> (4 * k)) & 0xFu)] << (4 * pp);
        }
      }
    cm_o = fixpoint_words(NN, NR, RN);
  }
  unsigned long long bad = 0;
#pragma unroll
  for (int a = 1; a <= 3; ++a) {
    const int o = (j + a) & 3, pp = pair_index(min(j, o), max(j, o));
    if ((g.near6 >> pp) & 1u) {
      // pair (o, j), j second -> row [d][eff_o]; pair (j, o), j first -> by symmetry row [40 - d][eff_o]
      uint32_t d = (uint32_t)((dw >> (8 * pp)) & 0xFFu);
      if (j < o) d = (uint32_t)(N_DELTA - 1) - d;
      const unsigned long long on = *reinterpret_cast<const unsigned long long*>(T.rowmask4[d * N_EFF + ((eo >> (4 * o)) & 0xFu)]);
      const unsigned long long st = *reinterpret_cast<const unsigned long long*>(T.rowmask4[d * N_EFF]);
      bad |= hits_on_j(on, st, cm_o, o);
    }
  }
  const uint32_t bad16 = (uint32_t)(bad | (bad >> 16) | (bad >> 32) | (bad >> 48)) & 0xFFFFu;
  const uint32_t count = (uint32_t)__popc(cand & ~bad16);
  return in_list ? count : count * 9u;
}

// The same with the six pairs as predicated compile-time bodies: more instructions, fewer dependent index computations.
// Measured equal or ~1 % faster in the thread-per-env step kernel (many resident warps); the form above has the shorter
// dependent chain and is the one the small-batch kernel and the operator kernels use.
__device__ __forceinline__ uint32_t count_valid_moves_p6(const SimTab& T, uint32_t cells, uint32_t eff_others, const PairGeom& g,
                                                      int j, bool in_list) {
  const uint32_t pj = pairs_of_agent(j);
  const uint32_t eo = eff_others & ~(0xFu << (4 * j));
  const uint32_t cand = in_list ? (uint32_t)T.unres[(cells >> (8 * j)) & 0xFFu] : 1u;
  // the others among themselves
  uint32_t cm_o = 0;
  const uint32_t nn_fixed = nn_masks(T, g, eo, 0x3Fu & ~pj);
  if (nn_fixed) {
    PairGeom go = g;
    go.near6 = g.near6 & ~pj;
    cm_o = resolve(T, go, eo, nn_fixed, 0u);
  }
  unsigned long long bad = 0;
  {
    int p = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = i + 1; k < 4; ++k, ++p)
        if ((i == j || k == j) && ((g.near6 >> p) & 1u)) {
          // pair (i, k): j second -> row [d][eff_i]; j first -> by symmetry row [40 - d][eff_k]
          const int o = (k == j) ? i : k;
          const uint32_t d = (k == j) ? geom_didx(g, p) : (uint32_t)(N_DELTA - 1) - geom_didx(g, p);
          const unsigned long long on = *reinterpret_cast<const unsigned long long*>(T.rowmask4[d * N_EFF + ((eo >> (4 * o)) & 0xFu)]);
          if (cm_o == 0) {
            bad |= on;                                                  // nobody else ever crashes
          } else {
            const unsigned long long st = *reinterpret_cast<const unsigned long long*>(T.rowmask4[d * N_EFF]);
            const uint32_t c = (cm_o >> o) & 0x1111u;                   // bit 4s: o in C_s
            const uint32_t cs = c, cp = c << 4;                         // C_s / C_{s-1} (C_{-1} empty)
            unsigned long long m_on = 0, m_st = 0;
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) {
              if (!((cp >> (4 * ss)) & 1u)) m_on |= 0xFFFFull << (16 * ss);
              if ((cs >> (4 * ss)) & 1u) m_st |= 0xFFFFull << (16 * ss);
            }
            bad |= (on & m_on) | (st & m_st);
          }
        }
  }
  const uint32_t bad16 = (uint32_t)(bad | (bad >> 16) | (bad >> 32) | (bad >> 48)) & 0xFFFFu;
  const uint32_t count = (uint32_t)__popc(cand & ~bad16);
  return in_list ? count : count * 9u;
}

// np.sum over the Resp matrix (one non-zero row): numpy's 8-lane pairwise reduction adds the row as
// first + (second + third) for n = 4, first + second for n = 3 (numpy association checked in tests/).
__device__ __forceinline__ double fear_sum3(int n, double r0, double r1, double r2) {
  return n == 4 ? r0 + (r1 + r2) : (n == 3 ? r0 + r1 : r0);
}

}  // namespace gw
