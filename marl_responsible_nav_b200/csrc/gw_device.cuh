// Device-side building blocks of the batched grid world (sm_100a).
//
// Formulation.  The reference grows a Python list of cells per agent and per
// sub-step (custom/grid_world.py:458-518) and re-derives floor/ceil positions
// from it in every collision pass (:255-264).  Here an agent's whole nominal
// trajectory is three packed cells (p0 start, p1 after move 1, p2 after move 2):
// a crash reverts every later path entry to p0 (:200-208), so at any sub-step an
// agent is either on its nominal trajectory or "reverted" (A = B = p0).  The
// fix-point of :247-405 then runs on a 4-bit crashed mask.  A cell is one byte:
// (row << 4) | col  (W = 16), so cell equality / direction equality are integer
// compares.  tests/ diff this against the literal restatement in oracle/.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/gridworld_b200.h"

namespace gw {

struct Tables {                         // device-global, read-only, built by gw_create
  uint16_t map_rows[GW_MAX_H];
  uint8_t mdr_map[GW_MAX_H * GW_W];
  uint8_t policy_map[GW_MAX_H * GW_W];
  uint32_t policy_thr[GW_MAX_POLICIES][2][8];   // 31-bit cdf thresholds, [policy][perturbed][k]
  uint8_t active_cell[GW_MAX_H * GW_W];         // row-major list of active cells (np.where order)
  int32_t n_active;
  int32_t pad_;
  double resp_lut[10][10];                      // clip((m-a)/(m+1e-6),-1,1), Responsibility.py:194-198
};

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// ---------------------------------------------------------------- geometry
__device__ __forceinline__ int manhattan(uint32_t a, uint32_t b) {
  return abs((int)(a >> 4) - (int)(b >> 4)) + abs((int)(a & 15) - (int)(b & 15));
}

__device__ __forceinline__ bool cell_ok(const uint16_t* rows, int H, int r, int c) {
  return (unsigned)r < (unsigned)H && (unsigned)c < (unsigned)GW_W && ((rows[r] >> c) & 1);
}

// Action table custom/custom_agent.py:140-150: 0 Stay, 1 Up, 2 Down, 3 Left, 4 Right, 5-8 the same twice.
__device__ __forceinline__ void action_delta(int a, int& dr, int& dc, int& len) {
  len = a >= 5 ? 2 : 1;
  const int d = (a == 0) ? -1 : ((a - 1) & 3);
  dr = (d == 0) ? -1 : (d == 1) ? 1 : 0;
  dc = (d == 2) ? -1 : (d == 3) ? 1 : 0;
}

// One move of grid_world.py:481-518: off-grid (clip) or inactive target => stay + restricted.
__device__ __forceinline__ uint32_t try_move(const uint16_t* rows, int H, uint32_t p, int dr, int dc, bool& blocked) {
  const int r = (int)(p >> 4) + dr, c = (int)(p & 15) + dc;
  const bool ok = cell_ok(rows, H, r, c);
  blocked = !ok;
  return ok ? (uint32_t)((r << 4) | c) : p;
}

// get_action_mask, custom/ma_customenv.py:467-506: only the TARGET cell is tested.
__device__ __forceinline__ uint32_t action_mask_bits(const uint16_t* rows, int H, uint32_t p) {
  uint32_t m = 1u;
#pragma unroll
  for (int a = 1; a < GW_N_ACTIONS; ++a) {
    int dr, dc, len;
    action_delta(a, dr, dc, len);
    if (cell_ok(rows, H, (int)(p >> 4) + dr * len, (int)(p & 15) + dc * len)) m |= 1u << a;
  }
  return m;
}

// ---------------------------------------------------------------- pair test
// custom/grid_world.py:276-390, first match wins.  q = (step+1)*len in quarter
// sub-steps, f = floor(q/4), c = ceil(q/4); A = path[f], B = path[c].
__device__ __forceinline__ bool pair_hit(uint32_t Ai, uint32_t Bi, uint32_t Pi, int qi, int fi, int ci,
                                         uint32_t Aj, uint32_t Bj, uint32_t Pj, int qj, int fj, int cj) {
  if (Ai == Aj || Bi == Bj) return true;                                   // :276-278
  if (Ai == Bj && Bi == Aj) return true;                                   // :291-294
  const bool same_dir = ((int)Bi - (int)Ai) == ((int)Bj - (int)Aj);        // :314-318 (cell codes: unique per direction)
  if (Ai == Bj) return !(((4 * ci - qi) + (qj - 4 * fj) <= 4) && same_dir);   // :307-326
  if (Bi == Aj) return !(((4 * cj - qj) + (qi - 4 * fi) <= 4) && same_dir);   // :339-357
  return (Ai == Pj && Pi == Aj) || (Bi == Pj && Pi == Bj) ||               // :371-378
         (Ai == Pj && Pi == Bj) || (Bi == Pj && Pi == Aj);
}

struct SimResult {
  uint32_t cells;       // 4 x 8-bit final cells
  uint32_t crash;       // bit i
  uint32_t restr;       // bit i
  uint32_t caught;      // 3-bit counters, field (eater*2 + apple)
  uint32_t unresolved;
};

// GWorld.UpdateGWorld with explicit actions for all n agents (custom/grid_world.py:424-563).
// cells: 4 x 8-bit start cells; acts: 4 x 4-bit action ids.
template <bool APPLES>
__device__ __forceinline__ SimResult simulate(const uint16_t* rows, int H, int n, uint32_t cells, uint32_t acts,
                                              uint32_t apple_cells = 0, uint32_t apple_on = 0, int n_eaters = 0) {
  uint32_t p0[4], p1[4], p2[4];
  int len[4];
  uint32_t r1 = 0, r2 = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    p0[i] = (cells >> (8 * i)) & 0xFFu;
    p1[i] = p2[i] = p0[i];
    len[i] = 1;
    if (i < n) {
      const int a = (acts >> (4 * i)) & 0xF;
      int dr, dc;
      action_delta(a, dr, dc, len[i]);
      bool b1 = false, b2 = false;
      p1[i] = try_move(rows, H, p0[i], dr, dc, b1);
      p2[i] = p1[i];
      if (len[i] == 2) p2[i] = try_move(rows, H, p1[i], dr, dc, b2);
      if (a != 0 && b1) r1 |= 1u << i;
      if (b2) r2 |= 1u << i;
    }
  }
  uint32_t crashed = 0, crashed_after0 = 0, caught = 0, unresolved = 0;
  uint32_t A[4], B[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    int q[4], f[4], c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q[i] = (s + 1) * len[i];
      f[i] = q[i] >> 2;
      c[i] = (q[i] + 3) >> 2;
      const uint32_t pf = f[i] == 0 ? p0[i] : (f[i] == 1 ? p1[i] : p2[i]);
      const uint32_t pc = c[i] == 1 ? p1[i] : p2[i];
      const bool rev = (crashed >> i) & 1;
      A[i] = rev ? p0[i] : pf;
      B[i] = rev ? p0[i] : pc;
    }
    if (n >= 2) {
      int loops = 0;
      while (true) {                                                       // :250-405
        ++loops;
        uint32_t hits = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
          for (int j = i + 1; j < 4; ++j) {
            if (j < n && pair_hit(A[i], B[i], p0[i], q[i], f[i], c[i], A[j], B[j], p0[j], q[j], f[j], c[j]))
              hits |= (1u << i) | (1u << j);
          }
        }
        crashed |= hits;                                                   // record_collision :407-412
#pragma unroll
        for (int i = 0; i < 4; ++i)                                        // revertStepsWithCollisions :190-209
          if ((crashed >> i) & 1) A[i] = B[i] = p0[i];
        if (hits == 0) break;
        if (loops >= 2 * n) { unresolved = 1; break; }                     // 'Collisions Not Resolved' :400-402
      }
    }
    if (s == 0) crashed_after0 = crashed;
    if (APPLES) {                                                          // :531-540, every sub-step
#pragma unroll
      for (int e = 0; e < 2; ++e) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint32_t cur = (n >= 2) ? A[e] : p0[e];
          if (e < n_eaters && ((apple_on >> k) & 1) && cur == ((apple_cells >> (8 * k)) & 0xFFu))
            caught += 1u << (3 * (e * 2 + k));
        }
      }
    }
  }
  SimResult out;
  out.cells = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) out.cells |= ((n >= 2) ? A[i] : p0[i]) << (8 * i);   // :213-231 (a lone agent never moves)
  out.crash = crashed;
  out.restr = r1 | (r2 & ~crashed_after0);          // the 2nd move is attempted only if not crashed in sub-step 0 (:464)
  out.caught = caught;
  out.unresolved = unresolved;
  return out;
}

// ---------------------------------------------------------------- FeAR (one warp)
// FeAR_4_one_actor (custom/Responsibility.py:135-210) for actor x on one warp: lanes 0..26 are
// (affected slot, affected action); two rounds (actor plays MdR / actor plays its action);
// counts by ballot + popc.  `in_list` = agents present in ActionID4Agents (others Stay, :43).
// Returns the per-slot counts packed 4 bits each: bits [4*js .. ] n_mdr, bits [16 + 4*js ..] n_act.
__device__ __forceinline__ uint32_t fear_counts_warp(const uint16_t* rows, int H, int n, uint32_t cells, uint32_t acts,
                                                     uint32_t in_list, int x, int mdr_x, int lane) {
  uint32_t base = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if ((in_list >> k) & 1) base |= ((acts >> (4 * k)) & 0xFu) << (4 * k);
  const int js = lane / 9, ap = lane - js * 9;
  const int j = js + (js >= x ? 1 : 0);
  const bool lane_on = lane < 27 && j < n;
  uint32_t packed = 0;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const uint32_t av = v == 0 ? (uint32_t)mdr_x : ((acts >> (4 * x)) & 0xFu);
    uint32_t a4 = (base & ~(0xFu << (4 * x))) | (av << (4 * x));
    bool valid = false;
    if (lane_on) {
      if ((in_list >> j) & 1) a4 = (a4 & ~(0xFu << (4 * j))) | ((uint32_t)ap << (4 * j));   // SwapActionIDs4Agents
      const SimResult r = simulate<false>(rows, H, n, cells, a4);
      valid = (((r.crash | r.restr) >> j) & 1) == 0;                        // Responsibility.py:46
    }
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
    for (int s = 0; s < 3; ++s) packed |= (uint32_t)__popc(b & (0x1FFu << (9 * s))) << (16 * v + 4 * s);
  }
  return packed;
}

// np.sum over the Resp matrix (one non-zero row): numpy's 8-lane pairwise reduction adds the row as
// first + (second + third) for n = 4, first + second for n = 3 (checked in tests).
__device__ __forceinline__ double fear_sum_from_counts(const Tables* T, int n, uint32_t packed) {
  double r[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int s = 0; s < 3; ++s)
    if (s < n - 1) r[s] = T->resp_lut[(packed >> (4 * s)) & 0xF][(packed >> (16 + 4 * s)) & 0xF];
  return n == 4 ? r[0] + (r[1] + r[2]) : (n == 3 ? r[0] + r[1] : r[0]);
}

}  // namespace gw
