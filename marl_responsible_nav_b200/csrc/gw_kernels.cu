// libgridworld_b200.so -- kernels + C-ABI (include/gridworld_b200.h).  sm_100a only.
//
// v1 layout: ONE WARP PER ENVIRONMENT.
//   state  : uint4 per env {cells 4x8b, meta, rng tick, episode return 2xi16}, 16 B, L2-resident
//   step   : every lane holds the env's scalars (broadcast load); lanes 0..n-1 fetch / draw the
//            agents' actions; FeAR = 27 lanes x 4 rounds of counterfactual sims, ballot+popc counts;
//            the real step is evaluated redundantly by all lanes (no shuffles needed afterwards);
//            observations leave as 128-bit coalesced stores, 32 lanes x 16 B per instruction.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <new>
#include <string>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "gw_device.cuh"

namespace gw {

// ------------------------------------------------------------------ meta word
// bits 0-1 apples left | 2-3 sticky terminations | 4 sticky truncation | 5-6 prev-distance valid
// 7-11 prev distance learner 0 | 12-16 learner 1 | 17-28 steps in episode
constexpr uint32_t M_APPLES = 0x3u, M_TERM_SH = 2, M_TRUNC = 1u << 4, M_PDV_SH = 5, M_PD0_SH = 7, M_PD1_SH = 12,
                   M_STEPS_SH = 17, M_STEPS_MASK = 0xFFFu;

struct StepParams {
  const Tables* tables;
  uint4* state;
  unsigned long long* stats;      // [STAT_SLOTS][8]
  gw_io io;
  long long E;
  long long env_id_base;
  int n, nl, kind, H, fear_radius, max_steps, auto_reset;
  uint32_t apple_cells, apple_init;   // 2 x 8-bit cells, initial apples-left bits
  uint32_t perturb_thr;               // P(perturb) * 2^32
  uint32_t seed_lo, seed_hi;
  double fear_weight;
  const uint8_t* reset_mask;          // gw_reset only
};

constexpr int STAT_SLOTS = 1024;
enum { ST_EPISODES = 0, ST_LEN, ST_CRASH, ST_APPLES, ST_UNRES, ST_FEAR_NZ, ST_RETURN_MILLI, ST_FEAR_BITS };

// ------------------------------------------------------------------ observation render
// value of one cell for learner k.  custom/ma_customenv.py:303-322 (step) / :198-209 (reset),
// custom/customenv.py:161-163 / :341-344 (single env: raw ids, every remaining apple).
__device__ __forceinline__ float obs_value(const uint16_t* rows, int n, int kind, uint32_t cells, uint32_t apples_left,
                                           uint32_t apple_cells, bool fresh, int k, uint32_t cell) {
  float v = ((rows[cell >> 4] >> (cell & 15)) & 1) ? 0.0f : -1.0f;
  int who = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < n && ((cells >> (8 * i)) & 0xFFu) == cell) who = i;           // WorldState[loc] = idx+1 in agent order
  bool apple_here;
  if (kind == GW_ENV_MULTI)
    apple_here = ((apples_left >> k) & 1) && ((apple_cells >> (8 * k)) & 0xFFu) == cell;   // own apple only
  else
    apple_here = (apples_left & 1) && (apple_cells & 0xFFu) == cell;
  if (who >= 0) {
    if (fresh) v = 0.5f;                                                   // AddAgent marker, grid_world.py:140
    else if (apple_here || kind == GW_ENV_SINGLE) v = (float)(who + 1);    // id+9 is not remapped (A.7)
    else v = (who == k) ? 1.0f : 5.0f;
  }
  if (apple_here) v += 9.0f;
  return v;
}

template <int OBS>
__device__ __forceinline__ void render_obs(void* dst_base, long long e, const uint16_t* rows, int H, int n, int nl,
                                           int kind, uint32_t cells, uint32_t apples_left, uint32_t apple_cells,
                                           bool fresh, int lane) {
  const int cells_per_obs = H * GW_W;
  if (OBS == GW_OBS_F32) {
    float4* dst = reinterpret_cast<float4*>(dst_base) + e * (long long)(nl * cells_per_obs / 4);
    const int quads = nl * cells_per_obs / 4;
    for (int q = lane; q < quads; q += 32) {
      const int k = q / (cells_per_obs / 4);
      const uint32_t c0 = (uint32_t)(q - k * (cells_per_obs / 4)) * 4;
      float4 v;
      v.x = obs_value(rows, n, kind, cells, apples_left, apple_cells, fresh, k, c0 + 0);
      v.y = obs_value(rows, n, kind, cells, apples_left, apple_cells, fresh, k, c0 + 1);
      v.z = obs_value(rows, n, kind, cells, apples_left, apple_cells, fresh, k, c0 + 2);
      v.w = obs_value(rows, n, kind, cells, apples_left, apple_cells, fresh, k, c0 + 3);
      __stcs(dst + q, v);                                                  // streaming: never re-read by this kernel
    }
  } else {
    uint4* dst = reinterpret_cast<uint4*>(dst_base) + e * (long long)(nl * cells_per_obs / 8);
    const int octs = nl * cells_per_obs / 8;
    for (int q = lane; q < octs; q += 32) {
      const int k = q / (cells_per_obs / 8);
      const uint32_t c0 = (uint32_t)(q - k * (cells_per_obs / 8)) * 8;
      uint32_t w[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const __nv_bfloat16 lo = __float2bfloat16(obs_value(rows, n, kind, cells, apples_left, apple_cells, fresh, k, c0 + 2 * h));
        const __nv_bfloat16 hi = __float2bfloat16(obs_value(rows, n, kind, cells, apples_left, apple_cells, fresh, k, c0 + 2 * h + 1));
        w[h] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      }
      __stcs(dst + q, make_uint4(w[0], w[1], w[2], w[3]));
    }
  }
}

__device__ __forceinline__ void write_masks(int8_t* dst, long long e, const uint16_t* rows, int H, int nl, uint32_t cells,
                                            int lane) {
  if (dst == nullptr) return;
  if (lane < nl * GW_N_ACTIONS) {
    const int k = lane / GW_N_ACTIONS, a = lane - k * GW_N_ACTIONS;
    const uint32_t m = action_mask_bits(rows, H, (cells >> (8 * k)) & 0xFFu);
    dst[e * (nl * GW_N_ACTIONS) + lane] = (int8_t)((m >> a) & 1);
  }
}

// ------------------------------------------------------------------ spawn
// setup_env, custom/ma_customenv.py:372-380: a sorted n-subset of the active cells (row-major order).
// Replay mode reads the recorded cells; native mode draws them from Philox (uniform over subsets).
__device__ __forceinline__ uint32_t spawn_cells(const StepParams& p, long long e, uint32_t tick, int lane) {
  uint32_t cells = 0;
  if (p.io.spawn != nullptr) {
    for (int i = 0; i < p.n; ++i) {
      const int r = p.io.spawn[(e * p.n + i) * 2 + 0], c = p.io.spawn[(e * p.n + i) * 2 + 1];
      cells |= (uint32_t)(((r & 15) << 4) | (c & 15)) << (8 * i);
    }
    return cells;
  }
  const unsigned long long gid = (unsigned long long)(p.env_id_base + e);
  const uint4 w = philox4x32(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), tick, 0x100u),
                             make_uint2(p.seed_lo, p.seed_hi));
  const uint32_t words[4] = {w.x, w.y, w.z, w.w};
  int chosen[4] = {0, 0, 0, 0};
  const int na = p.tables->n_active;
  for (int k = 0; k < p.n; ++k) {
    int d = (int)__umulhi(words[k], (uint32_t)(na - k));                   // k-th draw among the remaining cells
    int pos = 0;
    for (int t = 0; t < k; ++t)                                            // chosen[] ascending
      if (d >= chosen[t]) { ++d; pos = t + 1; }
    for (int t = k; t > pos; --t) chosen[t] = chosen[t - 1];
    chosen[pos] = d;
  }
  for (int i = 0; i < p.n; ++i) cells |= (uint32_t)p.tables->active_cell[chosen[i]] << (8 * i);
  return cells;
}

__device__ __forceinline__ uint32_t fresh_meta(const StepParams& p, uint32_t cells) {
  uint32_t meta = p.apple_init;
  if (p.kind == GW_ENV_SINGLE) {                                           // customenv.py:349-352: distance known from reset
    meta |= 1u << M_PDV_SH;
    meta |= (uint32_t)manhattan(cells & 0xFFu, p.apple_cells & 0xFFu) << M_PD0_SH;
  }
  return meta;                                                             // multi: prev_distance = None (:212)
}

__device__ __forceinline__ void stat_add(unsigned long long* stats, int slot, int which, unsigned long long v) {
  atomicAdd(&stats[slot * 8 + which], v);
}

// ------------------------------------------------------------------ reset kernel
template <int OBS>
__global__ void __launch_bounds__(128) gw_reset_kernel(StepParams p) {
  __shared__ uint16_t s_rows[GW_MAX_H];
  if (threadIdx.x < GW_MAX_H) s_rows[threadIdx.x] = p.tables->map_rows[threadIdx.x];
  __syncthreads();
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= p.E) return;
  if (p.reset_mask != nullptr && p.reset_mask[e] == 0) return;
  uint4 st = p.state[e];
  const uint32_t cells = spawn_cells(p, e, st.z, lane);
  const uint32_t meta = fresh_meta(p, cells);
  if (p.io.obs) render_obs<OBS>(p.io.obs, e, s_rows, p.H, p.n, p.nl, p.kind, cells, meta & M_APPLES, p.apple_cells, true, lane);
  write_masks(p.io.action_mask, e, s_rows, p.H, p.nl, cells, lane);
  if (lane < p.n * 2 && p.io.positions) {
    const uint32_t c = (cells >> (8 * (lane >> 1))) & 0xFFu;
    p.io.positions[e * p.n * 2 + lane] = (int8_t)((lane & 1) ? (c & 15) : (c >> 4));
  }
  if (lane == 0) p.state[e] = make_uint4(cells, meta, st.z + 1, 0u);
}

// ------------------------------------------------------------------ step kernel
template <bool FEAR, int OBS>
__global__ void __launch_bounds__(128) gw_step_kernel(StepParams p) {
  __shared__ uint16_t s_rows[GW_MAX_H];
  if (threadIdx.x < GW_MAX_H) s_rows[threadIdx.x] = p.tables->map_rows[threadIdx.x];
  __syncthreads();
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= p.E) return;
  const Tables* T = p.tables;
  const int n = p.n, nl = p.nl;

  const uint4 st = p.state[e];                                             // same address on all lanes: one broadcast load
  const uint32_t cells = st.x;
  uint32_t meta = st.y;
  const uint32_t tick = st.z;

  // ---- setup_step (ma_customenv.py:432-452): lane i < n owns agent i's action and MdR
  int my_act = 0, my_mdr = 0;
  if (lane < n) {
    const uint32_t c = (cells >> (8 * lane)) & 0xFFu;
    my_mdr = T->mdr_map[c];
    if (lane < nl) {
      my_act = p.io.learner_actions[e * nl + lane];                        // :239-242
    } else if (p.io.npc_actions != nullptr) {
      my_act = p.io.npc_actions[e * n + lane];
    } else {
      const unsigned long long gid = (unsigned long long)(p.env_id_base + e);
      const uint4 w = philox4x32(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), tick, (uint32_t)lane),
                                 make_uint2(p.seed_lo, p.seed_hi));
      const int pert = w.x < p.perturb_thr ? 1 : 0;                        // random.random() < 0.25 (:441)
      const uint32_t* thr = T->policy_thr[T->policy_map[c]][pert];
      const uint32_t u = w.y >> 1;
      int a = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += (u >= thr[k]) ? 1 : 0;              // np.random.choice(9, p) (custom_agent.py:31)
      my_act = a;
    }
    my_act = min(max(my_act, 0), GW_N_ACTIONS - 1);
  }
  uint32_t acts = 0, mdrs = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acts |= ((uint32_t)__shfl_sync(0xFFFFFFFFu, my_act, i) & 0xFu) << (4 * i);
    mdrs |= ((uint32_t)__shfl_sync(0xFFFFFFFFu, my_mdr, i) & 0xFu) << (4 * i);
  }

  // ---- FeAR on the pre-step positions (ma_customenv.py:245-252 / customenv.py:113-120)
  double fear[GW_MAX_LEARNERS] = {0.0, 0.0};
  if (FEAR) {
    for (int x = 0; x < nl; ++x) {
      uint32_t close = 0;                                                  // close_agents :456-464
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < n && (k == x || manhattan((cells >> (8 * x)) & 0xFFu, (cells >> (8 * k)) & 0xFFu) <= p.fear_radius))
          close |= 1u << k;
      const uint32_t packed = fear_counts_warp(s_rows, p.H, n, cells, acts, close, x, (mdrs >> (4 * x)) & 0xF, lane);
      fear[x] = fear_sum_from_counts(T, n, packed);
    }
  }

  // ---- the real update (:254), evaluated identically on every lane
  const uint32_t apples_before = meta & M_APPLES;
  const SimResult r = simulate<true>(s_rows, p.H, n, cells, acts, p.apple_cells, apples_before, nl);
  const uint32_t cells_new = r.cells;

  double reward[GW_MAX_LEARNERS] = {0.0, 0.0};
  uint32_t apples_left = apples_before;
  uint32_t term_now = 0, trunc_now = 0, apples_rewarded = 0, crash_count = 0, shaped = 0;
  if (p.kind == GW_ENV_MULTI) {
    uint32_t term = (meta >> M_TERM_SH) & 3u, trunc = (meta & M_TRUNC) ? 1u : 0u;
    int ri[GW_MAX_LEARNERS] = {0, 0};
#pragma unroll
    for (int k = 0; k < GW_MAX_LEARNERS; ++k)                              // own apple only (:258-271)
      if (k < nl && ((apples_left >> k) & 1) && ((r.caught >> (3 * (k * 2 + k))) & 7u)) {
        apples_left &= ~(1u << k);
        ri[k] += 20;
        ++apples_rewarded;
      }
    if (apples_rewarded && apples_left == 0) {                             // last apple: +20 to all, truncate (:272-275)
#pragma unroll
      for (int k = 0; k < GW_MAX_LEARNERS; ++k) if (k < nl) ri[k] += 20;
      trunc = 1;
    }
    uint32_t pdv = 0, pd[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < GW_MAX_LEARNERS; ++k) {
      if (k >= nl) continue;
      if ((r.crash >> k) & 1) {                                            // :281-285
        ri[k] -= 10;
        ++crash_count;
        trunc = 1;
        term |= 1u << k;
      }
      if ((apples_left >> k) & 1) {                                        // :287-300
        const uint32_t d = (uint32_t)manhattan((cells_new >> (8 * k)) & 0xFFu, (p.apple_cells >> (8 * k)) & 0xFFu);
        const uint32_t prev_valid = (meta >> (M_PDV_SH + k)) & 1u;
        const uint32_t prev = (meta >> (k == 0 ? M_PD0_SH : M_PD1_SH)) & 31u;
        if (prev_valid && prev > d) { ri[k] += 1; shaped |= 1u << k; }
        pdv |= 1u << k;
        pd[k] = d;
      }
      reward[k] = (double)ri[k];
    }
    term_now = term;
    trunc_now = trunc ? ((1u << nl) - 1u) : 0u;
    const uint32_t steps = min(((meta >> M_STEPS_SH) & M_STEPS_MASK) + 1u, M_STEPS_MASK);
    meta = apples_left | (term << M_TERM_SH) | (trunc ? M_TRUNC : 0u) | (pdv << M_PDV_SH) | (pd[0] << M_PD0_SH) |
           (pd[1] << M_PD1_SH) | (steps << M_STEPS_SH);
  } else {                                                                 // customenv.py:126-158
    double rew = 0.0;
    const uint32_t apple = p.apple_cells & 0xFFu;
    const uint32_t d = (uint32_t)manhattan(cells_new & 0xFFu, apple);
    if (r.crash & 1u) { rew -= 10.0; term_now = 1; crash_count = 1; }
    if ((apples_left & 1u) && (r.caught & 7u) == 1u) {                     // len(apples_caught) == 1 (:143)
      apples_left &= ~1u;
      rew += 20.0;
      trunc_now = 1;
      apples_rewarded = 1;
    }
    const uint32_t prev = (meta >> M_PD0_SH) & 31u;
    if (d < prev) { rew += 0.1; shaped = 1; }                              // :157-158
    reward[0] = rew;
    const uint32_t steps = min(((meta >> M_STEPS_SH) & M_STEPS_MASK) + 1u, M_STEPS_MASK);
    meta = apples_left | (1u << M_PDV_SH) | (d << M_PD0_SH) | (steps << M_STEPS_SH);
  }
  const uint32_t steps_now = (meta >> M_STEPS_SH) & M_STEPS_MASK;
  const bool episode_over = (p.kind == GW_ENV_MULTI) ? (trunc_now != 0) : ((term_now | trunc_now) != 0);
  const bool ended = episode_over || (p.max_steps > 0 && (int)steps_now >= p.max_steps);

  // ---- scalar outputs
  if (lane < nl) {
    const long long o = e * nl + lane;
    const double f = fear[lane == 0 ? 0 : 1], rw = reward[lane == 0 ? 0 : 1];
    if (p.io.reward) p.io.reward[o] = (float)rw;
    if (p.io.fear) p.io.fear[o] = f;
    if (p.io.shaped_reward) p.io.shaped_reward[o] = (float)(p.fear_weight * f + rw);   // maddpg/agent.py:130
    if (p.io.terminated) p.io.terminated[o] = (uint8_t)((term_now >> lane) & 1u);
    if (p.io.truncated) p.io.truncated[o] = (uint8_t)((trunc_now >> lane) & 1u);
  }
  if (lane < n * 2 && p.io.positions) {
    const uint32_t c = (cells_new >> (8 * (lane >> 1))) & 0xFFu;
    p.io.positions[e * n * 2 + lane] = (int8_t)((lane & 1) ? (c & 15) : (c >> 4));
  }
  if (lane == 0) {
    if (p.io.ended) p.io.ended[e] = ended ? 1 : 0;
    if (p.io.info)
      p.io.info[e] = (r.crash & 15u) | ((r.restr & 15u) << 4) | (crash_count << 8) | (apples_rewarded << 10) |
                     ((ended ? 1u : 0u) << 12) | (r.unresolved << 13) | (shaped << 14);
  }

  // ---- episode return (reward units: 1 multi, 0.1 single) and statistics
  int ret0 = (int)(short)(st.w & 0xFFFFu), ret1 = (int)(short)(st.w >> 16);
  const double unit = (p.kind == GW_ENV_MULTI) ? 1.0 : 10.0;
  ret0 += (int)lrint(reward[0] * unit);
  ret1 += (int)lrint(reward[1] * unit);
  if (lane == 0) {
    const int slot = (int)(e & (STAT_SLOTS - 1));
    if (ended) {
      stat_add(p.stats, slot, ST_EPISODES, 1);
      stat_add(p.stats, slot, ST_LEN, steps_now);
      stat_add(p.stats, slot, ST_RETURN_MILLI, (unsigned long long)(long long)llrint((ret0 + ret1) * (1000.0 / unit)));
    }
    if (crash_count) stat_add(p.stats, slot, ST_CRASH, crash_count);
    if (apples_rewarded) stat_add(p.stats, slot, ST_APPLES, apples_rewarded);
    if (r.unresolved) stat_add(p.stats, slot, ST_UNRES, 1);
    if (FEAR) {
      const int nz = (fear[0] != 0.0) + (fear[1] != 0.0);
      if (nz) {
        stat_add(p.stats, slot, ST_FEAR_NZ, nz);
        atomicAdd(reinterpret_cast<double*>(&p.stats[slot * 8 + ST_FEAR_BITS]), fear[0] + fear[1]);
      }
    }
  }

  // ---- observations, masks, auto-reset
  uint32_t cells_out = cells_new, meta_out = meta, tick_out = tick + 1, ret_out;
  if (ended && p.auto_reset) {
    if (p.io.final_obs)
      render_obs<OBS>(p.io.final_obs, e, s_rows, p.H, n, nl, p.kind, cells_new, apples_left, p.apple_cells, false, lane);
    cells_out = spawn_cells(p, e, tick, lane);
    meta_out = fresh_meta(p, cells_out);
    ret_out = 0;
    if (p.io.obs)
      render_obs<OBS>(p.io.obs, e, s_rows, p.H, n, nl, p.kind, cells_out, meta_out & M_APPLES, p.apple_cells, true, lane);
  } else {
    ret_out = ((uint32_t)ret0 & 0xFFFFu) | ((uint32_t)ret1 << 16);
    if (p.io.obs)
      render_obs<OBS>(p.io.obs, e, s_rows, p.H, n, nl, p.kind, cells_new, apples_left, p.apple_cells, false, lane);
  }
  write_masks(p.io.action_mask, e, s_rows, p.H, nl, cells_out, lane);
  if (lane == 0) p.state[e] = make_uint4(cells_out, meta_out, tick_out, ret_out);
}

// ------------------------------------------------------------------ operator-level kernels
__global__ void __launch_bounds__(128) gw_update_world_kernel(const Tables* T, int H, int n_default, long long C,
                                                              const int8_t* n_per, const int8_t* pos, const int8_t* act,
                                                              const int8_t* apples, int8_t* new_pos, uint8_t* crash,
                                                              uint8_t* restr, int8_t* caught) {
  __shared__ uint16_t s_rows[GW_MAX_H];
  if (threadIdx.x < GW_MAX_H) s_rows[threadIdx.x] = T->map_rows[threadIdx.x];
  __syncthreads();
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int n = n_per ? n_per[c] : n_default;
  uint32_t cells = 0, acts = 0, apple_cells = 0, apple_on = 0;
  for (int i = 0; i < n; ++i) {
    cells |= (uint32_t)(((pos[(c * 4 + i) * 2] & 15) << 4) | (pos[(c * 4 + i) * 2 + 1] & 15)) << (8 * i);
    acts |= ((uint32_t)act[c * 4 + i] & 0xFu) << (4 * i);
  }
  if (apples)
    for (int k = 0; k < 2; ++k)
      if (apples[(c * 2 + k) * 2] >= 0) {
        apple_on |= 1u << k;
        apple_cells |= (uint32_t)(((apples[(c * 2 + k) * 2] & 15) << 4) | (apples[(c * 2 + k) * 2 + 1] & 15)) << (8 * k);
      }
  const SimResult r = simulate<true>(s_rows, H, n, cells, acts, apple_cells, apple_on, min(2, n));
  for (int i = 0; i < 4; ++i) {
    const uint32_t cc = (r.cells >> (8 * i)) & 0xFFu;
    new_pos[(c * 4 + i) * 2] = i < n ? (int8_t)(cc >> 4) : (int8_t)-1;
    new_pos[(c * 4 + i) * 2 + 1] = i < n ? (int8_t)(cc & 15) : (int8_t)-1;
    crash[c * 4 + i] = i < n ? (uint8_t)((r.crash >> i) & 1) : 0;
    restr[c * 4 + i] = i < n ? (uint8_t)((r.restr >> i) & 1) : 0;
  }
  if (caught)
    for (int f = 0; f < 4; ++f) caught[c * 4 + f] = (int8_t)((r.caught >> (3 * f)) & 7u);
}

__global__ void __launch_bounds__(128) gw_fear_kernel(const Tables* T, int H, int n_default, long long C,
                                                      const int8_t* n_per, const int8_t* pos, const int8_t* act,
                                                      const int8_t* mdr, const int8_t* actor, const uint8_t* in_list,
                                                      double* resp, int8_t* n_mdr, int8_t* n_act) {
  __shared__ uint16_t s_rows[GW_MAX_H];
  if (threadIdx.x < GW_MAX_H) s_rows[threadIdx.x] = T->map_rows[threadIdx.x];
  __syncthreads();
  const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const int n = n_per ? n_per[c] : n_default;
  const int x = actor[c];
  uint32_t cells = 0, acts = 0, lst = 1u << x;
  for (int i = 0; i < n; ++i) {
    cells |= (uint32_t)(((pos[(c * 4 + i) * 2] & 15) << 4) | (pos[(c * 4 + i) * 2 + 1] & 15)) << (8 * i);
    acts |= ((uint32_t)act[c * 4 + i] & 0xFu) << (4 * i);
    if (in_list == nullptr || in_list[c * 4 + i]) lst |= 1u << i;
  }
  const uint32_t packed = fear_counts_warp(s_rows, H, n, cells, acts, lst, x, mdr[c * 4 + x], lane);
  if (lane < 4) {
    double rv = 0.0;
    int8_t m = 0, a = 0;
    if (lane < n && lane != x) {
      const int js = lane - (lane > x ? 1 : 0);
      m = (int8_t)((packed >> (4 * js)) & 0xF);
      a = (int8_t)((packed >> (16 + 4 * js)) & 0xF);
      rv = T->resp_lut[m][a];
    }
    resp[c * 4 + lane] = rv;
    if (n_mdr) n_mdr[c * 4 + lane] = m;
    if (n_act) n_act[c * 4 + lane] = a;
  }
}

}  // namespace gw

// ====================================================================== host side / C-ABI
struct gw_handle {
  gw_config cfg;
  gw::Tables* d_tables = nullptr;
  uint4* d_state = nullptr;
  unsigned long long* d_stats = nullptr;
  bool reset_done = false;
  uint64_t launches = 0;
  uint64_t env_steps = 0;
  std::string err;
};

static thread_local std::string g_create_err;

static int fail(gw_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}
static int cuda_fail(gw_handle* h, cudaError_t e, const char* what) {
  return fail(h, GW_ECUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define GW_CUDA(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(h, e_, #call); } while (0)

extern "C" {

int gw_abi_version(void) { return GW_ABI_VERSION; }

const char* gw_build_info(void) {
  return "gridworld_b200 abi " "1" " sm_100a nvcc "
#ifdef __CUDACC_VER_MAJOR__
#define GW_STR2(x) #x
#define GW_STR(x) GW_STR2(x)
      GW_STR(__CUDACC_VER_MAJOR__) "." GW_STR(__CUDACC_VER_MINOR__)
#endif
      ;
}

const char* gw_last_error(const gw_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int gw_default_config(gw_config* cfg) {
  if (!cfg) return GW_EINVAL;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->struct_size = (int32_t)sizeof(gw_config);
  cfg->abi_version = GW_ABI_VERSION;
  cfg->height = 10;
  cfg->width = GW_W;
  cfg->n_agents = 4;
  cfg->n_learners = 2;
  cfg->env_kind = GW_ENV_MULTI;
  cfg->apple_row[0] = 9; cfg->apple_col[0] = 0;      // ma_customenv.py:422
  cfg->apple_row[1] = 5; cfg->apple_col[1] = 10;
  cfg->perturb_prob = 0.25;
  cfg->fear = 1;
  cfg->fear_radius = 5;
  cfg->fear_weight = 0.0;
  cfg->max_steps = 150;
  cfg->auto_reset = 0;
  cfg->obs_dtype = GW_OBS_F32;
  cfg->num_envs = 1;
  cfg->n_policies = 1;
  cfg->step_weights[0][0] = cfg->step_weights[0][1] = cfg->step_weights[0][2] = 1.f;
  for (int d = 0; d < 4; ++d) cfg->dir_weights[0][d] = 1.f;
  return GW_OK;
}

static int validate(const gw_config* c, std::string& why) {
  if (c->struct_size != (int32_t)sizeof(gw_config)) { why = "gw_config.struct_size mismatch"; return GW_EINVAL; }
  if (c->abi_version != GW_ABI_VERSION) { why = "gw_config.abi_version mismatch"; return GW_EINVAL; }
  if (c->width != GW_W || c->height < 1 || c->height > GW_MAX_H) { why = "grid must be H<=16 x W==16"; return GW_EINVAL; }
  if ((c->height * c->width) % 8 != 0) { why = "H*W must be a multiple of 8"; return GW_EINVAL; }
  if (c->n_agents < 2 || c->n_agents > GW_MAX_AGENTS) { why = "n_agents must be 2..4"; return GW_EINVAL; }
  if (c->n_learners < 1 || c->n_learners > GW_MAX_LEARNERS || c->n_learners > c->n_agents) { why = "n_learners must be 1..2"; return GW_EINVAL; }
  if (c->env_kind != GW_ENV_MULTI && c->env_kind != GW_ENV_SINGLE) { why = "env_kind"; return GW_EINVAL; }
  if (c->env_kind == GW_ENV_SINGLE && c->n_learners != 1) { why = "single env has exactly one learner"; return GW_EINVAL; }
  if (c->obs_dtype != GW_OBS_F32 && c->obs_dtype != GW_OBS_BF16) { why = "obs_dtype"; return GW_EINVAL; }
  if (c->num_envs < 1) { why = "num_envs must be >= 1"; return GW_EINVAL; }
  if (c->n_policies < 1 || c->n_policies > GW_MAX_POLICIES) { why = "n_policies must be 1..16"; return GW_EINVAL; }
  if (c->fear_radius < 0 || c->max_steps < 0 || c->max_steps > 4000) { why = "fear_radius/max_steps out of range"; return GW_EINVAL; }
  if (!(c->perturb_prob >= 0.0 && c->perturb_prob <= 1.0)) { why = "perturb_prob"; return GW_EINVAL; }
  int active = 0;
  for (int r = 0; r < c->height; ++r) active += __builtin_popcount(c->map_rows[r]);
  if (active < c->n_agents) { why = "fewer active cells than agents"; return GW_EINVAL; }
  for (int k = 0; k < c->n_learners; ++k) {
    const int r = c->apple_row[k], col = c->apple_col[k];
    if (r < 0) continue;
    if (r >= c->height || col < 0 || col >= c->width) { why = "apple outside the grid"; return GW_EINVAL; }
  }
  if (c->env_kind == GW_ENV_SINGLE && c->apple_row[0] < 0) { why = "single env needs an apple"; return GW_EINVAL; }
  for (int i = 0; i < c->height * GW_W; ++i) {
    if (c->policy_map[i] >= c->n_policies) { why = "policy_map entry >= n_policies"; return GW_EINVAL; }
    if (c->mdr_map[i] >= GW_N_ACTIONS) { why = "mdr_map entry is not an action id"; return GW_EINVAL; }
  }
  for (int p = 0; p < c->n_policies; ++p) {
    double tot_b = c->step_weights[p][0], tot_p = c->step_weights[p][0], dsum = 0;
    for (int d = 0; d < 4; ++d) { if (c->dir_weights[p][d] < 0) { why = "negative weight"; return GW_EINVAL; } dsum += c->dir_weights[p][d]; }
    for (int s = 0; s < 3; ++s) if (c->step_weights[p][s] < 0) { why = "negative weight"; return GW_EINVAL; }
    tot_b += (c->step_weights[p][1] + c->step_weights[p][2]) * dsum;
    tot_p += (c->step_weights[p][1] + c->step_weights[p][2]) * 4.0;
    if (!(tot_b > 0) || !(tot_p > 0)) { why = "policy weights sum to zero"; return GW_EINVAL; }
  }
  return GW_OK;
}

// GeneratePolicy (custom/custom_agent.py:181-197) -> 31-bit cdf thresholds for the device sampler.
static void policy_thresholds(const float sw[3], const float dw_in[4], bool perturbed, uint32_t thr[8]) {
  double dw[4];
  for (int d = 0; d < 4; ++d) dw[d] = perturbed ? 1.0 : (double)dw_in[d];   // random.shuffle -> None -> [1,1,1,1]
  double p[9];
  p[0] = sw[0];
  for (int d = 0; d < 4; ++d) { p[1 + d] = (double)sw[1] * dw[d]; p[5 + d] = (double)sw[2] * dw[d]; }
  double tot = 0;
  for (int k = 0; k < 9; ++k) tot += p[k];
  double cdf = 0;
  for (int k = 0; k < 8; ++k) {
    cdf += p[k] / tot;
    double t = std::floor(cdf * 2147483648.0 + 0.5);
    if (t > 2147483648.0) t = 2147483648.0;
    if (t < 0) t = 0;
    thr[k] = (uint32_t)t;
  }
  // every action after the last one with p > 0 must be unreachable
  int last = 8;
  while (last > 0 && p[last] == 0.0) --last;
  for (int k = last; k < 8; ++k) thr[k] = 0x80000000u;
}

int gw_create(const gw_config* cfg, gw_handle** out) {
  if (!cfg || !out) return fail(nullptr, GW_EINVAL, "gw_create: null argument");
  *out = nullptr;
  std::string why;
  if (int rc = validate(cfg, why)) return fail(nullptr, rc, "gw_create: " + why);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, GW_ENODEV, std::string("gw_create: no CUDA device (") + cudaGetErrorString(e) +
                                        "); this library has no CPU path");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, GW_EINVAL, "gw_create: bad device ordinal");
  GW_CUDA(nullptr, cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  GW_CUDA(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(nullptr, GW_ENODEV, "gw_create: built for sm_100a (B200) only, found compute capability " +
                                        std::to_string(prop.major) + "." + std::to_string(prop.minor));
  gw_handle* h = new (std::nothrow) gw_handle();
  if (!h) return fail(nullptr, GW_ENOMEM, "gw_create: host allocation failed");
  h->cfg = *cfg;

  gw::Tables* t = new gw::Tables();
  std::memset(t, 0, sizeof(*t));
  std::memcpy(t->map_rows, cfg->map_rows, sizeof(t->map_rows));
  std::memcpy(t->mdr_map, cfg->mdr_map, sizeof(t->mdr_map));
  std::memcpy(t->policy_map, cfg->policy_map, sizeof(t->policy_map));
  for (int p = 0; p < cfg->n_policies; ++p) {
    policy_thresholds(cfg->step_weights[p], cfg->dir_weights[p], false, t->policy_thr[p][0]);
    policy_thresholds(cfg->step_weights[p], cfg->dir_weights[p], true, t->policy_thr[p][1]);
  }
  int na = 0;
  for (int r = 0; r < cfg->height; ++r)
    for (int c = 0; c < cfg->width; ++c)
      if ((cfg->map_rows[r] >> c) & 1) t->active_cell[na++] = (uint8_t)((r << 4) | c);
  t->n_active = na;
  for (int m = 0; m < 10; ++m)
    for (int a = 0; a < 10; ++a) {
      volatile double v = ((double)m - (double)a) / ((double)m + 0.000001);   // Responsibility.py:194-195, EPS :12
      double cl = v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v);                      // np.clip :198
      t->resp_lut[m][a] = cl;
    }
  auto cleanup = [&](int rc) { delete t; gw_destroy(h); return rc; };
  if ((e = cudaMalloc(&h->d_tables, sizeof(gw::Tables))) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMalloc tables"));
  if ((e = cudaMalloc(&h->d_state, sizeof(uint4) * (size_t)cfg->num_envs)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMalloc state"));
  if ((e = cudaMalloc(&h->d_stats, sizeof(unsigned long long) * gw::STAT_SLOTS * 8)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMalloc stats"));
  if ((e = cudaMemcpy(h->d_tables, t, sizeof(gw::Tables), cudaMemcpyHostToDevice)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMemcpy tables"));
  if ((e = cudaMemset(h->d_state, 0, sizeof(uint4) * (size_t)cfg->num_envs)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMemset state"));
  if ((e = cudaMemset(h->d_stats, 0, sizeof(unsigned long long) * gw::STAT_SLOTS * 8)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMemset stats"));
  delete t;
  *out = h;
  return GW_OK;
}

int gw_destroy(gw_handle* h) {
  if (!h) return GW_OK;
  cudaSetDevice(h->cfg.device);
  if (h->d_tables) cudaFree(h->d_tables);
  if (h->d_state) cudaFree(h->d_state);
  if (h->d_stats) cudaFree(h->d_stats);
  delete h;
  return GW_OK;
}

static gw::StepParams make_params(gw_handle* h, const gw_io* io) {
  gw::StepParams p;
  std::memset(&p, 0, sizeof(p));
  const gw_config& c = h->cfg;
  p.tables = h->d_tables;
  p.state = h->d_state;
  p.stats = h->d_stats;
  if (io) p.io = *io;
  p.E = c.num_envs;
  p.env_id_base = c.env_id_base;
  p.n = c.n_agents; p.nl = c.n_learners; p.kind = c.env_kind; p.H = c.height;
  p.fear_radius = c.fear_radius; p.max_steps = c.max_steps; p.auto_reset = c.auto_reset;
  for (int k = 0; k < c.n_learners; ++k)
    if (c.apple_row[k] >= 0) {
      p.apple_cells |= (uint32_t)((c.apple_row[k] << 4) | c.apple_col[k]) << (8 * k);
      p.apple_init |= 1u << k;
    }
  double thr = c.perturb_prob * 4294967296.0;
  p.perturb_thr = thr >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)thr;
  p.seed_lo = (uint32_t)c.seed; p.seed_hi = (uint32_t)(c.seed >> 32);
  p.fear_weight = c.fear_weight;
  return p;
}

static bool misaligned(const void* p, size_t a) { return p && (reinterpret_cast<uintptr_t>(p) % a) != 0; }

static int check_io(gw_handle* h, const gw_io* io, bool step) {
  if (!io) return fail(h, GW_EINVAL, "null gw_io");
  if (step && !io->learner_actions) return fail(h, GW_EINVAL, "gw_step: learner_actions is required");
  if (misaligned(io->obs, 16) || misaligned(io->final_obs, 16)) return fail(h, GW_EINVAL, "obs/final_obs must be 16-byte aligned");
  if (misaligned(io->reward, 4) || misaligned(io->shaped_reward, 4) || misaligned(io->fear, 8) || misaligned(io->info, 4))
    return fail(h, GW_EINVAL, "misaligned output pointer");
  return GW_OK;
}

int gw_reset(gw_handle* h, const uint8_t* reset_mask, const gw_io* io, void* stream) {
  if (!h) return GW_EINVAL;
  if (int rc = check_io(h, io, false)) return rc;
  if (!h->reset_done && reset_mask) return fail(h, GW_ESTATE, "gw_reset: the first reset must cover all envs (reset_mask = NULL)");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  gw::StepParams p = make_params(h, io);
  p.reset_mask = reset_mask;
  const int threads = 128;
  const long long blocks = (h->cfg.num_envs * 32 + threads - 1) / threads;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->cfg.obs_dtype == GW_OBS_F32) gw::gw_reset_kernel<GW_OBS_F32><<<(unsigned)blocks, threads, 0, s>>>(p);
  else gw::gw_reset_kernel<GW_OBS_BF16><<<(unsigned)blocks, threads, 0, s>>>(p);
  GW_CUDA(h, cudaGetLastError());
  h->reset_done = true;
  h->launches += 1;
  return GW_OK;
}

int gw_step(gw_handle* h, const gw_io* io, void* stream) {
  if (!h) return GW_EINVAL;
  if (int rc = check_io(h, io, true)) return rc;
  if (!h->reset_done) return fail(h, GW_ESTATE, "gw_step: call gw_reset first");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  gw::StepParams p = make_params(h, io);
  const int threads = 128;
  const long long blocks = (h->cfg.num_envs * 32 + threads - 1) / threads;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool f32 = h->cfg.obs_dtype == GW_OBS_F32;
  if (h->cfg.fear) {
    if (f32) gw::gw_step_kernel<true, GW_OBS_F32><<<(unsigned)blocks, threads, 0, s>>>(p);
    else gw::gw_step_kernel<true, GW_OBS_BF16><<<(unsigned)blocks, threads, 0, s>>>(p);
  } else {
    if (f32) gw::gw_step_kernel<false, GW_OBS_F32><<<(unsigned)blocks, threads, 0, s>>>(p);
    else gw::gw_step_kernel<false, GW_OBS_BF16><<<(unsigned)blocks, threads, 0, s>>>(p);
  }
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  h->env_steps += (uint64_t)h->cfg.num_envs;
  return GW_OK;
}

int gw_sync(gw_handle* h, void* stream) {
  if (!h) return GW_EINVAL;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  GW_CUDA(h, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  GW_CUDA(h, cudaGetLastError());
  return GW_OK;
}

size_t gw_state_bytes(const gw_handle* h) { return h ? sizeof(uint4) * (size_t)h->cfg.num_envs : 0; }

int gw_get_state(gw_handle* h, void* dst, int dst_is_device, void* stream) {
  if (!h || !dst) return fail(h, GW_EINVAL, "gw_get_state: null argument");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GW_CUDA(h, cudaMemcpyAsync(dst, h->d_state, gw_state_bytes(h), dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  if (!dst_is_device) GW_CUDA(h, cudaStreamSynchronize(s));
  return GW_OK;
}

int gw_set_state(gw_handle* h, const void* src, int src_is_device, void* stream) {
  if (!h || !src) return fail(h, GW_EINVAL, "gw_set_state: null argument");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GW_CUDA(h, cudaMemcpyAsync(h->d_state, src, gw_state_bytes(h), src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  if (!src_is_device) GW_CUDA(h, cudaStreamSynchronize(s));
  h->reset_done = true;
  return GW_OK;
}

int gw_get_stats(gw_handle* h, gw_stats* out, void* stream) {
  if (!h || !out) return fail(h, GW_EINVAL, "gw_get_stats: null argument");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static_assert(sizeof(unsigned long long) == 8, "");
  unsigned long long* host = (unsigned long long*)std::malloc(sizeof(unsigned long long) * gw::STAT_SLOTS * 8);
  if (!host) return fail(h, GW_ENOMEM, "gw_get_stats: host allocation failed");
  cudaError_t e = cudaMemcpyAsync(host, h->d_stats, sizeof(unsigned long long) * gw::STAT_SLOTS * 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { std::free(host); return cuda_fail(h, e, "gw_get_stats copy"); }
  std::memset(out, 0, sizeof(*out));
  long long ret_milli = 0;
  double fear_sum = 0;
  for (int sl = 0; sl < gw::STAT_SLOTS; ++sl) {
    const unsigned long long* r = host + sl * 8;
    out->episodes += r[gw::ST_EPISODES];
    out->episode_len_sum += r[gw::ST_LEN];
    out->crashes += r[gw::ST_CRASH];
    out->apples += r[gw::ST_APPLES];
    out->unresolved += r[gw::ST_UNRES];
    out->fear_nonzero += r[gw::ST_FEAR_NZ];
    ret_milli += (long long)r[gw::ST_RETURN_MILLI];
    double f;
    std::memcpy(&f, &r[gw::ST_FEAR_BITS], 8);
    fear_sum += f;
  }
  std::free(host);
  out->env_steps = h->env_steps;
  out->agent_steps = h->env_steps * (uint64_t)h->cfg.n_learners;
  out->return_sum = (double)ret_milli / 1000.0;
  out->fear_sum = fear_sum;
  return GW_OK;
}

int gw_reset_stats(gw_handle* h, void* stream) {
  if (!h) return GW_EINVAL;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  GW_CUDA(h, cudaMemsetAsync(h->d_stats, 0, sizeof(unsigned long long) * gw::STAT_SLOTS * 8, static_cast<cudaStream_t>(stream)));
  h->env_steps = 0;
  return GW_OK;
}

int gw_launch_count(const gw_handle* h, uint64_t* n) {
  if (!h || !n) return GW_EINVAL;
  *n = h->launches;
  return GW_OK;
}

int gw_update_world(gw_handle* h, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
                    const int8_t* apples, int8_t* new_positions, uint8_t* crash, uint8_t* restricted, int8_t* caught,
                    void* stream) {
  if (!h) return GW_EINVAL;
  if (n_cases < 0 || !positions || !actions || !new_positions || !crash || !restricted)
    return fail(h, GW_EINVAL, "gw_update_world: null/invalid argument");
  if (n_cases == 0) return GW_OK;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  const int threads = 128;
  const long long blocks = (n_cases + threads - 1) / threads;
  gw::gw_update_world_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d_tables, h->cfg.height, h->cfg.n_agents, n_cases, n_per, positions, actions, apples, new_positions, crash,
      restricted, caught);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_fear_one_actor(gw_handle* h, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
                      const int8_t* mdr, const int8_t* actor, const uint8_t* in_list, double* resp, int8_t* n_mdr,
                      int8_t* n_act, void* stream) {
  if (!h) return GW_EINVAL;
  if (n_cases < 0 || !positions || !actions || !mdr || !actor || !resp)
    return fail(h, GW_EINVAL, "gw_fear_one_actor: null/invalid argument");
  if (n_cases == 0) return GW_OK;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  const int threads = 128;
  const long long blocks = (n_cases * 32 + threads - 1) / threads;
  gw::gw_fear_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d_tables, h->cfg.height, h->cfg.n_agents, n_cases, n_per, positions, actions, mdr, actor, in_list, resp, n_mdr,
      n_act);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

}  // extern "C"
