// libgridworld_b200.so -- kernels + C-ABI (include/gridworld_b200.h).  sm_100a only.
//
// Three step kernels share the device code of gw_device.cuh (pair-mask table, fix-point, bit-parallel FeAR counting);
// gw_step picks one by batch size (pick_small / pick_tile, thresholds measured on B200, profiles/README.md):
//   gw_step_small_kernel   <= 6144 envs: 8 lanes per env, 32 envs per 256-thread CTA, no CTA-wide phase -- every warp steps
//                          its four envs, patches their observation rows in shared memory and issues one TMA bulk store
//   gw_step_kernel<TILE>   thread per env, persistent CTAs walking tiles of 32 / 128 / 256 envs in four phases:
//       P1  thread per env          : 16-byte packed state (coalesced), actions / Philox, world update via the pair-mask
//                                     table, rewards / done flags / auto-reset, scalar outputs, FeAR task queue
//       P2  thread per (task, variant): count_valid_moves (no loop over actions), counts into 4-bit fields of a shared word
//       P3  thread per env          : counts -> Resp LUT -> fear, shaped reward, statistics (warp-reduced atomics), masks
//       P4  warp per env            : observation row = staging copy of the constant template + <= 10 patched cells,
//                                     leaves as ONE TMA bulk store of whole 128-byte lines (double-buffered rows)
//     tiles after a CTA's first come from a global counter when there are >= 4 per CTA (self-resetting)
//   gw_step_server_kernel  the small kernel's body inside a RESIDENT kernel for the host-driven step (gw_step_host mode 2):
//                          doorbell / completion word in pinned host memory, no launch or stream sync per step
// Prologue of all of them: one thread brings the 15.6 KB table block and the pre-replicated staging rows in by TMA
// (cp.async.bulk + mbarrier).  State is 16 B/env and lives in HBM/L2 between launches; nothing else is kept by the library.
// DESIGN.md section 5 has the reasoning and the measurements behind each choice.
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <new>
#include <string>

#include <cuda_bf16.h>
#include <type_traits>
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
  const uint8_t* stage_init;      // GW_STAGE_ROWS replicated observation rows (template), see load_tables
  uint4* rng_cache;               // [E][2] small-batch kernel: next step's Philox words + re-spawn draw, tagged with their tick
  unsigned int* tile_ctr;         // [0] tiles handed out beyond the first wave, [1] CTAs that finished (self-resetting)
  uint4* state;
  unsigned long long* stats;      // [STAT_SLOTS][8]
  gw_io io;
  long long E;
  long long env_id_base;
  int n, nl, kind, H, fear_radius, max_steps, auto_reset, n_active;
  uint32_t apple_cells, apple_init;   // 2 x 8-bit cells, initial apples-left bits
  uint32_t perturb_thr;               // P(perturb) * 2^32
  uint32_t seed_lo, seed_hi;
  double fear_weight;
  const uint8_t* reset_mask;          // gw_reset only
  int dyn_tiles;                      // tiles after the first from a global counter (GW_DYN=0: fixed stride)
  int pdl_early;                      // small grids: let the next step's grid start its prologue right away
  unsigned long long* trace;          // GW_TRACE (dev): per CTA 16 globaltimer stamps at the phase boundaries
};

constexpr int STAGE_ROWS = 32;         // rows in the replicated block: 8 warps x up to 4 staging rows
constexpr int STAT_SLOTS = 1024;
enum { ST_EPISODES = 0, ST_LEN, ST_CRASH, ST_APPLES, ST_TASKS, ST_FEAR_NZ, ST_RETURN_MILLI, ST_FEAR_BITS };   // the 2N-pass cap never binds: no 'unresolved' counter

// render record per env (shared memory): what P4 needs
constexpr uint32_t R_FRESH = 1u << 2, R_FINAL = 1u << 3, R_SKIP = 1u << 6;   // bits 0-1 apples shown in obs, 4-5 apples at final

constexpr int N_SPEC = GW_MAX_LEARNERS * (GW_MAX_AGENTS + 1);   // per learner: every agent + the own apple

template <int TILE>
struct Smem {
  SimTab sim;
  alignas(16) SmallTables small;
  alignas(8) unsigned long long bar;                        // mbarrier of the prologue's TMA copies
  union {                                                   // P1b/P2 scratch of the FeAR tasks, then the mask staging (P3/P4)
    struct { uint32_t cells_old[TILE], effs[TILE], geom_lo[TILE], geom_hi[TILE], close[TILE]; };
    alignas(16) uint8_t mask[TILE * GW_MAX_LEARNERS * GW_N_ACTIONS + 16];
  };
  uint32_t cnt[TILE * 2];
  uint32_t cells_fin[TILE], rinfo[TILE];
  uint16_t spec[TILE][N_SPEC];                          // special cells of the tile's observations (offset | value*2 << 9)
  uint16_t queue[TILE * 6];
  uint32_t qn;
  long long next_tile;                                      // dynamic tile scheduling: what thread 0 fetched for the next round
};
// Behind the struct (dynamic shared memory): per warp two observation staging rows of n_learners*H*W elements each.
template <int TILE>
__host__ __device__ constexpr size_t smem_fixed_bytes() { return (sizeof(Smem<TILE>) + 15) / 16 * 16; }

// ---- prologue: tables and staging rows arrive by TMA (cp.async.bulk global -> shared, completion on an mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "GW_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra GW_WAIT_DONE;\n\t"
      "bra GW_WAIT_LOOP;\n\t"
      "GW_WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One thread issues two bulk copies: the lookup tables (sim + small, 15.6 KB) and `stage_rows` pre-replicated copies of
// the constant observation row (-1 inactive / 0 active, grid_world.py:433-434) for the warps' staging rows.  They
// overlap the first tile's state loads and RNG (and, under programmatic dependent launch, the previous step); everybody
// waits on the mbarrier (tables_wait) before the first table lookup.  The replicated rows live in a per-handle global
// block, so the CTAs stream distinct L2 lines instead of all hammering the five lines of a single template row.
template <int TILE>
__device__ __forceinline__ void load_tables(Smem<TILE>& s, uint8_t* stage, const StepParams& p, int stage_rows, int row_bytes) {
  const uint32_t bar = smem_u32(&s.bar);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    constexpr uint32_t tab_bytes = (uint32_t)(offsetof(Smem<TILE>, small) - offsetof(Smem<TILE>, sim) + sizeof(SmallTables));
    static_assert(tab_bytes % 16 == 0, "bulk copies move multiples of 16 bytes");
    static_assert(offsetof(Smem<TILE>, small) - offsetof(Smem<TILE>, sim) == offsetof(Tables, small) - offsetof(Tables, sim), "layout");
    const uint32_t st_bytes = (uint32_t)(stage_rows * row_bytes);
    mbar_expect_tx(bar, tab_bytes + st_bytes);
    bulk_load(&s.sim, &p.tables->sim, tab_bytes, bar);
    bulk_load(stage, p.stage_init, st_bytes, bar);
  }
}
template <int TILE>
__device__ __forceinline__ void tables_wait(Smem<TILE>& s) { mbar_wait(smem_u32(&s.bar), 0u); }

// value of an agent / apple cell.  custom/ma_customenv.py:303-322 (step) / :198-209 (reset),
// custom/customenv.py:161-163 / :341-344 (single env: raw ids, every remaining apple).
__device__ __forceinline__ float special_value(int kind, bool fresh, int who, int k, bool apple_here) {
  float v = 0.0f;
  if (who >= 0) {
    if (fresh) v = 0.5f;                                                   // AddAgent marker, grid_world.py:140
    else if (apple_here || kind == GW_ENV_SINGLE) v = (float)(who + 1);    // id+9 is never remapped
    else v = (who == k) ? 1.0f : 5.0f;
  }
  return apple_here ? v + 9.0f : v;
}

template <int OBS>
__device__ __forceinline__ void store_cell(void* obs_base, long long elem, float v) {
  if (OBS == GW_OBS_F32) reinterpret_cast<float*>(obs_base)[elem] = v;
  else reinterpret_cast<__nv_bfloat16*>(obs_base)[elem] = __float2bfloat16(v);
}

// One env's observations (all learners) written by one warp.  The warp keeps a private copy of the constant
// observation template in shared memory; the <= 5 special cells per learner (agents, own apple) are patched in
// with scalar shared-memory stores, the rows leave as full 128-bit coalesced streaming stores (whole 128-byte lines,
// never a partial sector), and the patched cells are set back to 0 (agents and apples only stand on active cells).
template <int OBS>
__device__ __noinline__ void stage_and_store_env(uint8_t* stage, void* obs_base, long long e, int H, int n, int nl,
                                                    int kind, uint32_t cells, uint32_t apples_left, uint32_t apple_cells,
                                                    bool fresh, int lane) {
  const int cpo = H * GW_W;
  const int Q = (OBS == GW_OBS_F32) ? cpo / 4 : cpo / 8;
  // lane -> (learner k, item i): items 0..n-1 are the agents, item n is the learner's apple
  const int per = n + 1;
  const int k = lane >= per ? 1 : 0, i = lane - k * per;
  int my_cell = -1;
  float my_val = 0.0f;
  if (lane < nl * per) {
    const bool apple_on = (kind == GW_ENV_MULTI) ? ((apples_left >> k) & 1u) : (apples_left & 1u);
    const uint32_t apple = (kind == GW_ENV_MULTI) ? (apple_cells >> (8 * k)) & 0xFFu : apple_cells & 0xFFu;
    if (i < n) {
      const uint32_t c = (cells >> (8 * i)) & 0xFFu;
      my_cell = (int)c;
      my_val = special_value(kind, fresh, i, k, apple_on && c == apple);
    } else if (apple_on) {
      bool covered = false;
#pragma unroll
      for (int a = 0; a < 4; ++a) covered |= (a < n) && ((cells >> (8 * a)) & 0xFFu) == apple;
      if (!covered) { my_cell = (int)apple; my_val = 9.0f; }
    }
    if (my_cell >= 0) {
      if (OBS == GW_OBS_F32) reinterpret_cast<float*>(stage)[k * cpo + my_cell] = my_val;
      else reinterpret_cast<__nv_bfloat16*>(stage)[k * cpo + my_cell] = __float2bfloat16(my_val);
    }
  }
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(obs_base) + e * (long long)(nl * Q);
  const uint4* src = reinterpret_cast<const uint4*>(stage);
  const int V = nl * Q;
  uint4 v0, v1, v2, v3;
  if (lane < V) v0 = src[lane];
  if (lane + 32 < V) v1 = src[lane + 32];
  if (lane + 64 < V) v2 = src[lane + 64];
  if (lane + 96 < V) v3 = src[lane + 96];
  if (lane < V) __stcs(dst + lane, v0);
  if (lane + 32 < V) __stcs(dst + lane + 32, v1);
  if (lane + 64 < V) __stcs(dst + lane + 64, v2);
  if (lane + 96 < V) __stcs(dst + lane + 96, v3);
  __syncwarp();
  if (my_cell >= 0) {
    if (OBS == GW_OBS_F32) reinterpret_cast<float*>(stage)[k * cpo + my_cell] = 0.0f;
    else reinterpret_cast<__nv_bfloat16*>(stage)[k * cpo + my_cell] = __float2bfloat16(0.0f);
  }
  __syncwarp();
}

// Special cells of one env's observations, computed by the thread that owns the env (P1b / reset) and parked in
// shared memory for P4: entry = element offset inside the env's [n_learners, H*W] block | (2 * value) << 9.
__device__ __forceinline__ void encode_specials(uint16_t* out, int cpo, int n, int nl, int kind, uint32_t cells,
                                                uint32_t apples_left, uint32_t apple_cells, bool fresh) {
#pragma unroll
  for (int k = 0; k < GW_MAX_LEARNERS; ++k) {
    const bool apple_on = (k < nl) && ((kind == GW_ENV_MULTI) ? ((apples_left >> k) & 1u) : (apples_left & 1u));
    const uint32_t apple = (kind == GW_ENV_MULTI) ? (apple_cells >> (8 * k)) & 0xFFu : apple_cells & 0xFFu;
    bool covered = false;
#pragma unroll
    for (int i = 0; i < GW_MAX_AGENTS; ++i) {
      uint32_t enc = 0xFFFFu;
      if (k < nl && i < n) {
        const uint32_t c = (cells >> (8 * i)) & 0xFFu;
        const bool here = apple_on && c == apple;
        covered |= here;
        enc = (uint32_t)(k * cpo + (int)c) | ((uint32_t)(2.0f * special_value(kind, fresh, i, k, here)) << 9);
      }
      out[k * (GW_MAX_AGENTS + 1) + i] = (uint16_t)enc;
    }
    out[k * (GW_MAX_AGENTS + 1) + GW_MAX_AGENTS] =
        (apple_on && !covered) ? (uint16_t)((uint32_t)(k * cpo + (int)apple) | (18u << 9)) : (uint16_t)0xFFFFu;
  }
}

// P4: observations + action masks of the tile.  Each warp owns two staging rows that hold the constant template.
// Per env: lanes 0..9 drop the pre-computed special cells into row A, one __syncwarp, the cells patched into row B
// for the previous env are zeroed again (agents and apples only stand on active cells, whose template value is 0),
// and row A leaves as 128-bit streaming stores: whole 128-byte lines, never a partial sector.  Rows alternate.
template <int OBS>
__device__ __forceinline__ void patch_cell(uint8_t* row, uint32_t enc, bool set) {
  if (enc == 0xFFFFu) return;
  const float v = set ? 0.5f * (float)(enc >> 9) : 0.0f;
  if (OBS == GW_OBS_F32) reinterpret_cast<float*>(row)[enc & 0x1FFu] = v;
  else reinterpret_cast<__nv_bfloat16*>(row)[enc & 0x1FFu] = __float2bfloat16(v);
}

// ---- TMA bulk copies shared -> global (cp.async.bulk, bulk async-groups)
__device__ __forceinline__ void fence_proxy_async_smem() {    // generic-proxy writes to shared memory -> visible to the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_row(void* gdst, const void* ssrc, int bytes) {
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\ncp.async.bulk.commit_group;"
               ::"l"(gdst), "r"(src), "r"(bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {             // at most N of this thread's bulk groups still read shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// el0 / el_step / el_end: the envs of the tile this warp renders (big kernel: warp, +NWARPS, ...; small kernel: its own four).
template <int THREADS, int TILE, int OBS>
__device__ __forceinline__ void render_obs(Smem<TILE>& s, uint8_t* stage, const StepParams& p, long long tile_base,
                                           int el0, int el_step, int el_end) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cpo = p.H * GW_W;
  const int V = p.nl * ((OBS == GW_OBS_F32) ? cpo / 4 : cpo / 8);
  if (p.io.obs != nullptr) {
    // Rows leave through the TMA: lanes 0..9 patch the special cells into the warp's staging row, one lane issues a bulk
    // shared->global copy of the whole row (cp.async.bulk: whole 128-byte lines, no LSU traffic for the payload), and the
    // row is cleaned up two envs later, once the copy has read it.  Two rows alternate.
    const int row_bytes = V * 16;
    uint8_t* const row0 = stage + warp * 2 * row_bytes;
    uint8_t* dst = reinterpret_cast<uint8_t*>(p.io.obs) + (tile_base + el0) * (long long)row_bytes;
    const long long dst_step = (long long)el_step * row_bytes;
    uint32_t enc_a = 0xFFFFu, enc_b = 0xFFFFu;             // what this lane patched into row 0 / row 1
    int cur = 0;
    for (int el = el0; el < el_end; el += el_step, dst += dst_step) {
      if (s.rinfo[el] & R_SKIP) continue;                  // warp-uniform
      uint8_t* row = row0 + cur * row_bytes;
      const uint32_t enc = (lane < N_SPEC) ? (uint32_t)s.spec[el][lane] : 0xFFFFu;
      if (lane == 0) bulk_wait_read<1>();                  // the copy issued from this row two envs ago has read it
      __syncwarp();
      patch_cell<OBS>(row, cur ? enc_b : enc_a, false);
      __syncwarp();                                        // another lane may patch the cell this one just cleared
      patch_cell<OBS>(row, enc, true);
      if (cur) enc_b = enc; else enc_a = enc;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) bulk_store_row(dst, row, row_bytes);
      cur ^= 1;
    }
    if (lane == 0) bulk_wait_read<0>();                    // leave both rows clean for the next tile
    __syncwarp();
    patch_cell<OBS>(row0, enc_a, false);
    patch_cell<OBS>(row0 + row_bytes, enc_b, false);
    __syncwarp();
  }
  if (p.io.final_obs != nullptr) {                        // terminal observation of envs that were just re-spawned (rare)
    for (int el = el0; el < el_end; el += el_step) {
      const uint32_t ri = s.rinfo[el];
      if ((ri & R_FINAL) && !(ri & R_SKIP))
        stage_and_store_env<OBS>(stage + warp * 2 * V * 16, p.io.final_obs, tile_base + el, p.H, p.n, p.nl, p.kind, s.cells_fin[el],
                                 (ri >> 4) & 3u, p.apple_cells, false, lane);
    }
  }
}

template <int THREADS, int TILE, int OBS>
__device__ __forceinline__ void render_tile(Smem<TILE>& s, uint8_t* stage, const StepParams& p, long long tile_base,
                                            int tile_envs) {
  const int tid = threadIdx.x;
  render_obs<THREADS, TILE, OBS>(s, stage, p, tile_base, tid >> 5, THREADS / 32, tile_envs);
  if (p.io.action_mask != nullptr) {
    __syncthreads();                                       // s.mask was filled by the owner threads
    const int bytes = tile_envs * p.nl * GW_N_ACTIONS;
    int8_t* dst = p.io.action_mask + tile_base * (long long)(p.nl * GW_N_ACTIONS);
    const bool partial = p.reset_mask != nullptr;
    if (!partial && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      const int vecs = bytes / 16;
      for (int i = tid; i < vecs; i += THREADS) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(s.mask)[i];
      for (int i = vecs * 16 + tid; i < bytes; i += THREADS) dst[i] = (int8_t)s.mask[i];
    } else {
      for (int i = tid; i < bytes; i += THREADS)
        if (!partial || !(s.rinfo[i / (p.nl * GW_N_ACTIONS)] & R_SKIP)) dst[i] = (int8_t)s.mask[i];
    }
  }
}

// owner thread: action masks of the (new) positions into the tile's staging area (custom/ma_customenv.py:467-506)
template <int TILE>
__device__ __forceinline__ void stage_masks(Smem<TILE>& s, const StepParams& p, int tid, uint32_t cells) {
  if (p.io.action_mask == nullptr) return;
#pragma unroll
  for (int k = 0; k < GW_MAX_LEARNERS; ++k) {
    if (k >= p.nl) break;
    const uint32_t m = action_mask_bits(s.small.map_rows, p.H, (cells >> (8 * k)) & 0xFFu);
#pragma unroll
    for (int a = 0; a < GW_N_ACTIONS; ++a) s.mask[(tid * p.nl + k) * GW_N_ACTIONS + a] = (uint8_t)((m >> a) & 1u);
  }
}

// ------------------------------------------------------------------ spawn
// setup_env, custom/ma_customenv.py:372-380: a sorted n-subset of the active cells (row-major order).
// Replay mode reads the recorded cells; native mode draws them from Philox (uniform over subsets).
__device__ __forceinline__ uint32_t spawn_choose(const StepParams& p, long long e, uint32_t tick);
__device__ __forceinline__ uint32_t cells_from_chosen(const StepParams& p, const SmallTables& st, uint32_t chosen);

__device__ __forceinline__ uint32_t spawn_cells(const StepParams& p, const SmallTables& st, long long e, uint32_t tick) {
  if (p.io.spawn != nullptr) {
    uint32_t cells = 0;
    for (int i = 0; i < p.n; ++i) {
      // L2-level loads: under the resident kernel the array may sit in host memory and change between two steps
      const int r = __ldcg(p.io.spawn + (e * p.n + i) * 2 + 0), c = __ldcg(p.io.spawn + (e * p.n + i) * 2 + 1);
      cells |= (uint32_t)(((r & 15) << 4) | (c & 15)) << (8 * i);
    }
    return cells;
  }
  return cells_from_chosen(p, st, spawn_choose(p, e, tick));
}

// the n sorted indices into the active-cell list (needs no table: the small-batch kernel draws them while the tables load)
__device__ __forceinline__ uint32_t spawn_choose(const StepParams& p, long long e, uint32_t tick) {
  const unsigned long long gid = (unsigned long long)(p.env_id_base + e);
  uint32_t w[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), tick, 0x100u};
  philox4x32(w, p.seed_lo, p.seed_hi);
  uint32_t chosen = 0;                       // up to 4 sorted indices, 8 bits each
  const int na = p.n_active;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k >= p.n) break;
    uint32_t d = __umulhi(w[k], (uint32_t)(na - k));                      // k-th draw among the remaining cells
    int pos = 0;
#pragma unroll
    for (int t = 0; t < 3; ++t)                                           // chosen ascending
      if (t < k && d >= ((chosen >> (8 * t)) & 0xFFu)) { ++d; pos = t + 1; }
    const uint32_t lowmask = (pos == 0) ? 0u : (0xFFFFFFFFu >> (32 - 8 * pos));
    chosen = (chosen & lowmask) | (d << (8 * pos)) | ((chosen & ~lowmask) << 8);
  }
  return chosen;
}

__device__ __forceinline__ uint32_t cells_from_chosen(const StepParams& p, const SmallTables& st, uint32_t chosen) {
  uint32_t cells = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < p.n) cells |= (uint32_t)st.active_cell[(chosen >> (8 * i)) & 0xFFu] << (8 * i);
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

__device__ __forceinline__ void write_positions(int8_t* dst, long long e, int n, uint32_t cells) {
  if (dst == nullptr) return;
  if (n == 4) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t c = (cells >> (8 * i)) & 0xFFu, d = (cells >> (8 * (i + 2))) & 0xFFu;
      lo |= ((c >> 4) | ((c & 15u) << 8)) << (16 * i);
      hi |= ((d >> 4) | ((d & 15u) << 8)) << (16 * i);
    }
    reinterpret_cast<uint2*>(dst)[e] = make_uint2(lo, hi);
  } else {
    for (int i = 0; i < n; ++i) {
      const uint32_t c = (cells >> (8 * i)) & 0xFFu;
      dst[(e * n + i) * 2] = (int8_t)(c >> 4);
      dst[(e * n + i) * 2 + 1] = (int8_t)(c & 15u);
    }
  }
}

// Reward / done block of one env step: custom/ma_customenv.py:258-302 (multi) and custom/customenv.py:126-158 (single).
// caught0 / caught1: sub-steps learner k stood on its own apple (grid_world.py:531-540).
struct RewardOut {
  double reward[GW_MAX_LEARNERS];
  uint32_t meta, apples_left, term_now, trunc_now, apples_rewarded, crash_count, shaped;
};

__device__ __forceinline__ RewardOut env_rewards(const StepParams& p, int nl, uint32_t meta, uint32_t cells_new, uint32_t crash,
                                                 uint32_t caught0, uint32_t caught1) {
  RewardOut o;
  o.reward[0] = o.reward[1] = 0.0;
  const uint32_t apples_before = meta & M_APPLES;
  uint32_t apples_left = apples_before;
  uint32_t term_now = 0, trunc_now = 0, apples_rewarded = 0, crash_count = 0, shaped = 0;
  if (p.kind == GW_ENV_MULTI) {
    uint32_t term = (meta >> M_TERM_SH) & 3u, trunc = (meta & M_TRUNC) ? 1u : 0u;
    int ri[GW_MAX_LEARNERS] = {0, 0};
#pragma unroll
    for (int k = 0; k < GW_MAX_LEARNERS; ++k)                          // own apple only (:258-271)
      if (k < nl && ((apples_left >> k) & 1u) && (k == 0 ? caught0 : caught1)) {
        apples_left &= ~(1u << k);
        ri[k] += 20;
        ++apples_rewarded;
      }
    if (apples_rewarded && apples_left == 0) {                         // last apple: +20 to all, truncate (:272-275)
#pragma unroll
      for (int k = 0; k < GW_MAX_LEARNERS; ++k) if (k < nl) ri[k] += 20;
      trunc = 1;
    }
    uint32_t pdv = 0, pd[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < GW_MAX_LEARNERS; ++k) {
      if (k >= nl) continue;
      if ((crash >> k) & 1u) {                                         // :281-285
        ri[k] -= 10;
        ++crash_count;
        trunc = 1;
        term |= 1u << k;
      }
      if ((apples_left >> k) & 1u) {                                   // :287-300
        const uint32_t d = (uint32_t)manhattan((cells_new >> (8 * k)) & 0xFFu, (p.apple_cells >> (8 * k)) & 0xFFu);
        const uint32_t prev_valid = (meta >> (M_PDV_SH + k)) & 1u;
        const uint32_t prev = (meta >> (k == 0 ? M_PD0_SH : M_PD1_SH)) & 31u;
        if (prev_valid && prev > d) { ri[k] += 1; shaped |= 1u << k; }
        pdv |= 1u << k;
        pd[k] = d;
      }
      o.reward[k] = (double)ri[k];
    }
    term_now = term;
    trunc_now = trunc ? ((1u << nl) - 1u) : 0u;
    const uint32_t steps = min(((meta >> M_STEPS_SH) & M_STEPS_MASK) + 1u, M_STEPS_MASK);
    meta = apples_left | (term << M_TERM_SH) | (trunc ? M_TRUNC : 0u) | (pdv << M_PDV_SH) | (pd[0] << M_PD0_SH) |
           (pd[1] << M_PD1_SH) | (steps << M_STEPS_SH);
  } else {                                                             // customenv.py:126-158
    double rew = 0.0;
    const uint32_t apple = p.apple_cells & 0xFFu;
    const uint32_t d = (uint32_t)manhattan(cells_new & 0xFFu, apple);
    if (crash & 1u) { rew -= 10.0; term_now = 1; crash_count = 1; }
    if ((apples_left & 1u) && caught0 == 1u) {                         // len(apples_caught) == 1 (:143)
      apples_left &= ~1u;
      rew += 20.0;
      trunc_now = 1;
      apples_rewarded = 1;
    }
    const uint32_t prev = (meta >> M_PD0_SH) & 31u;
    if (d < prev) { rew += 0.1; shaped = 1; }                          // :157-158
    o.reward[0] = rew;
    const uint32_t steps = min(((meta >> M_STEPS_SH) & M_STEPS_MASK) + 1u, M_STEPS_MASK);
    meta = apples_left | (1u << M_PDV_SH) | (d << M_PD0_SH) | (steps << M_STEPS_SH);
  }
  o.meta = meta; o.apples_left = apples_left; o.term_now = term_now; o.trunc_now = trunc_now;
  o.apples_rewarded = apples_rewarded; o.crash_count = crash_count; o.shaped = shaped;
  return o;
}

// ------------------------------------------------------------------ reset kernel
template <int THREADS, int TILE, int OBS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) gw_reset_kernel(StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem<TILE>& s = *reinterpret_cast<Smem<TILE>*>(smem_raw);
  uint8_t* stage = smem_raw + smem_fixed_bytes<TILE>();
  load_tables<TILE>(s, stage, p, (THREADS / 32) * 2, p.nl * p.H * GW_W * (OBS == GW_OBS_F32 ? 4 : 2));
  tables_wait(s);
  const long long tile_base = (long long)blockIdx.x * TILE;
  const int tile_envs = (int)min((long long)TILE, p.E - tile_base);
  const int tid = threadIdx.x;
  if (tid < tile_envs) {
    const long long e = tile_base + tid;
    if (p.reset_mask != nullptr && p.reset_mask[e] == 0) {
      s.rinfo[tid] = R_SKIP;
    } else {
      const uint4 st = p.state[e];
      const uint32_t cells = spawn_cells(p, s.small, e, st.z);
      const uint32_t meta = fresh_meta(p, cells);
      s.rinfo[tid] = (meta & M_APPLES) | R_FRESH;
      if (p.io.obs_code) p.io.obs_code[e] = (unsigned long long)cells | ((unsigned long long)(meta & M_APPLES) << 32) | (1ull << 34);
      encode_specials(s.spec[tid], p.H * GW_W, p.n, p.nl, p.kind, cells, meta & M_APPLES, p.apple_cells, true);
      stage_masks(s, p, tid, cells);
      write_positions(p.io.positions, e, p.n, cells);
      p.state[e] = make_uint4(cells, meta, st.z + 1, 0u);
    }
  }
  __syncthreads();
  render_tile<THREADS, TILE, OBS>(s, stage, p, tile_base, tile_envs);
}

// ------------------------------------------------------------------ step kernel
// Persistent CTAs: the grid is sized to the machine and each CTA walks tiles blockIdx.x, +gridDim.x, ...
// Phase stamps are a dev facility: compiled in only with -DGW_ENABLE_TRACE (scripts/trace_phases.py builds that variant).
__device__ __forceinline__ void trace_stamp(const StepParams& p, int slot) {
#ifndef GW_ENABLE_TRACE
  (void)p; (void)slot;
  return;
#endif
  if (p.trace != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[(size_t)blockIdx.x * 16 + slot] = t;
  }
}

template <int THREADS, int TILE, bool FEAR, int OBS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) gw_step_kernel(StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem<TILE>& s = *reinterpret_cast<Smem<TILE>*>(smem_raw);
  uint8_t* stage = smem_raw + smem_fixed_bytes<TILE>();
  trace_stamp(p, 0);
  // Programmatic dependent launch (when the launch carries the attribute): this grid's prologue (table copies) may run
  // while the previous step drains; the env state it wrote is only touched after griddepcontrol.wait.
  load_tables<TILE>(s, stage, p, (THREADS / 32) * 2, p.nl * p.H * GW_W * (OBS == GW_OBS_F32 ? 4 : 2));
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tid = threadIdx.x;
  const int n = p.n, nl = p.nl;
  const long long n_tiles = (p.E + TILE - 1) / TILE;
  bool tables_pending = true;

  // Tiles: the first one is blockIdx.x, the following ones come from a global counter, so that CTAs on SMs that run
  // ahead take more tiles and the grid drains together (a fixed stride left ~15 % of the SM cycles idle at the tail).
  // (measured: pays from ~4 tiles per CTA on; with fewer the fixed stride, which gives every SM the same count, is better)
  const bool dyn = p.dyn_tiles && n_tiles >= 4 * (long long)gridDim.x;
  for (long long tile = blockIdx.x; tile < n_tiles;) {
    const long long tile_base = tile * TILE;
    const int tile_envs = (int)min((long long)TILE, p.E - tile_base);
    const long long e = tile_base + tid;
    const bool own = tid < tile_envs;
    unsigned int fetched = 0;                              // issued now, first used right before P4's barrier: no stall on it
    if (dyn && tid == 0) fetched = atomicAdd(&p.tile_ctr[0], 1u);

    // per-env registers that live across the phases
    uint32_t task_bits = 0;            // bit (x*4 + j): (x, j) is an enqueued FeAR task
    double reward[GW_MAX_LEARNERS] = {0.0, 0.0};
    uint32_t stat_bits = 0, steps_now = 0;
    int ret0 = 0, ret1 = 0;
    uint4 st = make_uint4(0, 0, 0, 0);
    uint32_t acts = 0, mdrs = 0, cells_render = 0;
    uint32_t rw[4] = {0, 0, 0, 0}, rw2[4] = {0, 0, 0, 0};

    // ================================================================= P1a: loads and RNG (no shared tables needed)
    if (tid == 0) s.qn = 0;
    if (own) {
      st = p.state[e];
      const uint32_t tick = st.z;
      // ---- setup_step (ma_customenv.py:432-452): learner actions (:239-242), recorded NPC draws or Philox words
#pragma unroll
      for (int k = 0; k < GW_MAX_LEARNERS; ++k)
        if (k < nl) acts |= (uint32_t)min(max((int)p.io.learner_actions[e * nl + k], 0), 8) << (4 * k);
      if (p.io.npc_actions != nullptr) {
#pragma unroll
        for (int i = 1; i < 4; ++i)
          if (i >= nl && i < n) acts |= (uint32_t)min(max((int)p.io.npc_actions[e * n + i], 0), 8) << (4 * i);
      } else {
        // NPC number m = i - n_learners uses Philox call m/2 (counter = global env id, tick, call), words 2(m%2), 2(m%2)+1
        const unsigned long long gid = (unsigned long long)(p.env_id_base + e);
        rw[0] = (uint32_t)gid; rw[1] = (uint32_t)(gid >> 32); rw[2] = tick; rw[3] = 0u;
        philox4x32(rw, p.seed_lo, p.seed_hi);
        if (n - nl > 2) {
          rw2[0] = (uint32_t)gid; rw2[1] = (uint32_t)(gid >> 32); rw2[2] = tick; rw2[3] = 1u;
          philox4x32(rw2, p.seed_lo, p.seed_hi);
        }
      }
    }
    if (tables_pending) {
      tables_wait(s);
      tables_pending = false;
    }
    __syncthreads();

    trace_stamp(p, 1);
    // ================================================================= P1b: world update, rewards, flags, FeAR tasks
    if (own) {
      const uint32_t cells = st.x;
      uint32_t meta = st.y;
      const uint32_t tick = st.z;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < n) mdrs |= (uint32_t)s.small.mdr_map[(cells >> (8 * i)) & 0xFFu] << (4 * i);   // :445-447
      if (p.io.npc_actions == nullptr) {
#pragma unroll
        for (int i = 1; i < 4; ++i) {
          if (i < nl || i >= n) continue;
          const int m = i - nl;
          const uint32_t wa = m == 0 ? rw[0] : (m == 1 ? rw[2] : rw2[0]);
          const uint32_t wb = m == 0 ? rw[1] : (m == 1 ? rw[3] : rw2[1]);
          const int pert = wa < p.perturb_thr ? 1 : 0;                     // random.random() < 0.25 (:441)
          const uint32_t c = (cells >> (8 * i)) & 0xFFu;
          const uint4* thr4 = reinterpret_cast<const uint4*>(s.small.policy_thr[s.small.policy_map[c]][pert]);
          const uint4 t0 = thr4[0], t1 = thr4[1];
          const uint32_t u = wb >> 1;
          const uint32_t a = (u >= t0.x) + (u >= t0.y) + (u >= t0.z) + (u >= t0.w) + (u >= t1.x) + (u >= t1.y) +
                             (u >= t1.z) + (u >= t1.w);                    // np.random.choice(9, p) (custom_agent.py:31)
          acts |= a << (4 * i);
        }
      }
      trace_stamp(p, 8);
      const PairGeom g = pair_geometry(s.sim, n, cells);
      trace_stamp(p, 9);

      // ---- the real update (ma_customenv.py:254)
      const uint32_t apples_before = meta & M_APPLES;
      const StepResult r = world_update<true>(s.sim, n, cells, acts, g, p.apple_cells, apples_before, nl);
      const uint32_t cells_new = r.cells;
      trace_stamp(p, 10);

      // ---- FeAR tasks on the pre-step positions (ma_customenv.py:245-252 / customenv.py:113-120)
      if (FEAR) {
        uint32_t closew = 0, effw = r.effs;
#pragma unroll
        for (int x = 0; x < GW_MAX_LEARNERS; ++x) {
          if (x >= nl) break;
          const uint32_t ax = (acts >> (4 * x)) & 0xFu, mx = (mdrs >> (4 * x)) & 0xFu;
          // action == MdR: both counts are equal -> Resp = 0 exactly.  Agents that no chain of near pairs links to the
          // actor cannot be influenced by its move -> equal counts -> 0 as well.
          if (ax == mx) continue;
          const uint32_t js = reach_mask(s.sim, g.near6, x) & ~(1u << x) & ((1u << n) - 1u);
          if (js == 0) continue;
          uint32_t close = 0;                                              // close_agents :456-464
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < n && (k == x || manhattan((cells >> (8 * x)) & 0xFFu, (cells >> (8 * k)) & 0xFFu) <= p.fear_radius))
              close |= 1u << k;
          const uint32_t eff_mdr = make_traj(s.sim.next, (cells >> (8 * x)) & 0xFFu, mx).eff;   // actor plays its MdR
          if (eff_mdr == ((r.effs >> (4 * x)) & 0xFu)) continue;           // same trajectory (e.g. both blocked): counts equal
          closew |= close << (4 * x);
          effw |= eff_mdr << (16 + 4 * x);
          const uint32_t slot = atomicAdd(&s.qn, (uint32_t)__popc(js));
          uint32_t k2 = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((js >> j) & 1u) {
              s.queue[slot + k2] = (uint16_t)(tid | (x << 8) | (j << 9) | (((close >> j) & 1u) << 11));
              ++k2;
            }
          task_bits |= js << (4 * x);
        }
        if (task_bits) {
          s.cells_old[tid] = cells;
          s.effs[tid] = effw;
          s.geom_lo[tid] = g.didx_lo;
          s.geom_hi[tid] = g.didx_hi | (g.near6 << 16);
          s.close[tid] = closew;
          s.cnt[tid * 2] = 0;
          s.cnt[tid * 2 + 1] = 0;
        }
      }

      trace_stamp(p, 11);
      const RewardOut ro = env_rewards(p, nl, meta, cells_new, r.crash, r.caught & 7u, (r.caught >> 9) & 7u);
      reward[0] = ro.reward[0];
      reward[1] = ro.reward[1];
      meta = ro.meta;
      const uint32_t apples_left = ro.apples_left, term_now = ro.term_now, trunc_now = ro.trunc_now,
                     apples_rewarded = ro.apples_rewarded, crash_count = ro.crash_count, shaped = ro.shaped;
      steps_now = (meta >> M_STEPS_SH) & M_STEPS_MASK;
      const bool episode_over = (p.kind == GW_ENV_MULTI) ? (trunc_now != 0) : ((term_now | trunc_now) != 0);
      const bool ended = episode_over || (p.max_steps > 0 && (int)steps_now >= p.max_steps);

      trace_stamp(p, 12);
      // ---- scalar outputs (one thread per env: each array is written with unit stride across the warp)
      if (nl == 2) {
        if (p.io.reward) reinterpret_cast<float2*>(p.io.reward)[e] = make_float2((float)reward[0], (float)reward[1]);
        if (p.io.terminated) reinterpret_cast<uchar2*>(p.io.terminated)[e] = make_uchar2(term_now & 1u, (term_now >> 1) & 1u);
        if (p.io.truncated) reinterpret_cast<uchar2*>(p.io.truncated)[e] = make_uchar2(trunc_now & 1u, (trunc_now >> 1) & 1u);
      } else {
        if (p.io.reward) p.io.reward[e] = (float)reward[0];
        if (p.io.terminated) p.io.terminated[e] = (uint8_t)(term_now & 1u);
        if (p.io.truncated) p.io.truncated[e] = (uint8_t)(trunc_now & 1u);
      }
      write_positions(p.io.positions, e, n, cells_new);
      if (p.io.ended) p.io.ended[e] = ended ? 1 : 0;
      if (p.io.info)
        p.io.info[e] = (r.crash & 15u) | ((r.restr & 15u) << 4) | (crash_count << 8) | (apples_rewarded << 10) |
                       ((ended ? 1u : 0u) << 12) | (shaped << 14);

      trace_stamp(p, 13);
      // ---- episode return (reward units: 1 multi, 0.1 single), state for the next step, render record
      ret0 = (int)(short)(st.w & 0xFFFFu);
      ret1 = (int)(short)(st.w >> 16);
      const double unit = (p.kind == GW_ENV_MULTI) ? 1.0 : 10.0;
      ret0 += (int)lrint(reward[0] * unit);
      ret1 += (int)lrint(reward[1] * unit);
      stat_bits = (ended ? 1u : 0u) | (crash_count << 1) | (apples_rewarded << 3);
      uint32_t cells_r = cells_new, apples_r = apples_left, rflags = 0;
      uint4 st_out = make_uint4(cells_new, meta, tick + 1, ((uint32_t)ret0 & 0xFFFFu) | ((uint32_t)ret1 << 16));
      if (ended && p.auto_reset) {
        cells_r = spawn_cells(p, s.small, e, tick);
        const uint32_t meta_sp = fresh_meta(p, cells_r);
        apples_r = meta_sp & M_APPLES;
        rflags = R_FRESH | R_FINAL | (apples_left << 4);
        s.cells_fin[tid] = cells_new;
        st_out = make_uint4(cells_r, meta_sp, tick + 1, 0u);
      }
      trace_stamp(p, 14);
      p.state[e] = st_out;
      if (p.io.obs_code)
        p.io.obs_code[e] = (unsigned long long)cells_r | ((unsigned long long)apples_r << 32) |
                           ((rflags & R_FRESH) ? (1ull << 34) : 0ull);
      s.rinfo[tid] = apples_r | rflags;
      encode_specials(s.spec[tid], p.H * GW_W, n, nl, p.kind, cells_r, apples_r, p.apple_cells, (rflags & R_FRESH) != 0);
      cells_render = cells_r;
    }

    // ================================================================= P2: counterfactual sims (Responsibility.py:20-54)
    trace_stamp(p, 2);
    if (FEAR) {
      __syncthreads();
      trace_stamp(p, 3);
      // One thread per (task, actor variant): the nine counterfactual actions of the affected agent are counted at once.
      const uint32_t n_work = s.qn * 2u;
      if (tid == 0 && s.qn) atomicAdd(&p.stats[(int)(tile & (STAT_SLOTS - 1)) * 8 + ST_TASKS], (unsigned long long)s.qn);
      for (uint32_t w = tid; w < n_work; w += THREADS) {
        const uint32_t tk = s.queue[w >> 1], v = w & 1u;
        const uint32_t el = tk & 0xFFu, x = (tk >> 8) & 1u, j = (tk >> 9) & 3u, jc = (tk >> 11) & 1u;
        const uint32_t effw = s.effs[el];
        const uint32_t close = (s.close[el] >> (4 * x)) & 0xFu;
        const uint32_t keep = ((close & 1u) ? 0xFu : 0u) | ((close & 2u) ? 0xF0u : 0u) | ((close & 4u) ? 0xF00u : 0u) |
                              ((close & 8u) ? 0xF000u : 0u);
        uint32_t eo = effw & keep;                       // agents outside the close list Stay (defaultAction='stay')
        if (v == 0) eo = (eo & ~(0xFu << (4 * x))) | (((effw >> (16 + 4 * x)) & 0xFu) << (4 * x));   // actor plays its MdR (:169-172)
        PairGeom g;
        g.didx_lo = s.geom_lo[el];
        g.didx_hi = s.geom_hi[el] & 0xFFFFu;
        g.near6 = s.geom_hi[el] >> 16;
        const uint32_t cnt = count_valid_moves_p6(s.sim, s.cells_old[el], eo, g, (int)j, jc != 0);
        const uint32_t jslot = j - (j > x ? 1u : 0u);
        if (cnt) atomicAdd(&s.cnt[el * 2 + x], cnt << (4 * (jslot * 2 + v)));
      }
      __syncthreads();
    }

    trace_stamp(p, 4);
    // ================================================================= P3: fear, shaped reward, statistics, action masks
    if (own) {
      stage_masks(s, p, tid, cells_render);                // the FeAR scratch is dead now: its space stages the masks
      double fear[GW_MAX_LEARNERS] = {0.0, 0.0};
      if (FEAR) {
#pragma unroll
        for (int x = 0; x < GW_MAX_LEARNERS; ++x) {
          if (x >= nl) break;
          const uint32_t tb = (task_bits >> (4 * x)) & 0xFu;
          if (tb == 0) continue;
          const uint32_t c = s.cnt[tid * 2 + x];
          double rs[3] = {0.0, 0.0, 0.0};
#pragma unroll
          for (int js = 0; js < 3; ++js) {
            const int j = js + (js >= x ? 1 : 0);
            if ((tb >> j) & 1u) rs[js] = s.small.resp_lut[(c >> (8 * js)) & 0xFu][(c >> (8 * js + 4)) & 0xFu];
          }
          fear[x] = fear_sum3(n, rs[0], rs[1], rs[2]);
        }
      }
      if (nl == 2) {
        if (p.io.fear) reinterpret_cast<double2*>(p.io.fear)[e] = make_double2(fear[0], fear[1]);
        if (p.io.shaped_reward)
          reinterpret_cast<float2*>(p.io.shaped_reward)[e] =
              make_float2((float)(p.fear_weight * fear[0] + reward[0]), (float)(p.fear_weight * fear[1] + reward[1]));   // maddpg/agent.py:130
      } else {
        if (p.io.fear) p.io.fear[e] = fear[0];
        if (p.io.shaped_reward) p.io.shaped_reward[e] = (float)(p.fear_weight * fear[0] + reward[0]);
      }
      // statistics: most lanes contribute nothing, so reduce over the warp first
      const uint32_t ended = stat_bits & 1u, crashes = (stat_bits >> 1) & 3u, apples = (stat_bits >> 3) & 3u;
      const int nz = FEAR ? ((fear[0] != 0.0) + (fear[1] != 0.0)) : 0;
      const unsigned act = __activemask();
      const unsigned any = __ballot_sync(act, ended | crashes | apples | (uint32_t)nz);
      if (any) {
        const double unit = (p.kind == GW_ENV_MULTI) ? 1.0 : 10.0;
        const int slot = (int)((e >> 5) & (STAT_SLOTS - 1));
        const int lane = tid & 31;
        const int leader = __ffs(act) - 1;
        const unsigned w_end = __reduce_add_sync(act, ended), w_len = __reduce_add_sync(act, ended ? steps_now : 0u);
        const unsigned w_cr = __reduce_add_sync(act, crashes), w_ap = __reduce_add_sync(act, apples);
        const unsigned w_nz = __reduce_add_sync(act, (unsigned)nz);
        const int w_ret = __reduce_add_sync(act, ended ? (ret0 + ret1) : 0);
        if (lane == leader) {
          if (w_end) {
            atomicAdd(&p.stats[slot * 8 + ST_EPISODES], (unsigned long long)w_end);
            atomicAdd(&p.stats[slot * 8 + ST_LEN], (unsigned long long)w_len);
            atomicAdd(&p.stats[slot * 8 + ST_RETURN_MILLI], (unsigned long long)(long long)llrint(w_ret * (1000.0 / unit)));
          }
          if (w_cr) atomicAdd(&p.stats[slot * 8 + ST_CRASH], (unsigned long long)w_cr);
          if (w_ap) atomicAdd(&p.stats[slot * 8 + ST_APPLES], (unsigned long long)w_ap);
          if (w_nz) atomicAdd(&p.stats[slot * 8 + ST_FEAR_NZ], (unsigned long long)w_nz);
        }
        if (nz) atomicAdd(reinterpret_cast<double*>(&p.stats[slot * 8 + ST_FEAR_BITS]), fear[0] + fear[1]);
      }
    }

    trace_stamp(p, 5);
    // ================================================================= P4
    if (tid == 0) s.next_tile = dyn ? (long long)gridDim.x + (long long)fetched : n_tiles;
    __syncthreads();
    trace_stamp(p, 6);
    // last tile of this CTA: only observation stores are left, the next step's grid may start its prologue
    const long long next = s.next_tile;
    if (next >= n_tiles) asm volatile("griddepcontrol.launch_dependents;");
    render_tile<THREADS, TILE, OBS>(s, stage, p, tile_base, tile_envs);
    __syncthreads();                                       // shared arrays are reused by the next tile
    trace_stamp(p, 7);
    tile = dyn ? next : tile + gridDim.x;
  }
  if (tables_pending) tables_wait(s);                      // a CTA without tiles must not exit with copies in flight
  if (dyn && tid == 0 && atomicAdd(&p.tile_ctr[1], 1u) == gridDim.x - 1) {    // last CTA out: counters ready for the next launch
    p.tile_ctr[0] = 0;
    p.tile_ctr[1] = 0;
    __threadfence();
  }
}

// ------------------------------------------------------------------ step kernel for small batches (latency regime)
// One env per 8 lanes, 32 envs per CTA.  With a few thousand envs the thread-per-env kernel above leaves one warp per
// SM walking a ~1500-instruction dependent chain; here the per-agent and per-pair work of an env runs on different
// lanes (lanes 0-3: the agents, lanes 4-5: the learners' Move-de-Rigueur trajectories, lanes 0-5: the six pairs) and
// the words the fix-point needs are gathered with redux.sync over the 8-lane group.  The env's FeAR tasks are counted
// by the same 8 lanes (one (task, variant, three actions) item each), so there is no CTA-wide queue and no barrier
// after the table copy: every warp steps and renders its own four envs and leaves.
__device__ __forceinline__ uint32_t special_entry(int q, int cpo, int n, int nl, int kind, uint32_t cells, uint32_t apples_left,
                                                  uint32_t apple_cells, bool fresh) {
  const int k = q >= (GW_MAX_AGENTS + 1) ? 1 : 0, i = q - k * (GW_MAX_AGENTS + 1);
  const bool apple_on = (k < nl) && ((kind == GW_ENV_MULTI) ? ((apples_left >> k) & 1u) : (apples_left & 1u));
  const uint32_t apple = (kind == GW_ENV_MULTI) ? (apple_cells >> (8 * k)) & 0xFFu : apple_cells & 0xFFu;
  if (i < GW_MAX_AGENTS) {
    if (k >= nl || i >= n) return 0xFFFFu;
    const uint32_t c = (cells >> (8 * i)) & 0xFFu;
    const bool here = apple_on && c == apple;
    // twice special_value(): 0.5 fresh marker / raw id (apple cell, single env) / 1 self / 5 other; +9 on the apple
    const uint32_t v2 = (fresh ? 1u : ((here || kind == GW_ENV_SINGLE) ? 2u * (uint32_t)(i + 1) : (i == k ? 2u : 10u))) + (here ? 18u : 0u);
    return (uint32_t)(k * cpo + (int)c) | (v2 << 9);
  }
  bool covered = false;
#pragma unroll
  for (int a = 0; a < GW_MAX_AGENTS; ++a) covered |= (a < n) && ((cells >> (8 * a)) & 0xFFu) == apple;
  return (apple_on && !covered) ? ((uint32_t)(k * cpo + (int)apple) | (18u << 9)) : 0xFFFFu;
}

// OR of the BITS-wide values held by lanes base .. base+N-1, value i shifted to bit BITS*i.  N independent shuffles: on a
// latency-bound warp they pipeline, unlike a butterfly (and redux.sync takes its slow path for a part-warp mask).
template <int N, int BITS>
__device__ __forceinline__ uint32_t gather_lanes(unsigned gmask, uint32_t v, int base) {
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) w |= __shfl_sync(gmask, v, base + i) << (BITS * i);
  return w;
}

// Philox words of one env step: NPC call 0 (rw), NPC call 1 (rw2, three NPCs only) and the re-spawn draw.
__device__ __forceinline__ void draw_step_randoms(const StepParams& p, long long e, uint32_t tick, bool three_npcs, uint32_t rw[4],
                                                  uint32_t rw2[4], uint32_t& chosen) {
  const unsigned long long gid = (unsigned long long)(p.env_id_base + e);
  rw[0] = (uint32_t)gid; rw[1] = (uint32_t)(gid >> 32); rw[2] = tick; rw[3] = 0u;
  philox4x32(rw, p.seed_lo, p.seed_hi);
  if (three_npcs) {
    rw2[0] = (uint32_t)gid; rw2[1] = (uint32_t)(gid >> 32); rw2[2] = tick; rw2[3] = 1u;
    philox4x32(rw2, p.seed_lo, p.seed_hi);
  }
  chosen = spawn_choose(p, e, tick);
}

// One result array of a tile, shared memory -> (host or device) global memory: a single bulk copy when the destination
// and the size allow it (16-byte granules), plain stores otherwise (tile tails, odd alignments).  Called by a whole warp.
__device__ __forceinline__ void server_flush(void* gdst, const void* ssrc, int bytes, int lane) {
  if (((reinterpret_cast<uintptr_t>(gdst) | (uintptr_t)bytes) & 15u) == 0) {
    if (lane == 0 && bytes > 0) bulk_store_row(gdst, ssrc, bytes);
  } else if (((reinterpret_cast<uintptr_t>(gdst) | (uintptr_t)bytes) & 3u) == 0) {
    for (int i = lane; i < bytes / 4; i += 32) static_cast<uint32_t*>(gdst)[i] = static_cast<const uint32_t*>(ssrc)[i];
  } else {
    for (int i = lane; i < bytes; i += 32) static_cast<uint8_t*>(gdst)[i] = static_cast<const uint8_t*>(ssrc)[i];
  }
}

// gw_rollout: several steps per launch (see the ROLL notes in small_step_tiles)
struct RollParams {
  int steps;
  long long ring_slots, first_slot;       // time-major output rings: slot of step 0's transition
  long long action_slots, first_action;   // time-major action arrays: slot of step 0's actions
  int wait_full;                          // steps > ring_slots - 1: a slot is written more than once per launch
};

// The step of the CTA's tiles (tile = blockIdx.x, += gridDim.x).  SERVER: called once per command by the resident kernel
// below -- inputs that the host rewrites between two calls are loaded past the L1 (ordinary L2-level loads after the
// system-scope acquire fence that follows the doorbell: `ld.volatile` reads of host memory are served one PCIe round
// trip after the other, ~50 ns per request, measured 50-100 us per step at 4096 envs).
// ---- FeAR of one step for the warp's four envs (ma_customenv.py:245-252, Responsibility.py:135-210).  Inputs are per lane
// what the step computed before the world update: the env's packed cells, the trajectories played / the learners' MdR
// trajectories (effw), which learners deviate from their MdR (neqb), the near pairs (near6) and this lane's pair geometry
// (didx).  Nothing here depends on the update itself, which is what lets gw_rollout's helper warps run it beside the update.
struct FearOut { double f0, f1; uint32_t tasks; };
__device__ __forceinline__ FearOut fear_block(const StepParams& p, Smem<32>& s, int lane, int gsh, int r, int n, int nl, bool own,
                                              uint32_t cells, uint32_t effw, uint32_t neqb, uint32_t near6, uint32_t didx) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  double fear0 = 0.0, fear1 = 0.0;
  uint32_t tm = 0;                                               // bit 4x + j: task (actor x, affected j)
  {
    const int xx = r >> 2, kk = r & 3;                             // close_agents :456-464 for both actors at once
    const bool cl = xx < nl && kk < n &&
                    (kk == xx || manhattan((cells >> (8 * xx)) & 0xFFu, (cells >> (8 * kk)) & 0xFFu) <= p.fear_radius);
    const uint32_t closeb = (__ballot_sync(FULL, cl) >> gsh) & 0xFFu;
#pragma unroll
    for (int x = 0; x < GW_MAX_LEARNERS; ++x) {
      if (x >= nl) break;
      if (((neqb >> x) & 1u) == 0) continue;                                            // action == MdR: Resp = 0 exactly
      const uint32_t js = reach_mask(s.sim, near6, x) & ~(1u << x) & ((1u << n) - 1u);
      if (js == 0) continue;
      if (((effw >> (16 + 4 * x)) & 0xFu) == ((effw >> (4 * x)) & 0xFu)) continue;      // same trajectory: counts equal
      tm |= js << (4 * x);
    }
    if (!own) tm = 0;
    // Work items (task, variant) of the warp's four envs are dealt to its 32 lanes, whichever env they belong to: a
    // lane fetches the env's words from that group's first lane, counts (count_valid_moves: no per-action loop), and
    // the counts travel back as warp-wide sums per group.  One round unless the four envs hold more than 16 tasks.
    uint32_t cw0 = 0, cw1 = 0;
    const int n_it = 2 * __popc(tm);
    const int c0 = __shfl_sync(FULL, n_it, 0), c1 = __shfl_sync(FULL, n_it, 8), c2 = __shfl_sync(FULL, n_it, 16),
              c3 = __shfl_sync(FULL, n_it, 24);
    const int total = c0 + c1 + c2 + c3;
    if (total > 0) {                                               // warp-uniform
      const uint32_t my_lo = gather_lanes<4, 8>(FULL, didx, gsh), my_hi = gather_lanes<2, 8>(FULL, didx, gsh + 4);
      const uint32_t my_misc = closeb | (tm << 8) | (near6 << 16);
      uint32_t f0 = 0, f1 = 0;                                     // this lane's contributions (to the env of its item)
      int item_g = 0;
      for (int base = 0; base < total; base += 32) {
        const int it = base + lane;
        const int gq = (it >= c0) + (it >= c0 + c1) + (it >= c0 + c1 + c2);
        const int local = it - (gq > 0 ? c0 : 0) - (gq > 1 ? c1 : 0) - (gq > 2 ? c2 : 0);
        const int src = 8 * gq;
        const uint32_t cells_g = __shfl_sync(FULL, cells, src), effw_g = __shfl_sync(FULL, effw, src),
                       misc_g = __shfl_sync(FULL, my_misc, src), lo_g = __shfl_sync(FULL, my_lo, src),
                       hi_g = __shfl_sync(FULL, my_hi, src);
        if (it < total) {
          const uint32_t v = (uint32_t)local & 1u;
          uint32_t mm = (misc_g >> 8) & 0xFFu;
          for (int q = 0; q < (local >> 1); ++q) mm &= mm - 1u;
          const uint32_t bit = (uint32_t)__ffs(mm) - 1u, x = bit >> 2, j = bit & 3u;
          const uint32_t close = (misc_g >> (4 * x)) & 0xFu;
          const uint32_t keep = ((close & 1u) ? 0xFu : 0u) | ((close & 2u) ? 0xF0u : 0u) | ((close & 4u) ? 0xF00u : 0u) |
                                ((close & 8u) ? 0xF000u : 0u);
          uint32_t eo = effw_g & keep;                             // agents outside the close list Stay (defaultAction='stay')
          if (v == 0) eo = (eo & ~(0xFu << (4 * x))) | (((effw_g >> (16 + 4 * x)) & 0xFu) << (4 * x));   // actor plays its MdR (:169-172)
          PairGeom g;
          g.near6 = (misc_g >> 16) & 0x3Fu;
          g.didx_lo = lo_g;
          g.didx_hi = hi_g;
          const uint32_t cnt = count_valid_moves(s.sim, cells_g, eo, g, (int)j, ((close >> j) & 1u) != 0);
          const uint32_t jslot = j - (j > x ? 1u : 0u);
          const uint32_t field = cnt << (4 * (jslot * 2 + v));
          if (x == 0) f0 += field; else f1 += field;
          item_g = gq;
        }
        // the sums of this round go back to their groups (4-bit fields, at most 9 each: no carry)
        uint32_t cw0r = 0, cw1r = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t s0 = __reduce_add_sync(FULL, (it < total && item_g == q) ? f0 : 0u);
          const uint32_t s1 = __reduce_add_sync(FULL, (it < total && item_g == q) ? f1 : 0u);
          if ((lane >> 3) == q) { cw0r = s0; cw1r = s1; }
        }
        cw0 += cw0r; cw1 += cw1r;
        f0 = f1 = 0;
      }
#pragma unroll
      for (int x = 0; x < GW_MAX_LEARNERS; ++x) {
        const uint32_t tb = (tm >> (4 * x)) & 0xFu;
        if (tb == 0) continue;
        const uint32_t c = x == 0 ? cw0 : cw1;
        double rs[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int js = 0; js < 3; ++js) {
          const int j = js + (js >= x ? 1 : 0);
          if ((tb >> j) & 1u) rs[js] = s.small.resp_lut[(c >> (8 * js)) & 0xFu][(c >> (8 * js + 4)) & 0xFu];
        }
        const double f = fear_sum3(n, rs[0], rs[1], rs[2]);
        if (x == 0) fear0 = f; else fear1 = f;
      }
      
    }
  }
  FearOut o;
  o.f0 = fear0; o.f1 = fear1; o.tasks = (uint32_t)__popc(tm);
  return o;
}

// gw_rollout with FeAR, split form: warp w of the CTA's first eight steps its four envs, warp w + 8 computes their FeAR beside
// it.  The pair talk through a per-warp mailbox in shared memory and two mbarriers (inputs written / results written).
struct SplitBox {
  uint32_t in[8][4][32];                   // per lane of the main warp: cells, effw, neqb | near6 << 2 | own << 8, didx
  uint32_t rin[8][9][32];                  // the render message (see render_post): what the step's output stage needs
  double fear[8][4][2];                    // per env of the warp: info["fear"] of learner 0 / 1
  uint32_t tasks[8][4];
  unsigned long long full[8], done[8];
};
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// The step's own 256 threads: with helper warps in the CTA (SPLIT) a CTA-wide barrier would wait for them too.
template <bool SPLIT>
__device__ __forceinline__ void main_sync() {
  if (SPLIT) asm volatile("bar.sync 1, 256;" ::: "memory");
  else __syncthreads();
}
// the main warps tell their helpers to leave (gw_*_split_kernel, at the end of the kernel)
__device__ __forceinline__ void split_post_exit(SplitBox* box) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  box->in[warp][2][lane] = 0x200u;
  __syncwarp();
  if (lane == 0) mbar_arrive(smem_u32(&box->full[warp]));
}

template <bool FEAR, int OBS, bool SERVER, bool ROLL = false, bool SPLIT = false>
__device__ __forceinline__ void small_step_tiles(const StepParams& p, Smem<32>& s, uint8_t* stage, bool& tables_pending,
                                                 const RollParams* rp = nullptr, SplitBox* box = nullptr, uint32_t* split_phase_io = nullptr) {
  uint32_t split_phase = SPLIT ? *split_phase_io : 0u;              // parity of the helper's "results written" barrier: lives across calls
  constexpr bool RS = SPLIT && !SERVER;      // the helper warp also stores the observation rows (fear_helper_loop)
  constexpr int TILE = 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = tid & 7, gsh = lane & 24;
  constexpr unsigned FULL = 0xFFFFFFFFu;     // every shuffle / ballot below is executed by the whole converged warp: a
                                             // part-warp mask would send each of them through the compiler's slow path
  const int n = p.n, nl = p.nl;
  const long long n_tiles = (p.E + TILE - 1) / TILE;
  const int cpo = p.H * GW_W;
  const int Q = (OBS == GW_OBS_F32) ? cpo / 4 : cpo / 8;               // 16-byte vectors per learner observation
  const int V = nl * Q, row_bytes = V * 16;
  uint8_t* const rows4 = stage + (size_t)warp * 4 * row_bytes;         // this warp's staging rows
  uint8_t* const myrow = rows4 + (lane >> 3) * row_bytes;              // this group's env
  if (!SERVER && !ROLL && p.pdl_early) asm volatile("griddepcontrol.launch_dependents;");

  // SERVER: rewards / shaped rewards / ended flags (the results the host reads) are collected per tile and leave as one
  // bulk copy per array: written lane by lane into host memory they are ~12k small PCIe writes per step (measured: 40 us).
  __shared__ __align__(16) float out_rew[SERVER ? TILE * GW_MAX_LEARNERS : 1], out_shp[SERVER ? TILE * GW_MAX_LEARNERS : 1];
  __shared__ __align__(16) uint8_t out_end[SERVER ? TILE : 1];
  __shared__ __align__(16) int8_t act_s[SERVER ? TILE * GW_MAX_LEARNERS + 16 : 16];   // SERVER: the tile's learner actions, fetched once

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long tile_base = tile * TILE;
    const int tile_envs = (int)min((long long)TILE, p.E - tile_base);
    const int el = tid >> 3;
    const long long e = tile_base + el;
    const bool own = el < tile_envs;
    if (SERVER && tile != (long long)blockIdx.x) {                      // the previous tile's result copies have read the staging
      if (tid == 0) bulk_wait_read<0>();
      main_sync<SPLIT>();
    }

    // ================================================================= loads and RNG (no table needed)
    uint4 st = make_uint4(0, 0, 0, 0);
    uint32_t la = 0, npc_a = 0, chosen = 0;
    uint32_t rw[4] = {0, 0, 0, 0}, rw2[4] = {0, 0, 0, 0};
    bool act_staged = false;
    // ROLL (gw_rollout): the tile is stepped n_steps times by this CTA; state and random words stay in registers between
    // the steps, the outputs of step k go to slot (first_slot + k) of the caller's time-major rings (observation, obs_code
    // and action mask -- what the policy reads next -- to the slot after it), the actions come from slot (first_action + k).
    const int n_steps = ROLL ? rp->steps : 1;
    long long slot_t = ROLL ? rp->first_slot : 0, slot_a = ROLL ? rp->first_action : 0;
    uint32_t la_next = 0;
    uint4 st_carry = make_uint4(0, 0, 0, 0);
    for (int k_step = 0; k_step < n_steps; ++k_step) {
    const long long slot_o = (ROLL && slot_t + 1 >= rp->ring_slots) ? 0 : slot_t + 1;
    const long long et = ROLL ? slot_t * p.E + e : e;                  // index into the transition arrays (rewards, flags ...)
    const long long eo = ROLL ? slot_o * p.E + e : e;                  // ... the observation-side arrays
    const long long ea = ROLL ? slot_a * p.E + e : e;                  // ... the action arrays
    const bool first_step = !ROLL || k_step == 0, last_step = !ROLL || k_step + 1 == n_steps;
    la = 0; npc_a = 0;
    if (SERVER) {
      // The actions come out of pinned HOST memory: fetched lane by lane, the tile's 64 bytes turn into 16 read requests
      // per CTA and the PCIe read queue becomes the step's critical path (measured: CTAs finish 5-8 us apart).  A few
      // lanes fetch them in 16-byte pieces instead; the state loads below overlap the round trip.
      const int8_t* ap = p.io.learner_actions + tile_base * nl;
      act_staged = (reinterpret_cast<uintptr_t>(ap) & 15u) == 0;       // CTA-uniform
      if (act_staged && tid < (tile_envs * nl + 15) / 16)
        reinterpret_cast<uint4*>(act_s)[tid] = __ldcg(reinterpret_cast<const uint4*>(ap) + tid);   // stays inside the last 16-byte granule
    }
    if (own && first_step) {
      st = p.state[e];
      // The random words of this step were drawn at the end of the previous one (they only depend on seed, env id and
      // tick) and parked next to the state; the tag tells whether they belong to this tick (not after reset / set_state).
      uint4 c0 = make_uint4(0, 0, 0, 0), c1 = make_uint4(0, 0, 0, ~st.z);
      if (p.rng_cache != nullptr) { c0 = p.rng_cache[2 * e]; c1 = p.rng_cache[2 * e + 1]; }
      if (c1.w == st.z) {
        rw[0] = c0.x; rw[1] = c0.y; rw[2] = c0.z; rw[3] = c0.w;
        rw2[0] = c1.x; rw2[1] = c1.y;
        chosen = c1.z;
      } else {
        draw_step_randoms(p, e, st.z, n - nl > 2, rw, rw2, chosen);
      }
    }
    if (own) {
      if (!ROLL || first_step) {
#pragma unroll
        for (int k = 0; k < GW_MAX_LEARNERS; ++k)
          if (k < nl && !act_staged) la |= (uint32_t)min(max((int)(SERVER ? __ldcg(p.io.learner_actions + ea * nl + k) : p.io.learner_actions[ea * nl + k]), 0), 8) << (4 * k);
      } else {
        la = la_next;                                                  // fetched while the previous step ran
      }
      if (ROLL && !last_step) {                                        // the next step's actions: in flight during this step
        const long long slot_a1 = (slot_a + 1 >= rp->action_slots) ? 0 : slot_a + 1;
        const long long ea1 = slot_a1 * p.E + e;
        la_next = 0;
#pragma unroll
        for (int k = 0; k < GW_MAX_LEARNERS; ++k)
          if (k < nl) la_next |= (uint32_t)min(max((int)p.io.learner_actions[ea1 * nl + k], 0), 8) << (4 * k);
      }
      if (p.io.npc_actions != nullptr && r < 4 && r >= nl && r < n)
        npc_a = (uint32_t)min(max((int)(SERVER ? __ldcg(p.io.npc_actions + ea * n + r) : p.io.npc_actions[ea * n + r]), 0), 8);
    }
    if (SERVER && act_staged) {
      main_sync<SPLIT>();
#pragma unroll
      for (int k = 0; k < GW_MAX_LEARNERS; ++k)
        if (k < nl && own) la |= (uint32_t)min(max((int)act_s[el * nl + k], 0), 8) << (4 * k);
    }
    trace_stamp(p, 5);
    if (tables_pending) {
      tables_wait(s);                                                  // tables + the warps' staging rows have landed
      trace_stamp(p, 14);
      tables_pending = false;
    }
    trace_stamp(p, 1);

    uint32_t enc0 = 0xFFFFu, enc1 = 0xFFFFu;                           // the two special cells this lane patches into its env's row
    uint32_t v_end = 0, v_len = 0, v_cr = 0, v_ap = 0, v_tasks = 0, v_nz = 0;     // statistics: lane r == 0 of every group
    int v_ret = 0;
    double v_fear = 0.0;
    {                                        // groups without an env (tile tail) run along on zeros; their stores are masked
      const uint32_t cells = st.x, tick = st.z;
      uint32_t meta = st.y;
      // ---- agent role (setup_step :432-452, DefineActions, grid_world.py:481-518)
      const int ai = r & 3;
      const uint32_t c_r = (cells >> (8 * ai)) & 0xFFu;
      const uint32_t mdr_r = (ai < n) ? (uint32_t)s.small.mdr_map[c_r] : 0u;                // :445-447
      uint32_t a_r = 0;
      if (r >= 4) {
        if (r < 4 + nl) a_r = mdr_r;                                   // lanes 4, 5: learner r-4 plays its MdR
      } else if (ai < nl) {
        a_r = (la >> (4 * ai)) & 0xFu;                                 // :239-242
      } else if (ai < n) {
        if (p.io.npc_actions != nullptr) {
          a_r = npc_a;
        } else {
          const int m = ai - nl;
          const uint32_t wa = m == 0 ? rw[0] : (m == 1 ? rw[2] : rw2[0]);
          const uint32_t wb = m == 0 ? rw[1] : (m == 1 ? rw[3] : rw2[1]);
          const int pert = wa < p.perturb_thr ? 1 : 0;                 // random.random() < 0.25 (:441)
          const uint4* thr4 = reinterpret_cast<const uint4*>(s.small.policy_thr[s.small.policy_map[c_r]][pert]);
          const uint4 t0 = thr4[0], t1 = thr4[1];
          const uint32_t u = wb >> 1;
          a_r = (u >= t0.x) + (u >= t0.y) + (u >= t0.z) + (u >= t0.w) + (u >= t1.x) + (u >= t1.y) + (u >= t1.z) +
                (u >= t1.w);                                           // np.random.choice(9, p) (custom_agent.py:31)
        }
      }
      const Traj t = make_traj(s.sim.next, c_r, a_r);
      // bits 0-15: the trajectories played, 16-23: the learners' MdR trajectories
      const uint32_t effw = gather_lanes<6, 4>(FULL, t.eff, gsh);
      // learners whose action differs from their MdR (lanes 4, 5 hold the MdR as their action)
      const uint32_t neqb = (__ballot_sync(FULL, r >= 4 && r < 4 + nl && a_r != ((la >> (4 * (r - 4))) & 0xFu)) >> (gsh + 4)) & 3u;

      trace_stamp(p, 8);
      // ---- pair role: lane r < 6 owns pair (01,02,03,12,13,23)[r]
      const int pi = (0x211000 >> (4 * r)) & 0xF, pj = (0x332321 >> (4 * r)) & 0xF;
      bool near = false;
      uint32_t didx = 0;
      if (r < 6 && pj < n) {
        const uint32_t a = (cells >> (8 * pi)) & 0xFFu, b = (cells >> (8 * pj)) & 0xFFu;
        const int dr = (int)(b >> 4) - (int)(a >> 4), dc = (int)(b & 15) - (int)(a & 15);
        if (abs(dr) + abs(dc) <= 4) {
          near = true;
          didx = s.sim.diamond[(dr + 4) * 9 + (dc + 4)];
        }
      }
      const uint32_t near6 = (__ballot_sync(FULL, near) >> gsh) & 0x3Fu;
      if (SPLIT) {                                                     // everything FeAR needs is known: hand it to the helper warp
        box->in[warp][0][lane] = cells;
        box->in[warp][1][lane] = effw;
        box->in[warp][2][lane] = neqb | (near6 << 2) | (own ? 0x100u : 0u);
        box->in[warp][3][lane] = didx;
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&box->full[warp]));
      }
      uint32_t cm = 0;
      {
        uint32_t nn = 0, nr = 0, rn = 0;
        if (near && n >= 2) {
          const uint32_t base = didx * (N_EFF * N_EFF), ei = (effw >> (4 * pi)) & 0xFu, ej = (effw >> (4 * pj)) & 0xFu;
          nn = s.sim.lut[base + ei * N_EFF + ej];
          nr = s.sim.lut[base + ei * N_EFF];
          rn = s.sim.lut[base + ej];
        }
        if (__any_sync(FULL, nn != 0u)) {                              // (warp-uniform) somebody collides while everyone is on course
          const uint32_t NNw = gather_lanes<6, 4>(FULL, nn, gsh), NRw = gather_lanes<6, 4>(FULL, nr, gsh),
                         RNw = gather_lanes<6, 4>(FULL, rn, gsh);
          if (NNw) cm = fixpoint_words(NNw, NRw, RNw);
        }
      }
      const uint32_t crash = (cm >> 12) & 0xFu;
      trace_stamp(p, 9);

      // ---- agent role again: final cell, restricted flag, own apple (grid_world.py:531-563)
      uint32_t fin = 0, caught_r = 0;
      bool restr_r = false;
      if (r < 4) {
        fin = (n >= 2) ? cell_at(c_r, t, (crash >> r) & 1u, 3) : c_r;
        restr_r = ai < n && (t.r1 | (t.r2 & ~(cm >> r) & 1u)) != 0;
        if (r < nl && ((meta >> r) & 1u)) {
          const uint32_t apple = (p.apple_cells >> (8 * r)) & 0xFFu;
#pragma unroll
          for (int ss = 0; ss < 4; ++ss) {
            const uint32_t cur = (n >= 2) ? cell_at(c_r, t, (cm >> (4 * ss + r)) & 1u, ss) : c_r;
            caught_r += (cur == apple) ? 1u : 0u;
          }
        }
      }
      const uint32_t cells_new = gather_lanes<4, 8>(FULL, fin, gsh);
      const uint32_t restr = (__ballot_sync(FULL, restr_r) >> gsh) & 0xFu;
      // own-apple catches: only "none / exactly one / more" matters (ma_customenv.py:264, customenv.py:143)
      const uint32_t c_any = (__ballot_sync(FULL, caught_r > 0u) >> gsh) & 3u, c_one = (__ballot_sync(FULL, caught_r == 1u) >> gsh) & 3u;
      const uint32_t caught0 = (c_one & 1u) ? 1u : ((c_any & 1u) ? 2u : 0u), caught1 = (c_one & 2u) ? 1u : ((c_any & 2u) ? 2u : 0u);

      trace_stamp(p, 10);
      // ---- rewards and flags (the same for every lane of the group)
      const RewardOut ro = env_rewards(p, nl, meta, cells_new, crash, caught0, caught1);
      meta = ro.meta;
      const uint32_t steps_now = (meta >> M_STEPS_SH) & M_STEPS_MASK;
      const bool episode_over = (p.kind == GW_ENV_MULTI) ? (ro.trunc_now != 0) : ((ro.term_now | ro.trunc_now) != 0);
      const bool ended = episode_over || (p.max_steps > 0 && (int)steps_now >= p.max_steps);
      int ret0 = (int)(short)(st.w & 0xFFFFu), ret1 = (int)(short)(st.w >> 16);
      const double unit = (p.kind == GW_ENV_MULTI) ? 1.0 : 10.0;
      ret0 += (int)lrint(ro.reward[0] * unit);
      ret1 += (int)lrint(ro.reward[1] * unit);
      uint32_t cells_r = cells_new, apples_r = ro.apples_left, rflags = 0;
      uint4 st_out = make_uint4(cells_new, meta, tick + 1, ((uint32_t)ret0 & 0xFFFFu) | ((uint32_t)ret1 << 16));
      if (ended && p.auto_reset && own) {
        cells_r = (p.io.spawn != nullptr) ? spawn_cells(p, s.small, ea, tick) : cells_from_chosen(p, s.small, chosen);
        const uint32_t meta_sp = fresh_meta(p, cells_r);
        apples_r = meta_sp & M_APPLES;
        rflags = R_FRESH | R_FINAL | (ro.apples_left << 4);
        st_out = make_uint4(cells_r, meta_sp, tick + 1, 0u);
      }

      trace_stamp(p, 11);
      // ---- outputs: one lane per array
      const double my_reward = (r == 0) ? ro.reward[0] : ro.reward[1];
      st_carry = st_out;
      if (r == 0 && own) {
        if (last_step) p.state[e] = st_out;
        s.rinfo[el] = apples_r | rflags;
        s.cells_fin[el] = cells_new;
        v_end = ended ? 1u : 0u; v_len = ended ? steps_now : 0u; v_cr = ro.crash_count; v_ap = ro.apples_rewarded;
        v_ret = ended ? (ret0 + ret1) : 0;
      }
      if (r < nl && own && p.io.reward) {
        if (SERVER) out_rew[el * nl + r] = (float)my_reward;
        else p.io.reward[et * nl + r] = (float)my_reward;
      }
      if (r == 1 && own) write_positions(p.io.positions, et, n, cells_new);
      if (r == 2 && own && p.io.info)
        p.io.info[et] = (crash & 15u) | ((restr & 15u) << 4) | (ro.crash_count << 8) | (ro.apples_rewarded << 10) |
                       ((ended ? 1u : 0u) << 12) | (ro.shaped << 14);
      if (r == 3 && own && p.io.ended) {
        if (SERVER) out_end[el] = ended ? 1 : 0;
        else p.io.ended[et] = ended ? 1 : 0;
      }
      if (r >= 4 && r < 4 + nl && own) {
        if (p.io.terminated) p.io.terminated[et * nl + (r - 4)] = (uint8_t)((ro.term_now >> (r - 4)) & 1u);
        if (p.io.truncated) p.io.truncated[et * nl + (r - 4)] = (uint8_t)((ro.trunc_now >> (r - 4)) & 1u);
      }
      if (r == 6 && own && p.io.obs_code)
        p.io.obs_code[eo] = (unsigned long long)cells_r | ((unsigned long long)apples_r << 32) |
                           ((rflags & R_FRESH) ? (1ull << 34) : 0ull);

      trace_stamp(p, 12);
      // ---- observation specials and action masks of the (new) positions
      if (own) {
        const bool fresh = (rflags & R_FRESH) != 0;
        enc0 = special_entry(r, cpo, n, nl, p.kind, cells_r, apples_r, p.apple_cells, fresh);
        if (r < N_SPEC - 8) enc1 = special_entry(r + 8, cpo, n, nl, p.kind, cells_r, apples_r, p.apple_cells, fresh);
      }
      if (p.io.action_mask != nullptr) {                               // get_action_mask :467-506: lane r tests action r + 1
        const int a = r + 1, len = a >= 5 ? 2 : 1, d = (a - 1) & 3;
        bool ok[GW_MAX_LEARNERS] = {false, false};
#pragma unroll
        for (int k = 0; k < GW_MAX_LEARNERS; ++k) {
          if (k >= nl) break;
          const uint32_t pc = (cells_r >> (8 * k)) & 0xFFu;
          const int tr = (int)(pc >> 4) + (d == 0 ? -len : d == 1 ? len : 0), tc = (int)(pc & 15) + (d == 2 ? -len : d == 3 ? len : 0);
          ok[k] = (unsigned)tr < (unsigned)p.H && (unsigned)tc < (unsigned)GW_W && ((s.small.map_rows[tr] >> tc) & 1);
        }
        const uint32_t m0 = 1u | (((__ballot_sync(FULL, ok[0]) >> gsh) & 0xFFu) << 1);
        const uint32_t m1 = 1u | (((__ballot_sync(FULL, ok[1]) >> gsh) & 0xFFu) << 1);
        const uint32_t bits = m0 | (m1 << GW_N_ACTIONS);
        int8_t* dst = p.io.action_mask + eo * (long long)(nl * GW_N_ACTIONS);
        if (own)
          for (int b = r; b < nl * GW_N_ACTIONS; b += 8) dst[b] = (int8_t)((bits >> b) & 1u);
      }

      trace_stamp(p, 13);
      // ---- FeAR (ma_customenv.py:245-252, Responsibility.py:135-210): the env's own 8 lanes count its tasks
      double fear0 = 0.0, fear1 = 0.0;
      if (FEAR) {
        FearOut fo;
        if (SPLIT) {                                                   // computed by this warp's helper warp meanwhile (gw_rollout_split_kernel)
          mbar_wait(smem_u32(&box->done[warp]), split_phase);
          split_phase ^= 1u;
          fo.f0 = box->fear[warp][lane >> 3][0]; fo.f1 = box->fear[warp][lane >> 3][1]; fo.tasks = box->tasks[warp][lane >> 3];
          if (RS) {                                                    // the helper is free again: the step's output stage goes to it
            const int n_here = max(0, min(4, tile_envs - warp * 4));
            const long long obs_base = (ROLL ? slot_o * p.E : 0) + tile_base, fin_base = (ROLL ? slot_t * p.E : 0) + tile_base;
            uint32_t (*m)[32] = box->rin[warp];
            m[0][lane] = cells_r;
            m[1][lane] = apples_r | (rflags << 8) | (own ? 0x10000u : 0u) | ((uint32_t)n_here << 20) | ((uint32_t)tile_envs << 24);
            m[2][lane] = enc0; m[3][lane] = enc1;                      // this lane's special cells of its env's rows
            m[4][lane] = (uint32_t)obs_base; m[5][lane] = (uint32_t)((unsigned long long)obs_base >> 32);
            m[6][lane] = (uint32_t)fin_base; m[7][lane] = (uint32_t)((unsigned long long)fin_base >> 32);
            m[8][lane] = cells_new | 0u;
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&box->full[warp]));
          }
        } else {
          fo = fear_block(p, s, lane, gsh, r, n, nl, own, cells, effw, neqb, near6, didx);
        }
        fear0 = fo.f0; fear1 = fo.f1;
        if (r == 0) { v_tasks = fo.tasks; v_nz = (fear0 != 0.0) + (fear1 != 0.0); v_fear = fear0 + fear1; }
      }
      if (r < nl && own) {
        const double f = (r == 0) ? fear0 : fear1;
        if (p.io.fear) p.io.fear[et * nl + r] = f;
        if (p.io.shaped_reward) {
          const float sh = (float)(p.fear_weight * f + my_reward);                                        // maddpg/agent.py:130
          if (SERVER) out_shp[el * nl + r] = sh;
          else p.io.shaped_reward[et * nl + r] = sh;
        }
      }
    }
    trace_stamp(p, 2);
    if (SERVER) {
      fence_proxy_async_smem();
      main_sync<SPLIT>();
      if (warp == 0) {
        const int nr = tile_envs * nl;
        if (p.io.reward) server_flush(p.io.reward + tile_base * nl, out_rew, nr * 4, lane);
        if (p.io.shaped_reward) server_flush(p.io.shaped_reward + tile_base * nl, out_shp, nr * 4, lane);
        if (p.io.ended) server_flush(p.io.ended + tile_base, out_end, tile_envs, lane);
      }
    }

    // ================================================================= observations of the warp's own four envs
    // Every lane drops its special cells into its env's staging row, the four rows (contiguous, like the four
    // observations in HBM) leave as 128-bit streaming stores, and the cells are set back to the template's 0.
    if (!SERVER && !ROLL && !p.pdl_early && tile + gridDim.x >= n_tiles) asm volatile("griddepcontrol.launch_dependents;");
    if (RS) {                                                          // rows and masks are the helper's; the next step's random words stay here
      if ((p.rng_cache != nullptr || (ROLL && !last_step)) && own) {
        uint32_t nrw[4] = {0, 0, 0, 0}, nrw2[4] = {0, 0, 0, 0}, nchosen = 0;
        draw_step_randoms(p, e, st.z + 1, n - nl > 2, nrw, nrw2, nchosen);
        if (r == 0 && last_step && p.rng_cache != nullptr) {
          p.rng_cache[2 * e] = make_uint4(nrw[0], nrw[1], nrw[2], nrw[3]);
          p.rng_cache[2 * e + 1] = make_uint4(nrw2[0], nrw2[1], nchosen, st.z + 1);
        }
        if (ROLL) {
#pragma unroll
          for (int q = 0; q < 4; ++q) { rw[q] = nrw[q]; rw2[q] = nrw2[q]; }
          chosen = nchosen;
        }
      }
    }
    if (!RS && p.io.obs != nullptr) {
      patch_cell<OBS>(myrow, enc0, true);
      patch_cell<OBS>(myrow, enc1, true);
      fence_proxy_async_smem();
      __syncwarp();
      const int n_here = max(0, min(4, tile_envs - warp * 4));
      if (lane == 0 && n_here > 0)                         // four rows, contiguous here and in HBM: one TMA copy
        bulk_store_row(reinterpret_cast<uint8_t*>(p.io.obs) + ((ROLL ? slot_o * p.E : 0) + tile_base + warp * 4) * (long long)row_bytes, rows4, n_here * row_bytes);
      if ((p.rng_cache != nullptr || (ROLL && !last_step)) && own) {   // while the copy reads the rows: the next step's random words
        uint32_t nrw[4] = {0, 0, 0, 0}, nrw2[4] = {0, 0, 0, 0}, nchosen = 0;
        draw_step_randoms(p, e, st.z + 1, n - nl > 2, nrw, nrw2, nchosen);
        if (r == 0 && last_step && p.rng_cache != nullptr) {
          p.rng_cache[2 * e] = make_uint4(nrw[0], nrw[1], nrw[2], nrw[3]);
          p.rng_cache[2 * e + 1] = make_uint4(nrw2[0], nrw2[1], nchosen, st.z + 1);
        }
        if (ROLL) {
#pragma unroll
          for (int q = 0; q < 4; ++q) { rw[q] = nrw[q]; rw2[q] = nrw2[q]; }
          chosen = nchosen;
        }
      }
      if (lane == 0 && n_here > 0) {
        // a ring shorter than the launch is written twice by this thread: the first copy must have landed, not just been read
        if (ROLL && rp->wait_full) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        else bulk_wait_read<0>();
      }
      __syncwarp();
      patch_cell<OBS>(myrow, enc0, false);
      patch_cell<OBS>(myrow, enc1, false);
      __syncwarp();
    }
    if (!RS && p.io.final_obs != nullptr) {                // terminal observation of envs that were just re-spawned (rare)
      for (int el4 = warp * 4; el4 < min(tile_envs, warp * 4 + 4); ++el4) {
        const uint32_t ri = s.rinfo[el4];
        if (ri & R_FINAL)
          stage_and_store_env<OBS>(rows4, p.io.final_obs, (ROLL ? slot_t * p.E : 0) + tile_base + el4, p.H, n, nl, p.kind, s.cells_fin[el4], (ri >> 4) & 3u,
                                   p.apple_cells, false, lane);
      }
    }
    trace_stamp(p, 6);

    // ================================================================= statistics (most warps contribute nothing)
    {
      const unsigned any = __ballot_sync(0xFFFFFFFFu, (v_end | v_cr | v_ap | v_tasks | v_nz) != 0u);
      if (any) {
        const double unit = (p.kind == GW_ENV_MULTI) ? 1.0 : 10.0;
        const int slot = (int)(((tile_base >> 5) * 8 + warp) & (STAT_SLOTS - 1));
        const unsigned w_end = __reduce_add_sync(0xFFFFFFFFu, v_end), w_len = __reduce_add_sync(0xFFFFFFFFu, v_len);
        const unsigned w_cr = __reduce_add_sync(0xFFFFFFFFu, v_cr), w_ap = __reduce_add_sync(0xFFFFFFFFu, v_ap);
        const unsigned w_nz = __reduce_add_sync(0xFFFFFFFFu, v_nz), w_tk = __reduce_add_sync(0xFFFFFFFFu, v_tasks);
        const int w_ret = __reduce_add_sync(0xFFFFFFFFu, v_ret);
        if (lane == 0) {
          if (w_end) {
            atomicAdd(&p.stats[slot * 8 + ST_EPISODES], (unsigned long long)w_end);
            atomicAdd(&p.stats[slot * 8 + ST_LEN], (unsigned long long)w_len);
            atomicAdd(&p.stats[slot * 8 + ST_RETURN_MILLI], (unsigned long long)(long long)llrint(w_ret * (1000.0 / unit)));
          }
          if (w_cr) atomicAdd(&p.stats[slot * 8 + ST_CRASH], (unsigned long long)w_cr);
          if (w_ap) atomicAdd(&p.stats[slot * 8 + ST_APPLES], (unsigned long long)w_ap);
          if (w_nz) atomicAdd(&p.stats[slot * 8 + ST_FEAR_NZ], (unsigned long long)w_nz);
          if (w_tk) atomicAdd(&p.stats[slot * 8 + ST_TASKS], (unsigned long long)w_tk);
        }
        if (v_nz) atomicAdd(reinterpret_cast<double*>(&p.stats[slot * 8 + ST_FEAR_BITS]), v_fear);
      }
    }
    __syncwarp();                                                      // s.spec / s.rinfo of this warp are rewritten by the next tile
    trace_stamp(p, 7);
    if (ROLL) {
      st = st_carry;
      slot_t = slot_o;
      slot_a = (slot_a + 1 >= rp->action_slots) ? 0 : slot_a + 1;
    }
    }                                                                  // steps of this tile
  }
  if (SPLIT) *split_phase_io = split_phase;
}
template <bool FEAR, int OBS>
__global__ void __launch_bounds__(256, 2) gw_step_small_kernel(StepParams p) {
  constexpr int TILE = 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem<TILE>& s = *reinterpret_cast<Smem<TILE>*>(smem_raw);
  uint8_t* stage = smem_raw + smem_fixed_bytes<TILE>();
  trace_stamp(p, 0);
  load_tables<TILE>(s, stage, p, 32, p.nl * p.H * GW_W * (OBS == GW_OBS_F32 ? 4 : 2));
  trace_stamp(p, 3);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  bool tables_pending = true;
  small_step_tiles<FEAR, OBS, false>(p, s, stage, tables_pending);
  if (tables_pending) tables_wait(s);                                  // a CTA without tiles must not exit with copies in flight
}

// ------------------------------------------------------------------ device-side rollout: T steps per launch (gw_rollout)
// Envs are independent, so a CTA steps its 32 envs T times with no grid-wide synchronisation: tables and staging rows are
// loaded once, packed state and random words stay in registers, each step's actions are fetched while the previous step
// runs, and every step's outputs go to their own slot of the caller's time-major rings (the replay ring's layout).
template <bool FEAR, int OBS>
__global__ void __launch_bounds__(256, 2) gw_rollout_kernel(StepParams p, RollParams rp) {
  constexpr int TILE = 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem<TILE>& s = *reinterpret_cast<Smem<TILE>*>(smem_raw);
  uint8_t* stage = smem_raw + smem_fixed_bytes<TILE>();
  load_tables<TILE>(s, stage, p, 32, p.nl * p.H * GW_W * (OBS == GW_OBS_F32 ? 4 : 2));
  bool tables_pending = true;
  small_step_tiles<FEAR, OBS, false, true>(p, s, stage, tables_pending, &rp);
  if (tables_pending) tables_wait(s);
}

// The helper half of the *_split_kernel's: warp w + 8 serves warp w until the main warp posts the exit word.  Per step one
// FeAR hand-over each way and, with RS (every form but the host-driven one), the observation store: the main warp posts its
// lanes' special cells and the row indices once it has picked FeAR's values up, and this warp patches its staging rows,
// issues the TMA store, waits for the engine to have read them and restores the rows -- while the main warp is already on
// the next step.  (The staging rows are touched by this warp only.  Handing over the special cells and the action masks
// as well was measured on the same box and dropped: 4.52 us per env step against 4.37 with the store alone and 4.56 with FeAR
// alone, 4 096 envs, T = 64 -- the helper's chain, FeAR + the whole output stage, then bounds the step.)
template <int OBS, bool RS>
__device__ void fear_helper_loop(const StepParams& p, Smem<32>& s, uint8_t* stage, SplitBox* box, bool wait_full) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const int tid = (int)threadIdx.x - 256, warp = tid >> 5, lane = tid & 31, r = tid & 7, gsh = lane & 24;
  const int n = p.n, nl = p.nl, cpo = p.H * GW_W;
  const int Q = (OBS == GW_OBS_F32) ? cpo / 4 : cpo / 8, row_bytes = nl * Q * 16;
  uint8_t* const rows4 = stage + (size_t)warp * 4 * row_bytes;
  uint8_t* const myrow = rows4 + (lane >> 3) * row_bytes;
  tables_wait(s);
  uint32_t phase = 0;
  for (;;) {
    mbar_wait(smem_u32(&box->full[warp]), phase);
    phase ^= 1u;
    const uint32_t misc = box->in[warp][2][lane];
    if (misc & 0x200u) break;
    {
      const uint32_t cells = box->in[warp][0][lane], effw = box->in[warp][1][lane], didx = box->in[warp][3][lane];
      const FearOut fo = fear_block(p, s, lane, gsh, r, n, nl, (misc & 0x100u) != 0, cells, effw, misc & 3u, (misc >> 2) & 0x3Fu, didx);
      if (r == 0) {
        box->fear[warp][lane >> 3][0] = fo.f0;
        box->fear[warp][lane >> 3][1] = fo.f1;
        box->tasks[warp][lane >> 3] = fo.tasks;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&box->done[warp]));
    }
    if (!RS) continue;
    // ---- the step's output stage (the same code as in small_step_tiles, on the posted values)
    mbar_wait(smem_u32(&box->full[warp]), phase);
    phase ^= 1u;
    const uint32_t (*m)[32] = box->rin[warp];
    const uint32_t meta = m[1][lane], cells_fin = m[8][lane];
    const int n_here = (int)((meta >> 20) & 7u), tile_envs = (int)((meta >> 24) & 63u);
    const uint32_t enc0 = m[2][lane], enc1 = m[3][lane];
    const long long obs_base = (long long)(((unsigned long long)m[5][lane] << 32) | m[4][lane]);
    const long long fin_base = (long long)(((unsigned long long)m[7][lane] << 32) | m[6][lane]);
    if (p.io.obs != nullptr) {
      patch_cell<OBS>(myrow, enc0, true);
      patch_cell<OBS>(myrow, enc1, true);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && n_here > 0) {                       // four rows, contiguous here and in HBM: one TMA copy
        bulk_store_row(reinterpret_cast<uint8_t*>(p.io.obs) + (obs_base + warp * 4) * (long long)row_bytes, rows4, n_here * row_bytes);
        // a ring shorter than the launch is written twice by this thread: the first copy must have landed, not just been read
        if (wait_full) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        else bulk_wait_read<0>();
      }
      __syncwarp();
      patch_cell<OBS>(myrow, enc0, false);
      patch_cell<OBS>(myrow, enc1, false);
      __syncwarp();
    }
    if (p.io.final_obs != nullptr) {                       // terminal observation of envs that were just re-spawned (rare)
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t meta_q = __shfl_sync(FULL, meta, 8 * q4), fin_q = __shfl_sync(FULL, cells_fin, 8 * q4);
        const int el4 = warp * 4 + q4;
        if (el4 < tile_envs && ((meta_q >> 8) & R_FINAL))
          stage_and_store_env<OBS>(rows4, p.io.final_obs, fin_base + el4, p.H, n, nl, p.kind, fin_q, (meta_q >> 12) & 3u, p.apple_cells, false, lane);
      }
    }
  }
}
__device__ __forceinline__ void split_init(SplitBox* box) {          // before load_tables (which fences the inits and syncs the CTA)
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&box->full[i]), 1); mbar_init(smem_u32(&box->done[i]), 1); }
}

// gw_rollout with FeAR on sixteen warps: the step's dependent chain is what bounds small batches (a warp issues one
// instruction every ~6 cycles and a step is ~2 000 of them), and a quarter of it is FeAR, which needs nothing of the world
// update it sits behind.  Warps 0-7 run the step as in gw_rollout_kernel and post FeAR's inputs as soon as they are known;
// warps 8-15 compute it meanwhile; the step picks the values up where it used to compute them.
template <int OBS>
__global__ void __launch_bounds__(512, 1) gw_rollout_split_kernel(StepParams p, RollParams rp) {
  constexpr int TILE = 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem<TILE>& s = *reinterpret_cast<Smem<TILE>*>(smem_raw);
  uint8_t* stage = smem_raw + smem_fixed_bytes<TILE>();
  __shared__ SplitBox box;
  split_init(&box);
  load_tables<TILE>(s, stage, p, 32, p.nl * p.H * GW_W * (OBS == GW_OBS_F32 ? 4 : 2));
  if (threadIdx.x < 256) {
    bool tables_pending = true;
    uint32_t phase = 0;
    small_step_tiles<true, OBS, false, true, true>(p, s, stage, tables_pending, &rp, &box, &phase);
    if (tables_pending) tables_wait(s);
    split_post_exit(&box);
  } else {
    fear_helper_loop<OBS, true>(p, s, stage, &box, rp.wait_full != 0);
  }
}

// One gw_step launch per step, the same way (up to one tile per SM; the 256-thread kernel above runs two CTAs per SM).
template <int OBS>
__global__ void __launch_bounds__(512, 1) gw_step_small_split_kernel(StepParams p) {
  constexpr int TILE = 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem<TILE>& s = *reinterpret_cast<Smem<TILE>*>(smem_raw);
  uint8_t* stage = smem_raw + smem_fixed_bytes<TILE>();
  __shared__ SplitBox box;
  split_init(&box);
  load_tables<TILE>(s, stage, p, 32, p.nl * p.H * GW_W * (OBS == GW_OBS_F32 ? 4 : 2));
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x < 256) {
    bool tables_pending = true;
    uint32_t phase = 0;
    small_step_tiles<true, OBS, false, false, true>(p, s, stage, tables_pending, nullptr, &box, &phase);
    if (tables_pending) tables_wait(s);
    split_post_exit(&box);
  } else {
    fear_helper_loop<OBS, true>(p, s, stage, &box, false);
  }
}

// ------------------------------------------------------------------ resident step server (gw_step_host, mode 2)
// A host-driven step costs a kernel launch and a stream synchronisation on top of the kernel (about 20 of the 29 us per
// step at 4096 envs).  The resident kernel stays on the SMs between steps instead: tables and staging rows are loaded
// once, CTA 0 polls a doorbell word in pinned host memory and relays it through a word in device memory, every CTA
// steps its tiles (same code as gw_step_small_kernel), and the last CTA to finish writes the step's sequence number
// into pinned host memory, where the host spins on it.  Liveness: only CTA 0 decides to leave (STOP from the host, or
// no doorbell for idle_ns) and says so in device memory (for the other CTAs) and in host memory (for the host, which
// relaunches and rings again if its doorbell crossed the exit).  The grid is sized to be co-resident.
struct ServerParams {
  const unsigned long long* host_bell;   // pinned host: (seq << 32) | (op << 16) | buffer set
  unsigned int* host_resp;               // pinned host: [0] last completed seq, [1] generation of the launch that has left
  unsigned long long* dev_bell;          // device relay of the doorbell word
  unsigned int* dev_arrive;              // CTAs done, monotonic within a launch
  const gw_io* io_table;                 // device: the registered buffer sets
  unsigned int first_seq, generation;
  unsigned long long idle_ns;
};
enum { SRV_OP_STEP = 1, SRV_OP_STOP = 2, SRV_OP_EXIT = 3 };

__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <bool FEAR, int OBS, bool SPLIT = false>
__global__ void __launch_bounds__(SPLIT ? 512 : 256, SPLIT ? 1 : 2) gw_step_server_kernel(StepParams p, ServerParams sp) {
  constexpr int TILE = 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem<TILE>& s = *reinterpret_cast<Smem<TILE>*>(smem_raw);
  uint8_t* stage = smem_raw + smem_fixed_bytes<TILE>();
  __shared__ unsigned long long cmd_s;
  __shared__ __align__(16) unsigned long long io_s[(sizeof(gw_io) + 7) / 8];
  __shared__ typename std::conditional<SPLIT, SplitBox, int>::type box_s;     // the mailboxes exist in the split kernel only
  SplitBox* const box = reinterpret_cast<SplitBox*>(&box_s);
  uint32_t split_phase = 0;
  if (SPLIT) split_init(box);
  load_tables<TILE>(s, stage, p, 32, p.nl * p.H * GW_W * (OBS == GW_OBS_F32 ? 4 : 2));
  tables_wait(s);
  if (SPLIT && threadIdx.x >= 256) {                                   // FeAR helper warps: served through the mailboxes until told to leave
    fear_helper_loop<OBS, false>(p, s, stage, box, false);
    return;
  }
  unsigned int seq = sp.first_seq, rounds = 0;
#ifdef GW_ENABLE_TRACE
  unsigned long long tr_acc[5] = {0, 0, 0, 0, 0}, tr_t = 0;   // thread 0: time in wait / table entry / step / drain / arrive
#define SRV_TRACE(k) do { if (threadIdx.x == 0) { const unsigned long long n_ = global_ns(); tr_acc[k] += n_ - tr_t; tr_t = n_; } } while (0)
#else
#define SRV_TRACE(k) do { } while (0)
#endif
  for (;;) {
#ifdef GW_ENABLE_TRACE
    if (threadIdx.x == 0) tr_t = global_ns();
#endif
    if (threadIdx.x == 0) {
      unsigned long long w;
      const unsigned long long t0 = global_ns();
      if (blockIdx.x == 0) {
        for (;;) {
          w = ld_volatile_u64(sp.host_bell);
          if ((unsigned int)(w >> 32) == seq) break;
          if (global_ns() - t0 > sp.idle_ns) { w = ((unsigned long long)seq << 32) | ((unsigned long long)SRV_OP_EXIT << 16); break; }
        }
        asm volatile("fence.acq_rel.sys;" ::: "memory");               // what the host wrote before ringing is visible from here on
        st_volatile_u64(sp.dev_bell, w);
      } else {
        for (;;) {
          w = ld_volatile_u64(sp.dev_bell);
          const int ahead = (int)((unsigned int)(w >> 32) - seq);
          if (ahead == 0) break;
          // never taken while CTA 0 lives (it always relays an exit); a CTA that only became resident after the others
          // had left finds a later exit word, and the time bound keeps a fault elsewhere from turning into a hang
          if ((ahead > 0 && ((w >> 16) & 0xFFu) != SRV_OP_STEP) || global_ns() - t0 > 4ull * sp.idle_ns + 2000000000ull) {
            w = (unsigned long long)SRV_OP_EXIT << 16;
            break;
          }
        }
        // acquire at GPU scope: CTA 0 acquired the host's writes at system scope before it relayed the word (causality is
        // transitive), and system-scope fences issued by every CTA are served one after the other (~50 ns each, measured)
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      cmd_s = w;
    }
    main_sync<SPLIT>();
    const unsigned long long w = cmd_s;
    if (((w >> 16) & 0xFFu) != SRV_OP_STEP) break;
    SRV_TRACE(0);
    if (threadIdx.x < (int)((sizeof(gw_io) + 7) / 8))
      io_s[threadIdx.x] = __ldcv(reinterpret_cast<const unsigned long long*>(sp.io_table + (w & 0xFFFFu)) + threadIdx.x);
    main_sync<SPLIT>();
    StepParams q = p;
    q.io = *reinterpret_cast<const gw_io*>(io_s);
    bool tables_pending = false;
    SRV_TRACE(1);
    small_step_tiles<FEAR, OBS, true, false, SPLIT>(q, s, stage, tables_pending, nullptr, box, &split_phase);
    SRV_TRACE(2);
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // this thread's observation rows have landed
    main_sync<SPLIT>();
    SRV_TRACE(3);
    if (threadIdx.x == 0) {
      // Release at GPU scope per CTA (cumulative over what the barrier ordered before it), one system-scope fence by the
      // last CTA before the completion word: fences are cumulative, so it covers every CTA's stores into host memory.
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      const unsigned int prev = atomicAdd(sp.dev_arrive, 1u);
      if (prev == gridDim.x * (rounds + 1u) - 1u) {                    // last CTA of this step
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        st_volatile_u32(sp.host_resp, seq);
      }
    }
    SRV_TRACE(4);
    ++seq;
    ++rounds;
  }
#ifdef GW_ENABLE_TRACE
  if (threadIdx.x == 0 && p.trace != nullptr) {
    for (int k = 0; k < 5; ++k) p.trace[(size_t)blockIdx.x * 16 + k] = tr_acc[k];   // this launch only
    p.trace[(size_t)blockIdx.x * 16 + 5] = rounds;
  }
#endif
  if (SPLIT) split_post_exit(box);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    st_volatile_u32(sp.host_resp + 1, sp.generation);
  }
}

static inline size_t stage_row_bytes(const gw_config& c) {
  return (size_t)c.n_learners * c.height * GW_W * (c.obs_dtype == GW_OBS_F32 ? 4 : 2);
}

// ------------------------------------------------------------------ operator-level kernels
typedef SimTab OpSmem;

__device__ __forceinline__ void load_op_tables(OpSmem& s, const Tables* __restrict__ T) {
  const uint4* src = reinterpret_cast<const uint4*>(&T->sim);
  uint4* dst = reinterpret_cast<uint4*>(&s);
  for (int i = threadIdx.x; i < (int)sizeof(SimTab) / 16; i += blockDim.x) dst[i] = __ldg(src + i);
  __syncthreads();
}

__device__ __forceinline__ void load_case(long long c, int n, const int8_t* pos, const int8_t* act, uint32_t& cells,
                                          uint32_t& acts) {
  cells = acts = 0;
  for (int i = 0; i < n; ++i) {
    cells |= (uint32_t)(((pos[(c * 4 + i) * 2] & 15) << 4) | (pos[(c * 4 + i) * 2 + 1] & 15)) << (8 * i);
    acts |= (uint32_t)min(max((int)act[c * 4 + i], 0), 8) << (4 * i);
  }
}

__global__ void __launch_bounds__(128) gw_update_world_kernel(const Tables* T, int n_default, long long C,
                                                              const int8_t* n_per, const int8_t* pos, const int8_t* act,
                                                              const int8_t* apples, int8_t* new_pos, uint8_t* crash,
                                                              uint8_t* restr, int8_t* caught) {
  __shared__ OpSmem s;
  load_op_tables(s, T);
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int n = n_per ? n_per[c] : n_default;
  uint32_t cells, acts, apple_cells = 0, apple_on = 0;
  load_case(c, n, pos, act, cells, acts);
  if (apples)
    for (int k = 0; k < 2; ++k)
      if (apples[(c * 2 + k) * 2] >= 0) {
        apple_on |= 1u << k;
        apple_cells |= (uint32_t)(((apples[(c * 2 + k) * 2] & 15) << 4) | (apples[(c * 2 + k) * 2 + 1] & 15)) << (8 * k);
      }
  const PairGeom g = pair_geometry(s, n, cells);
  const StepResult r = world_update(s, n, cells, acts, g, apple_cells, apple_on, min(2, n));
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

// one warp per case: lanes 0..26 = (affected slot, affected action), two rounds (actor plays MdR / its action)
__global__ void __launch_bounds__(128) gw_fear_kernel(const Tables* T, int n_default, long long C, const int8_t* n_per,
                                                      const int8_t* pos, const int8_t* act, const int8_t* mdr,
                                                      const int8_t* actor, const uint8_t* in_list, double* resp,
                                                      int8_t* n_mdr, int8_t* n_act) {
  __shared__ OpSmem s;
  load_op_tables(s, T);
  const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const int n = n_per ? n_per[c] : n_default;
  const int x = actor[c];
  uint32_t cells, acts, lst = 1u << x;
  load_case(c, n, pos, act, cells, acts);
  for (int i = 0; i < n; ++i)
    if (in_list == nullptr || in_list[c * 4 + i]) lst |= 1u << i;
  const PairGeom g = pair_geometry(s, n, cells);
  uint32_t base = 0;                                   // effective trajectories of the listed agents; the others Stay (:43)
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n && ((lst >> k) & 1u))
      base |= make_traj(s.next, (cells >> (8 * k)) & 0xFFu, (acts >> (4 * k)) & 0xFu).eff << (4 * k);
  const int js = lane >> 1, v = lane & 1;                // lanes 0..5 = (affected slot, actor plays MdR / its action)
  const int j = js + (js >= x ? 1 : 0);
  uint32_t my = 0;
  if (lane < 6 && j < n) {
    const uint32_t av = v == 0 ? (uint32_t)min(max((int)mdr[c * 4 + x], 0), 8) : ((acts >> (4 * x)) & 0xFu);
    const uint32_t eo = (base & ~(0xFu << (4 * x))) | (make_traj(s.next, (cells >> (8 * x)) & 0xFFu, av).eff << (4 * x));
    my = count_valid_moves(s, cells, eo, g, j, ((lst >> j) & 1u) != 0);   // SwapActionIDs4Agents: listed agents only
  }
  uint32_t packed = 0;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    packed |= (uint32_t)__shfl_sync(0xFFFFFFFFu, my, 2 * q) << (4 * q);
    packed |= (uint32_t)__shfl_sync(0xFFFFFFFFu, my, 2 * q + 1) << (16 + 4 * q);
  }
  if (lane < 4) {
    double rv = 0.0;
    int8_t m = 0, a = 0;
    if (lane < n && lane != x) {
      const int q = lane - (lane > x ? 1 : 0);
      m = (int8_t)((packed >> (4 * q)) & 0xF);
      a = (int8_t)((packed >> (16 + 4 * q)) & 0xF);
      rv = T->small.resp_lut[m][a];
    }
    resp[c * 4 + lane] = rv;
    if (n_mdr) n_mdr[c * 4 + lane] = m;
    if (n_act) n_act[c * 4 + lane] = a;
  }
}

// Responsibility.FeAR for all actors (one warp per (case, actor); lanes 0..5 = (affected slot, variant)) and
// Responsibility.FeAL (one warp per case; lanes 0..7 = (agent, variant)).
__global__ void __launch_bounds__(128) gw_fear_matrix_kernel(const Tables* T, int n_default, long long C, const int8_t* n_per,
                                                             const int8_t* pos, const int8_t* act, const int8_t* mdr,
                                                             const uint8_t* in_list, double* resp, int8_t* n_mdr,
                                                             int8_t* n_act) {
  __shared__ OpSmem s;
  load_op_tables(s, T);
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long c = wid >> 2;
  const int x = (int)(wid & 3);
  if (c >= C) return;
  const int n = n_per ? n_per[c] : n_default;
  uint32_t cells, acts, lst = 0;
  load_case(c, n, pos, act, cells, acts);
  for (int i = 0; i < n; ++i)
    if (in_list == nullptr || in_list[c * 4 + i]) lst |= 1u << i;
  const PairGeom g = pair_geometry(s, n, cells);
  uint32_t base = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n && ((lst >> k) & 1u))
      base |= make_traj(s.next, (cells >> (8 * k)) & 0xFFu, (acts >> (4 * k)) & 0xFu).eff << (4 * k);
  const int js = lane >> 1, v = lane & 1;
  const int j = js + (js >= x ? 1 : 0);
  uint32_t my = 0;
  if (x < n && lane < 6 && j < n) {
    uint32_t eo = base;
    if (v == 0 && ((lst >> x) & 1u))                    // SwapActionIDs4Agents: only an actor present in the list plays its MdR
      eo = (base & ~(0xFu << (4 * x))) |
           (make_traj(s.next, (cells >> (8 * x)) & 0xFFu, (uint32_t)min(max((int)mdr[c * 4 + x], 0), 8)).eff << (4 * x));
    my = count_valid_moves(s, cells, eo, g, j, ((lst >> j) & 1u) != 0);
  }
  const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, my, 1);
  if (lane < 6 && (lane & 1) == 0) {                    // even lane holds n_mdr, its neighbour n_act
    const long long o = c * 16 + x * 4 + j;
    if (x < n && j < n) {
      resp[o] = T->small.resp_lut[my][other];
      n_mdr[o] = (int8_t)my;
      n_act[o] = (int8_t)other;
    }
  }
  if (lane < 4) {                                       // diagonal and padding entries
    const long long o = c * 16 + x * 4 + lane;
    if (lane == x || x >= n || lane >= n) { resp[o] = 0.0; n_mdr[o] = 0; n_act[o] = 0; }
  }
}

__global__ void __launch_bounds__(128) gw_feal_kernel(const Tables* T, int n_default, long long C, const int8_t* n_per,
                                                      const int8_t* pos, const int8_t* act, const int8_t* mdr,
                                                      const uint8_t* in_list, double* feal, int8_t* n_mdr, int8_t* n_act) {
  __shared__ OpSmem s;
  load_op_tables(s, T);
  const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const int n = n_per ? n_per[c] : n_default;
  uint32_t cells, acts, lst = 0;
  load_case(c, n, pos, act, cells, acts);
  for (int i = 0; i < n; ++i)
    if (in_list == nullptr || in_list[c * 4 + i]) lst |= 1u << i;
  const PairGeom g = pair_geometry(s, n, cells);
  const int ii = lane >> 1, v = lane & 1;               // v = 0: the others play their MdR, v = 1: their actions
  uint32_t my = 0;
  if (lane < 8 && ii < n) {
    uint32_t eo = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < n && k != ii && ((lst >> k) & 1u)) {
        const uint32_t a = v == 0 ? (uint32_t)min(max((int)mdr[c * 4 + k], 0), 8) : (acts >> (4 * k)) & 0xFu;
        eo |= make_traj(s.next, (cells >> (8 * k)) & 0xFFu, a).eff << (4 * k);
      }
    my = count_valid_moves(s, cells, eo, g, ii, ((lst >> ii) & 1u) != 0);
  }
  const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, my, 1);
  if (lane < 8 && (lane & 1) == 0) {
    const long long o = c * 4 + ii;
    if (ii < n) {
      const double r = __ddiv_rn((double)other, __dadd_rn((double)my, 0.000001));     // Responsibility.py:287-288
      feal[o] = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
      n_mdr[o] = (int8_t)my;
      n_act[o] = (int8_t)other;
    } else {
      feal[o] = 0.0; n_mdr[o] = 0; n_act[o] = 0;
    }
  }
}

}  // namespace gw

// ====================================================================== host side / C-ABI
#include "gw_internal.h"

static bool pick_small(long long E);

thread_local std::string g_gw_create_err;

int gw_fail(gw_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_gw_create_err = msg;
  return code;
}
int gw_cuda_fail(gw_handle* h, cudaError_t e, const char* what) {
  return gw_fail(h, GW_ECUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
static int fail(gw_handle* h, int code, const std::string& msg) { return gw_fail(h, code, msg); }
static int cuda_fail(gw_handle* h, cudaError_t e, const char* what) { return gw_cuda_fail(h, e, what); }

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

const char* gw_last_error(const gw_handle* h) { return h ? h->err.c_str() : g_gw_create_err.c_str(); }

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
    if (!((c->map_rows[r] >> col) & 1)) { why = "apple on an inactive cell"; return GW_EINVAL; }
  }
  if (c->env_kind == GW_ENV_SINGLE && c->apple_row[0] < 0) { why = "single env needs an apple"; return GW_EINVAL; }
  if (c->n_blocked < 0 || c->n_blocked > GW_MAX_BLOCKED) { why = "n_blocked must be 0..256"; return GW_EINVAL; }
  for (int k = 0; k < c->n_blocked; ++k) {
    const int fr = c->blocked_from[k] >> 4, fc = c->blocked_from[k] & 15, tr = c->blocked_to[k] >> 4, tc = c->blocked_to[k] & 15;
    if (fr >= c->height || tr >= c->height || std::abs(fr - tr) + std::abs(fc - tc) != 1) { why = "restricted path: cells must be adjacent and inside the grid"; return GW_EINVAL; }
  }
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

// the same thresholds for the general layout's translation unit (gw_wide.cu)
void gw_policy_thresholds(const float sw[3], const float dw_in[4], bool perturbed, uint32_t thr[8]) {
  policy_thresholds(sw, dw_in, perturbed, thr);
}

// next-cell table: one move of grid_world.py:481-518 (off-grid clip or inactive target => stay)
static void build_next_cell(const gw_config* c, uint8_t* next) {
  static const int DR[4] = {-1, 1, 0, 0}, DC[4] = {0, 0, -1, 1};     // Up, Down, Left, Right (custom_agent.py:140-150)
  for (int r = 0; r < GW_MAX_H; ++r)
    for (int col = 0; col < GW_W; ++col)
      for (int d = 0; d < 4; ++d) {
        const int tr = r + DR[d], tc = col + DC[d];
        const bool ok = tr >= 0 && tr < c->height && tc >= 0 && tc < c->width && ((c->map_rows[tr] >> tc) & 1);
        next[((r << 4) | col) * 4 + d] = (uint8_t)(ok ? ((tr << 4) | tc) : ((r << 4) | col));
      }
  // restricted paths (walls / one-ways, grid_world.py:493-509): the move is refused like one into an inactive cell
  for (int k = 0; k < c->n_blocked; ++k) {
    const int from = c->blocked_from[k], to = c->blocked_to[k];
    for (int d = 0; d < 4; ++d)
      if ((from >> 4) + DR[d] == (to >> 4) && (from & 15) + DC[d] == (to & 15)) next[from * 4 + d] = (uint8_t)from;
  }
}

// pair-mask table: the literal five-rule test (gw::pair_hit) for every relative start offset within Manhattan
// distance 4 and every pair of effective trajectories, bit s = collision at sub-step s (both agents on course).
static int diamond_index(int dr, int dc) {           // row-major rank of (dr, dc) among the offsets with |dr| + |dc| <= 4
  int idx = 0;
  for (int r = -4; r <= 4; ++r)
    for (int c = -4; c <= 4; ++c) {
      if (std::abs(r) + std::abs(c) > 4) continue;
      if (r == dr && c == dc) return idx;
      ++idx;
    }
  return -1;
}

static void build_pair_lut(uint8_t* lut) {
  static const int DR[4] = {-1, 1, 0, 0}, DC[4] = {0, 0, -1, 1};
  std::memset(lut, 0, gw::LUT_BYTES);
  auto traj = [&](int p0, int eff, int& p1, int& p2, int& len) {
    p1 = p2 = p0;
    len = 1;
    if (eff == 0) return;
    const int d = (eff - 1) & 3, step = DR[d] * 64 + DC[d];            // virtual 64-wide grid: no wrap-around
    p1 = p0 + step;
    if (eff <= 4) { p2 = p1; len = 1; }
    else if (eff <= 8) { p2 = p1 + step; len = 2; }
    else { p2 = p1; len = 2; }                                         // second move blocked
  };
  for (int dr = -4; dr <= 4; ++dr)
    for (int dc = -4; dc <= 4; ++dc) {
      if (std::abs(dr) + std::abs(dc) > 4 || (dr == 0 && dc == 0)) continue;
      const int Pi = 16 * 64 + 16, Pj = Pi + dr * 64 + dc;
      for (int ei = 0; ei < gw::N_EFF; ++ei)
        for (int ej = 0; ej < gw::N_EFF; ++ej) {
          int i1, i2, li, j1, j2, lj;
          traj(Pi, ei, i1, i2, li);
          traj(Pj, ej, j1, j2, lj);
          uint8_t mask = 0;
          for (int s = 0; s < 4; ++s) {
            const int qi = (s + 1) * li, fi = qi >> 2, ci = (qi + 3) >> 2;
            const int qj = (s + 1) * lj, fj = qj >> 2, cj = (qj + 3) >> 2;
            const int Ai = fi == 0 ? Pi : (fi == 1 ? i1 : i2), Bi = ci == 1 ? i1 : i2;
            const int Aj = fj == 0 ? Pj : (fj == 1 ? j1 : j2), Bj = cj == 1 ? j1 : j2;
            if (gw::pair_hit(Ai, Bi, Pi, qi, fi, ci, Aj, Bj, Pj, qj, fj, cj)) mask |= (uint8_t)(1u << s);
          }
          lut[diamond_index(dr, dc) * (gw::N_EFF * gw::N_EFF) + ei * gw::N_EFF + ej] = mask;
        }
    }
}

// The tables derived from the pair-mask table and the map.  Returns false if the pair table is not symmetric under
// swapping the two agents (count_valid_moves relies on it; it is a property of the five rules, checked here anyway).
static bool build_sim_tables(const gw_config* cfg, gw::SimTab* T) {
  build_pair_lut(T->lut);
  build_next_cell(cfg, T->next);
  const int EE = gw::N_EFF * gw::N_EFF;
  for (int d = 0; d < gw::N_DELTA; ++d)
    for (int a = 0; a < gw::N_EFF; ++a) {
      for (int ss = 0; ss < 4; ++ss) T->rowmask4[d * gw::N_EFF + a][ss] = 0;
      for (int b = 0; b < gw::N_EFF; ++b) {
        if (T->lut[d * EE + a * gw::N_EFF + b] != T->lut[(gw::N_DELTA - 1 - d) * EE + b * gw::N_EFF + a]) return false;
        for (int ss = 0; ss < 4; ++ss)
          if ((T->lut[d * EE + a * gw::N_EFF + b] >> ss) & 1) T->rowmask4[d * gw::N_EFF + a][ss] |= (uint16_t)(1u << b);
      }
    }
  for (int cell = 0; cell < GW_MAX_H * GW_W; ++cell) {     // actions that are not restricted (grid_world.py:481-518)
    uint16_t m = 1;
    for (int d = 0; d < 4; ++d) {
      const int p1 = T->next[cell * 4 + d];
      if (p1 == cell) continue;
      m |= (uint16_t)(1u << (1 + d));
      if (T->next[p1 * 4 + d] != p1) m |= (uint16_t)(1u << (5 + d));
    }
    T->unres[cell] = m;
  }
  std::memset(T->diamond, 0xFF, sizeof(T->diamond));
  for (int dr = -4; dr <= 4; ++dr)
    for (int dc = -4; dc <= 4; ++dc)
      if (std::abs(dr) + std::abs(dc) <= 4) T->diamond[(dr + 4) * 9 + (dc + 4)] = (uint8_t)diamond_index(dr, dc);
  for (int near6 = 0; near6 < 64; ++near6) {
    uint16_t w = 0;
    for (int x = 0; x < 4; ++x) w |= (uint16_t)(gw::reach_mask_slow((uint32_t)near6, x) << (4 * x));
    T->reach[near6] = w;
  }
  return true;
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
  h->sm_count = prop.multiProcessorCount;

  gw::Tables* t = new gw::Tables();
  std::memset(t, 0, sizeof(*t));
  std::memcpy(t->small.map_rows, cfg->map_rows, sizeof(t->small.map_rows));
  std::memcpy(t->small.mdr_map, cfg->mdr_map, sizeof(t->small.mdr_map));
  std::memcpy(t->small.policy_map, cfg->policy_map, sizeof(t->small.policy_map));
  for (int p = 0; p < cfg->n_policies; ++p) {
    policy_thresholds(cfg->step_weights[p], cfg->dir_weights[p], false, t->small.policy_thr[p][0]);
    policy_thresholds(cfg->step_weights[p], cfg->dir_weights[p], true, t->small.policy_thr[p][1]);
  }
  int na = 0;
  for (int r = 0; r < cfg->height; ++r)
    for (int c = 0; c < cfg->width; ++c)
      if ((cfg->map_rows[r] >> c) & 1) t->small.active_cell[na++] = (uint8_t)((r << 4) | c);
  t->n_active = na;
  h->n_active = na;
  if (!build_sim_tables(cfg, &t->sim)) { delete t; delete h; return fail(nullptr, GW_EINVAL, "gw_create: internal: pair table is not symmetric"); }
  // replicated observation rows: n_learners copies of the map template (-1 inactive / 0 active) per row
  const size_t row_bytes = (size_t)cfg->n_learners * cfg->height * GW_W * (cfg->obs_dtype == GW_OBS_F32 ? 4 : 2);
  std::string stage_init(row_bytes * gw::STAGE_ROWS, '\0');
  for (int r = 0; r < gw::STAGE_ROWS; ++r)
    for (int k = 0; k < cfg->n_learners; ++k)
      for (int cell = 0; cell < cfg->height * GW_W; ++cell) {
        const bool active = (cfg->map_rows[cell >> 4] >> (cell & 15)) & 1;
        const size_t idx = ((size_t)r * cfg->n_learners + k) * cfg->height * GW_W + cell;
        if (cfg->obs_dtype == GW_OBS_F32) reinterpret_cast<uint32_t*>(&stage_init[0])[idx] = active ? 0u : 0xBF800000u;   // 0.0f / -1.0f
        else reinterpret_cast<uint16_t*>(&stage_init[0])[idx] = active ? (uint16_t)0 : (uint16_t)0xBF80u;               // bf16 0 / -1
      }
  for (int m = 0; m < 10; ++m)
    for (int a = 0; a < 10; ++a) {
      volatile double v = ((double)m - (double)a) / ((double)m + 0.000001);   // Responsibility.py:194-195, EPS :12
      double cl = v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v);                      // np.clip :198
      t->small.resp_lut[m][a] = cl;
    }
  auto cleanup = [&](int rc) { delete t; gw_destroy(h); return rc; };
  if ((e = cudaMalloc(&h->d_tables, sizeof(gw::Tables))) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMalloc tables"));
  if ((e = cudaMalloc(&h->d_state, sizeof(uint4) * (size_t)cfg->num_envs)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMalloc state"));
  if ((e = cudaMalloc(&h->d_stats, sizeof(unsigned long long) * gw::STAT_SLOTS * 8)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMalloc stats"));
  if ((e = cudaMemcpy(h->d_tables, t, sizeof(gw::Tables), cudaMemcpyHostToDevice)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMemcpy tables"));
  if (pick_small(cfg->num_envs)) {               // tag ~tick never matches: the first step draws its own words
    if ((e = cudaMalloc(&h->d_rng_cache, 2 * sizeof(uint4) * (size_t)cfg->num_envs)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMalloc rng cache"));
    if ((e = cudaMemset(h->d_rng_cache, 0xA5, 2 * sizeof(uint4) * (size_t)cfg->num_envs)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMemset rng cache"));
  }
  if ((e = cudaMalloc(&h->d_tile_ctr, 2 * sizeof(unsigned int))) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMalloc tile counter"));
  if ((e = cudaMemset(h->d_tile_ctr, 0, 2 * sizeof(unsigned int))) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMemset tile counter"));
  if ((e = cudaMalloc(&h->d_stage_init, stage_init.size())) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMalloc stage rows"));
  if ((e = cudaMemcpy(h->d_stage_init, stage_init.data(), stage_init.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMemcpy stage rows"));
  if ((e = cudaMemset(h->d_state, 0, sizeof(uint4) * (size_t)cfg->num_envs)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMemset state"));
  if ((e = cudaMemset(h->d_stats, 0, sizeof(unsigned long long) * gw::STAT_SLOTS * 8)) != cudaSuccess) return cleanup(cuda_fail(nullptr, e, "cudaMemset stats"));
  delete t;
  if (const char* tr = std::getenv("GW_TRACE")) {
    if (std::atoi(tr) != 0 && cudaMalloc(&h->d_trace, sizeof(unsigned long long) * 16 * 4096) == cudaSuccess)
      cudaMemset(h->d_trace, 0, sizeof(unsigned long long) * 16 * 4096);
  }
  *out = h;
  return GW_OK;
}

static int server_stop(gw_handle* h);
static void server_free(gw_handle* h);

int gw_destroy(gw_handle* h) {
  if (!h) return GW_OK;
  cudaSetDevice(h->cfg.device);
  server_stop(h);
  server_free(h);
  if (h->d_tables) cudaFree(h->d_tables);
  if (h->d_stage_init) cudaFree(h->d_stage_init);
  if (h->d_tile_ctr) cudaFree(h->d_tile_ctr);
  if (h->d_rng_cache) cudaFree(h->d_rng_cache);
  if (h->d_state) cudaFree(h->d_state);
  if (h->d_stats) cudaFree(h->d_stats);
  if (h->d_trace) cudaFree(h->d_trace);
  delete h;
  return GW_OK;
}

static gw::StepParams make_params(gw_handle* h, const gw_io* io) {
  gw::StepParams p;
  std::memset(&p, 0, sizeof(p));
  const gw_config& c = h->cfg;
  p.tables = h->d_tables;
  p.stage_init = h->d_stage_init;
  p.tile_ctr = h->d_tile_ctr;
  p.rng_cache = h->d_rng_cache;
  static const int dyn_env = [] { const char* v = std::getenv("GW_DYN"); return v ? std::atoi(v) : 1; }();
  p.dyn_tiles = dyn_env;
  p.state = h->d_state;
  p.stats = h->d_stats;
  if (io) p.io = *io;
  p.E = c.num_envs;
  p.env_id_base = c.env_id_base;
  p.n = c.n_agents; p.nl = c.n_learners; p.kind = c.env_kind; p.H = c.height;
  p.fear_radius = c.fear_radius; p.max_steps = c.max_steps; p.auto_reset = c.auto_reset;
  p.n_active = h->n_active;
  for (int k = 0; k < c.n_learners; ++k)
    if (c.apple_row[k] >= 0) {
      p.apple_cells |= (uint32_t)((c.apple_row[k] << 4) | c.apple_col[k]) << (8 * k);
      p.apple_init |= 1u << k;
    }
  double thr = c.perturb_prob * 4294967296.0;
  p.perturb_thr = thr >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)thr;
  p.seed_lo = (uint32_t)c.seed; p.seed_hi = (uint32_t)(c.seed >> 32);
  p.fear_weight = c.fear_weight;
  p.trace = h->d_trace;
  return p;
}

static inline cudaError_t use_device(int device) {
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == device) return cudaSuccess;
  return cudaSetDevice(device);
}

static bool misaligned(const void* p, size_t a) { return p && (reinterpret_cast<uintptr_t>(p) % a) != 0; }

static int check_io(gw_handle* h, const gw_io* io, bool step) {
  if (!io) return fail(h, GW_EINVAL, "null gw_io");
  if (step && !io->learner_actions) return fail(h, GW_EINVAL, "gw_step: learner_actions is required");
  if (misaligned(io->obs, 16) || misaligned(io->final_obs, 16)) return fail(h, GW_EINVAL, "obs/final_obs must be 16-byte aligned");
  if (misaligned(io->reward, 8) || misaligned(io->shaped_reward, 8) || misaligned(io->fear, 16) || misaligned(io->info, 4) ||
      misaligned(io->positions, 8) || misaligned(io->terminated, 2) || misaligned(io->truncated, 2))
    return fail(h, GW_EINVAL, "misaligned output pointer (reward/shaped_reward/positions 8 B, fear 16 B, info 4 B)");
  return GW_OK;
}

}  // extern "C" (templates below need C++ linkage)

// Launch shape (measured on B200, profiles/README.md).  Up to ~6k envs the step is latency-bound and the 8-lanes-per-env
// kernel wins (32-env CTAs: E = 4096 -> 128 CTAs, one per SM); up to ~24k envs thread-per-env with 32-env tiles still
// puts work on every SM; beyond that 128- and (from ~200k envs) 256-env tiles amortise the per-CTA table copy and the
// kernel runs at the memory system's pace.  GW_TILE=32|128|256 and GW_SMALL=0|1 override (dev).
static int pick_tile(long long E) {
  if (const char* s = std::getenv("GW_TILE")) {
    const int v = std::atoi(s);
    if (v == 32 || v == 128 || v == 256) return v;
  }
  return E <= 24576 ? 32 : (E <= 196608 ? 128 : 256);
}

// Resident step kernel (gw_step_host mode 2): the 8-lanes-per-env body walking several tiles per CTA still beats a launch
// per step well past the small-batch threshold (measured, profiles/README.md); GW_RESIDENT_MAX overrides (dev).
static bool pick_resident(long long E) {
  static const long long lim = [] { const char* v = std::getenv("GW_RESIDENT_MAX"); return v ? std::atoll(v) : 24576ll; }();
  return E <= lim;
}

static bool pick_small(long long E) {
  if (const char* s = std::getenv("GW_SMALL")) return std::atoi(s) != 0;
  return E <= 6144;
}

template <typename K>
static void launch_k(K kernel, const gw::StepParams& p, unsigned blocks, int threads, size_t smem, cudaStream_t s,
                     bool pdl = false) {
  // > 48 KB of dynamic shared memory needs the opt-in; raised (never lowered) per (device, function)
  {
    struct Seen { int dev; const void* fn; size_t smem; };
    static thread_local Seen seen[64];
    static thread_local int n_seen = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    int at = -1;
    for (int i = 0; i < n_seen; ++i)
      if (seen[i].dev == dev && seen[i].fn == (const void*)kernel) at = i;
    if (at < 0 || seen[at].smem < smem) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (at >= 0) seen[at].smem = smem;
      else if (n_seen < 64) seen[n_seen++] = Seen{dev, (const void*)kernel, smem};
    }
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  gw::StepParams pp = p;
  cudaLaunchKernelEx(&cfg, kernel, pp);
}

// Programmatic dependent launch is opt-in (GW_PDL=1): measured on B200 it saves 0.4 us per step at 4096 envs but costs
// 7 us at 65536 envs, where the early-launched CTAs of the next grid unbalance the SMs (profiles/README.md).
// FeAR on helper warps (the *_split_kernel's): on unless GW_SPLIT=0; they hold one 512-thread CTA per SM
static bool use_split() {
  static const bool on = [] { const char* e = std::getenv("GW_SPLIT"); return !(e && e[0] == '0'); }();
  return on;
}

static bool use_pdl() {
  static const bool v = [] { const char* s = std::getenv("GW_PDL"); return s && std::atoi(s) != 0; }();
  return v;
}

template <int THREADS, int TILE>
static void launch_step_t(const gw_config& c, const gw::StepParams& p, unsigned blocks, cudaStream_t s) {
  const bool f32 = c.obs_dtype == GW_OBS_F32;
  const size_t smem = gw::smem_fixed_bytes<TILE>() + (size_t)(THREADS / 32) * 2 * gw::stage_row_bytes(c);
  if (c.fear) {
    if (f32) launch_k(gw::gw_step_kernel<THREADS, TILE, true, GW_OBS_F32>, p, blocks, THREADS, smem, s, use_pdl());
    else launch_k(gw::gw_step_kernel<THREADS, TILE, true, GW_OBS_BF16>, p, blocks, THREADS, smem, s, use_pdl());
  } else {
    if (f32) launch_k(gw::gw_step_kernel<THREADS, TILE, false, GW_OBS_F32>, p, blocks, THREADS, smem, s, use_pdl());
    else launch_k(gw::gw_step_kernel<THREADS, TILE, false, GW_OBS_BF16>, p, blocks, THREADS, smem, s, use_pdl());
  }
}

static void launch_step_small(const gw_config& c, gw::StepParams& p, unsigned blocks, int sm_count, cudaStream_t s) {
  // Programmatic dependent launch: when this grid and the next one fit on the machine together (one CTA per SM each), the
  // next step's CTAs start at once, copy their tables while this step runs and wait (griddepcontrol.wait) for its results.
  // GW_PDL=0 turns it off, GW_PDL=1 turns the late trigger on for every size (measured: it only pays for small grids).
  static const int pdl_env = [] { const char* v = std::getenv("GW_PDL"); return v ? std::atoi(v) : -1; }();
  p.pdl_early = (pdl_env != 0 && (int)blocks <= sm_count) ? 1 : 0;
  const bool f32 = c.obs_dtype == GW_OBS_F32;
  const size_t smem = gw::smem_fixed_bytes<32>() + (size_t)8 * 4 * gw::stage_row_bytes(c);
  const bool pdl = p.pdl_early != 0 || use_pdl();
  if (c.fear && use_split() && (int)blocks <= sm_count) {               // one tile per SM: FeAR on helper warps beside the step
    if (f32) launch_k(gw::gw_step_small_split_kernel<GW_OBS_F32>, p, blocks, 512, smem, s, pdl);
    else launch_k(gw::gw_step_small_split_kernel<GW_OBS_BF16>, p, blocks, 512, smem, s, pdl);
  } else if (c.fear) {
    if (f32) launch_k(gw::gw_step_small_kernel<true, GW_OBS_F32>, p, blocks, 256, smem, s, pdl);
    else launch_k(gw::gw_step_small_kernel<true, GW_OBS_BF16>, p, blocks, 256, smem, s, pdl);
  } else {
    if (f32) launch_k(gw::gw_step_small_kernel<false, GW_OBS_F32>, p, blocks, 256, smem, s, pdl);
    else launch_k(gw::gw_step_small_kernel<false, GW_OBS_BF16>, p, blocks, 256, smem, s, pdl);
  }
}

template <int THREADS, int TILE>
static void launch_reset_t(const gw_config& c, const gw::StepParams& p, unsigned blocks, cudaStream_t s) {
  const size_t smem = gw::smem_fixed_bytes<TILE>() + (size_t)(THREADS / 32) * 2 * gw::stage_row_bytes(c);
  if (c.obs_dtype == GW_OBS_F32) launch_k(gw::gw_reset_kernel<THREADS, TILE, GW_OBS_F32>, p, blocks, THREADS, smem, s);
  else launch_k(gw::gw_reset_kernel<THREADS, TILE, GW_OBS_BF16>, p, blocks, THREADS, smem, s);
}

// ---- resident step server, host side -------------------------------------------------------------------------------
static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#elif defined(__aarch64__)
  asm volatile("yield" ::: "memory");
#endif
}
static inline double wall_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline unsigned int host_load_u32(const unsigned int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }

static void server_free(gw_handle* h) {
  gw_server& v = h->srv;
  if (v.h_bell) cudaFreeHost(v.h_bell);
  if (v.d_bell) cudaFree(v.d_bell);
  if (v.d_io_table) cudaFree(v.d_io_table);
  if (v.copy_stream) cudaStreamDestroy(v.copy_stream);
  v = gw_server();
}

static int server_alloc(gw_handle* h) {
  gw_server& v = h->srv;
  if (v.allocated) return GW_OK;
  void* hp = nullptr;
  GW_CUDA(h, cudaHostAlloc(&hp, 512, cudaHostAllocPortable | cudaHostAllocMapped));
  std::memset(hp, 0, 512);
  v.h_bell = static_cast<unsigned long long*>(hp);
  v.h_resp = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(hp) + 128);
  v.h_stage = reinterpret_cast<gw_io*>(static_cast<uint8_t*>(hp) + 256);
  GW_CUDA(h, cudaStreamCreateWithFlags(&v.copy_stream, cudaStreamNonBlocking));
  void* dp = nullptr;
  GW_CUDA(h, cudaMalloc(&dp, 256));
  GW_CUDA(h, cudaMemset(dp, 0, 256));
  v.d_bell = static_cast<unsigned long long*>(dp);
  v.d_arrive = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(dp) + 128);
  GW_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&v.d_io_table), sizeof(gw_io) * GW_SRV_SETS));
  if (const char* e = std::getenv("GW_SERVER_IDLE_US")) {
    const long long us = std::atoll(e);
    if (us > 0) v.idle_ns = (unsigned long long)us * 1000ull;
  }
  v.sets.reserve(64);
  v.allocated = true;
  return GW_OK;
}

// The resident kernel leaves (STOP, or it had already left after idle_ns without a doorbell).  Everything it wrote is
// visible to later work on any stream once this returns.
static int server_stop(gw_handle* h) {
  gw_server& v = h->srv;
  if (!v.running) return GW_OK;
  v.running = false;
  v.seq += 1;
  __atomic_store_n(v.h_bell, ((unsigned long long)v.seq << 32) | ((unsigned long long)gw::SRV_OP_STOP << 16), __ATOMIC_RELEASE);
  GW_CUDA(h, cudaStreamSynchronize(v.stream));     // bounded: CTA 0 answers within a poll, and gives up by itself after idle_ns
  if (host_load_u32(v.h_resp + 1) != v.generation) return fail(h, GW_ECUDA, "step server ended without signing off");
  return GW_OK;
}

template <typename K>
static void launch_server_k(K kernel, const gw::StepParams& p, const gw::ServerParams& sp, unsigned blocks, size_t smem, cudaStream_t s,
                            int threads = 256) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kernel<<<blocks, threads, smem, s>>>(p, sp);
}


static int server_start(gw_handle* h, cudaStream_t s, unsigned int first_seq) {
  gw_server& v = h->srv;
  const gw_config& c = h->cfg;
  GW_CUDA(h, use_device(c.device));
  v.generation += 1;
  if (v.generation == 0) v.generation = 1;
  __atomic_store_n(v.h_resp + 1, 0u, __ATOMIC_RELEASE);
  GW_CUDA(h, cudaMemsetAsync(v.d_bell, 0, 256, s));
  gw::StepParams p = make_params(h, nullptr);
  p.pdl_early = 0;
  gw::ServerParams sp;
  sp.host_bell = v.h_bell; sp.host_resp = v.h_resp; sp.dev_bell = v.d_bell; sp.dev_arrive = v.d_arrive;
  sp.io_table = v.d_io_table; sp.first_seq = first_seq; sp.generation = v.generation; sp.idle_ns = v.idle_ns;
  const long long n_tiles = (c.num_envs + 31) / 32, resident = (long long)h->sm_count * 2;   // __launch_bounds__(256, 2), < 114 KB each
  const unsigned blocks = (unsigned)(n_tiles < resident ? n_tiles : resident);
  const size_t smem = gw::smem_fixed_bytes<32>() + (size_t)8 * 4 * gw::stage_row_bytes(c);
  const bool f32 = c.obs_dtype == GW_OBS_F32;
  if (c.fear && use_split() && n_tiles <= (long long)h->sm_count) {      // one tile per SM: FeAR on helper warps beside the step
    if (f32) launch_server_k(gw::gw_step_server_kernel<true, GW_OBS_F32, true>, p, sp, blocks, smem, s, 512);
    else launch_server_k(gw::gw_step_server_kernel<true, GW_OBS_BF16, true>, p, sp, blocks, smem, s, 512);
  } else if (c.fear) {
    if (f32) launch_server_k(gw::gw_step_server_kernel<true, GW_OBS_F32>, p, sp, blocks, smem, s);
    else launch_server_k(gw::gw_step_server_kernel<true, GW_OBS_BF16>, p, sp, blocks, smem, s);
  } else {
    if (f32) launch_server_k(gw::gw_step_server_kernel<false, GW_OBS_F32>, p, sp, blocks, smem, s);
    else launch_server_k(gw::gw_step_server_kernel<false, GW_OBS_BF16>, p, sp, blocks, smem, s);
  }
  GW_CUDA(h, cudaGetLastError());
  // a kernel that has already left when its launch returns sat out its idle time inside the launch call
  v.blocked_starts = (host_load_u32(v.h_resp + 1) == v.generation) ? v.blocked_starts + 1 : 0;
  v.running = true;
  v.stream = s;
  v.launches += 1;
  h->launches += 1;
  return GW_OK;
}

// index of this buffer set in the device table (registered on first sight)
static int server_set_index(gw_handle* h, const gw_io& z, int* out) {
  gw_server& v = h->srv;
  if (v.last_set >= 0 && std::memcmp(&v.sets[v.last_set], &z, sizeof(gw_io)) == 0) { *out = v.last_set; return GW_OK; }
  const int n = (int)v.sets.size();
  for (int k = 1; k <= n; ++k) {                    // callers cycle through their buffers: look just after the last hit first
    const int i = (v.last_set + k + n) % n;
    if (std::memcmp(&v.sets[i], &z, sizeof(gw_io)) == 0) { *out = v.last_set = i; return GW_OK; }
  }
  if (n >= GW_SRV_SETS) {                           // full: start over (entries in use must not change under the kernel)
    if (int rc = server_stop(h)) return rc;
    v.sets.clear();
  }
  // Appending is safe while the kernel runs (no doorbell names the new entry yet); the copy goes through its own
  // stream, since anything on the kernel's stream would wait for the kernel.
  const int i = (int)v.sets.size();
  v.sets.push_back(z);
  std::memcpy(v.h_stage, &z, sizeof(gw_io));
  GW_CUDA(h, cudaMemcpyAsync(v.d_io_table + i, v.h_stage, sizeof(gw_io), cudaMemcpyHostToDevice, v.copy_stream));
  GW_CUDA(h, cudaStreamSynchronize(v.copy_stream));
  *out = v.last_set = i;
  return GW_OK;
}

constexpr int SRV_RETRY_PLAIN = 1;     // server_step: not served, take the launch-per-step path

static int server_step(gw_handle* h, const gw_io& z, cudaStream_t s) {
  gw_server& v = h->srv;
  if (v.blocked_starts >= 2) {
    // Launches do not return while the kernel runs (a profiler or CUDA_LAUNCH_BLOCKING serialises them): every step
    // would cost two idle periods.  This handle goes back to one launch per step.
    v.disabled = true;
    return SRV_RETRY_PLAIN;
  }
  if (int rc = server_alloc(h)) return rc;
  int set = 0;
  if (int rc = server_set_index(h, z, &set)) return rc;
  if (v.running && v.stream != s)
    if (int rc = server_stop(h)) return rc;
  if (v.seq >= 0xFFFFFF00u) {                        // sequence numbers are 32 bits: start over
    if (int rc = server_stop(h)) return rc;
    v.seq = 0;
    __atomic_store_n(v.h_bell, 0ull, __ATOMIC_RELEASE);
    __atomic_store_n(v.h_resp, 0u, __ATOMIC_RELEASE);
  }
  const unsigned int seq = ++v.seq;
  if (!v.running)
    if (int rc = server_start(h, s, seq)) return rc;
  __atomic_store_n(v.h_bell, ((unsigned long long)seq << 32) | ((unsigned long long)gw::SRV_OP_STEP << 16) | (unsigned long long)set,
                   __ATOMIC_RELEASE);
  double t0 = 0.0;
  int starts = 0;
  for (unsigned long long spins = 1;; ++spins) {
    if (host_load_u32(v.h_resp) == seq) break;
    if (host_load_u32(v.h_resp + 1) == v.generation) {
      // the kernel left (idle) -- either before it saw this doorbell, or right after completing it
      v.running = false;
      GW_CUDA(h, cudaStreamSynchronize(s));
      if (host_load_u32(v.h_resp) == seq) break;
      if (++starts > 2) {                       // keeps leaving before it sees the doorbell: one launch per step from now on
        v.disabled = true;                      // (nothing of this step has run: CTA 0 alone accepts a doorbell, and it did not)
        return SRV_RETRY_PLAIN;
      }
      v.relaunches += 1;
      if (int rc = server_start(h, s, seq)) return rc;   // the doorbell still rings
      t0 = 0.0;
      continue;
    }
    if ((spins & 0x3FFFu) == 0) {
      const double now = wall_s();
      if (t0 == 0.0) t0 = now;
      const cudaError_t q = cudaStreamQuery(s);
      if (q != cudaErrorNotReady && q != cudaSuccess) { v.running = false; return cuda_fail(h, q, "step server"); }
      if (now - t0 > 10.0) { v.running = false; return fail(h, GW_ECUDA, "step server: no completion within 10 s"); }
    }
    cpu_relax();
  }
  h->env_steps += (uint64_t)h->cfg.num_envs;
  return GW_OK;
}

extern "C" {

int gw_server_stop(gw_handle* h) {
  if (!h) return GW_EINVAL;
  return server_stop(h);
}

int gw_server_info(gw_handle* h, int* running, uint64_t* launches, uint64_t* relaunches, int* n_sets) {
  if (!h) return GW_EINVAL;
  gw_server& v = h->srv;
  if (running) *running = v.disabled ? -1 : ((v.running && host_load_u32(v.h_resp + 1) != v.generation) ? 1 : 0);
  if (launches) *launches = v.launches;
  if (relaunches) *relaunches = v.relaunches;
  if (n_sets) *n_sets = (int)v.sets.size();
  return GW_OK;
}

int gw_reset(gw_handle* h, const uint8_t* reset_mask, const gw_io* io, void* stream) {
  if (!h) return GW_EINVAL;
  if (int rc = server_stop(h)) return rc;
  if (int rc = check_io(h, io, false)) return rc;
  if (!h->reset_done && reset_mask) return fail(h, GW_ESTATE, "gw_reset: the first reset must cover all envs (reset_mask = NULL)");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  gw::StepParams p = make_params(h, io);
  p.reset_mask = reset_mask;
  const int tile = pick_tile(h->cfg.num_envs);
  const unsigned blocks = (unsigned)((h->cfg.num_envs + tile - 1) / tile);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (tile == 32) launch_reset_t<256, 32>(h->cfg, p, blocks, s);
  else if (tile == 128) launch_reset_t<256, 128>(h->cfg, p, blocks, s);
  else launch_reset_t<256, 256>(h->cfg, p, blocks, s);
  GW_CUDA(h, cudaGetLastError());
  h->reset_done = true;
  h->launches += 1;
  return GW_OK;
}

int gw_step(gw_handle* h, const gw_io* io, void* stream) {
  if (!h) return GW_EINVAL;
  if (int rc = server_stop(h)) return rc;
  if (int rc = check_io(h, io, true)) return rc;
  if (!h->reset_done) return fail(h, GW_ESTATE, "gw_step: call gw_reset first");
  GW_CUDA(h, use_device(h->cfg.device));
  gw::StepParams p = make_params(h, io);
  const int tile = pick_tile(h->cfg.num_envs);
  const long long n_tiles = (h->cfg.num_envs + tile - 1) / tile;
  // persistent CTAs: at most `sm_count x resident CTAs per SM`, each walks several tiles and loads the tables once
  const bool small_kernel = pick_small(h->cfg.num_envs);
  const long long resident = (long long)h->sm_count * ((tile == 32 && small_kernel) ? 2 : 4);
  const unsigned blocks = (unsigned)(n_tiles < resident ? n_tiles : resident);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (tile == 32 && small_kernel) launch_step_small(h->cfg, p, blocks, h->sm_count, s);   // latency regime: 8 lanes per env, no CTA-wide phases
  else if (tile == 32) launch_step_t<256, 32>(h->cfg, p, blocks, s);
  else if (tile == 128) launch_step_t<256, 128>(h->cfg, p, blocks, s);
  else launch_step_t<256, 256>(h->cfg, p, blocks, s);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  h->env_steps += (uint64_t)h->cfg.num_envs;
  return GW_OK;
}

}  // extern "C"

template <typename K>
static void launch_rollout_k(K kernel, const gw::StepParams& p, const gw::RollParams& rp, unsigned blocks, size_t smem, cudaStream_t s,
                             int threads = 256) {
  static thread_local const void* raised_fn = nullptr;
  static thread_local size_t raised = 0;
  if (raised_fn != (const void*)kernel || raised < smem) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    raised_fn = (const void*)kernel;
    raised = smem;
  }
  kernel<<<blocks, threads, smem, s>>>(p, rp);
}

extern "C" {

int gw_rollout(gw_handle* h, const gw_io* rings, const gw_rollout_plan* plan, void* stream) {
  if (!h) return GW_EINVAL;
  if (int rc = server_stop(h)) return rc;
  if (!plan || plan->struct_size != sizeof(gw_rollout_plan)) return fail(h, GW_EINVAL, "gw_rollout: plan missing or struct_size mismatch");
  if (int rc = check_io(h, rings, true)) return rc;
  if (!rings->obs) return fail(h, GW_EINVAL, "gw_rollout: the observation ring is required");
  if (plan->steps < 1 || plan->ring_slots < 1 || plan->action_slots < 1 || plan->first_slot < 0 || plan->first_slot >= plan->ring_slots ||
      plan->first_action < 0 || plan->first_action >= plan->action_slots)
    return fail(h, GW_EINVAL, "gw_rollout: steps >= 1, 0 <= first_slot < ring_slots, 0 <= first_action < action_slots");
  if (!h->reset_done) return fail(h, GW_ESTATE, "gw_rollout: call gw_reset first");
  const gw_config& c = h->cfg;
  // slot strides must keep every slot as aligned as slot 0 (check_io looked at slot 0)
  const size_t E = (size_t)c.num_envs, L = (size_t)c.n_learners;
  if ((rings->fear && (E * L * 8) % 16) || (rings->positions && (E * (size_t)c.n_agents * 2) % 8) ||
      ((rings->terminated || rings->truncated) && (E * L) % 2) || ((rings->reward || rings->shaped_reward) && (E * L * 4) % 8))
    return fail(h, GW_EINVAL, "gw_rollout: num_envs x n_learners must be even for time-major rings");
  GW_CUDA(h, use_device(c.device));
  gw::StepParams p = make_params(h, rings);
  p.pdl_early = 0;
  gw::RollParams rp;
  rp.steps = plan->steps; rp.ring_slots = plan->ring_slots; rp.first_slot = plan->first_slot;
  rp.action_slots = plan->action_slots; rp.first_action = plan->first_action;
  rp.wait_full = (long long)plan->steps >= plan->ring_slots ? 1 : 0;
  const unsigned blocks = (unsigned)((c.num_envs + 31) / 32);      // one tile per CTA: nothing is shared between tiles
  const size_t smem = gw::smem_fixed_bytes<32>() + (size_t)8 * 4 * gw::stage_row_bytes(c);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool f32 = c.obs_dtype == GW_OBS_F32;
  if (c.fear && use_split() && blocks <= (unsigned)h->sm_count) {       // one tile per SM: FeAR on helper warps beside the step
    if (f32) launch_rollout_k(gw::gw_rollout_split_kernel<GW_OBS_F32>, p, rp, blocks, smem, s, 512);
    else launch_rollout_k(gw::gw_rollout_split_kernel<GW_OBS_BF16>, p, rp, blocks, smem, s, 512);
  } else if (c.fear) {
    if (f32) launch_rollout_k(gw::gw_rollout_kernel<true, GW_OBS_F32>, p, rp, blocks, smem, s);
    else launch_rollout_k(gw::gw_rollout_kernel<true, GW_OBS_BF16>, p, rp, blocks, smem, s);
  } else {
    if (f32) launch_rollout_k(gw::gw_rollout_kernel<false, GW_OBS_F32>, p, rp, blocks, smem, s);
    else launch_rollout_k(gw::gw_rollout_kernel<false, GW_OBS_BF16>, p, rp, blocks, smem, s);
  }
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  h->env_steps += (uint64_t)c.num_envs * (uint64_t)plan->steps;
  return GW_OK;
}

int gw_step_host(gw_handle* h, const gw_io* io, const int8_t* host_actions, float* host_reward, float* host_shaped,
                 uint8_t* host_ended, int zero_copy, void* stream) {
  if (!h) return GW_EINVAL;
  if (zero_copy) {
    if (!io || !host_actions) return fail(h, GW_EINVAL, "gw_step_host: io and host_actions are required");
    gw_io z = *io;                                   // pinned host memory is device-addressable (UVA): point the kernel at it
    z.learner_actions = host_actions;
    if (host_reward) z.reward = host_reward;
    if (host_shaped) z.shaped_reward = host_shaped;
    if (host_ended) z.ended = host_ended;
    if (zero_copy == GW_HOST_RESIDENT && pick_resident(h->cfg.num_envs) && !h->srv.disabled) {
      if (int rc = check_io(h, &z, true)) return rc;
      if (!h->reset_done) return fail(h, GW_ESTATE, "gw_step_host: call gw_reset first");
      const int rc = server_step(h, z, static_cast<cudaStream_t>(stream));
      if (rc != SRV_RETRY_PLAIN) return rc;
    }
    if (int rc = gw_step(h, &z, stream)) return rc;
    GW_CUDA(h, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return GW_OK;
  }
  if (!io || !io->learner_actions || !host_actions) return fail(h, GW_EINVAL, "gw_step_host: io->learner_actions (device) and host_actions are required");
  if ((host_reward && !io->reward) || (host_shaped && !io->shaped_reward) || (host_ended && !io->ended))
    return fail(h, GW_EINVAL, "gw_step_host: a host destination needs the matching device output in io");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t E = (size_t)h->cfg.num_envs, L = (size_t)h->cfg.n_learners;
  GW_CUDA(h, cudaMemcpyAsync(const_cast<int8_t*>(io->learner_actions), host_actions, E * L, cudaMemcpyHostToDevice, s));
  if (int rc = gw_step(h, io, stream)) return rc;
  if (host_reward) GW_CUDA(h, cudaMemcpyAsync(host_reward, io->reward, E * L * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (host_shaped) GW_CUDA(h, cudaMemcpyAsync(host_shaped, io->shaped_reward, E * L * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (host_ended) GW_CUDA(h, cudaMemcpyAsync(host_ended, io->ended, E, cudaMemcpyDeviceToHost, s));
  GW_CUDA(h, cudaStreamSynchronize(s));
  return GW_OK;
}

int gw_host_call_prepare(gw_handle* h, const gw_io* io, const int8_t* host_actions, float* host_reward, float* host_shaped,
                         uint8_t* host_ended, int zero_copy, int* token) {
  if (!h || !io || !host_actions || !token) return fail(h, GW_EINVAL, "gw_host_call_prepare: null argument");
  if (h->host_calls.size() >= (size_t)1 << 20) return fail(h, GW_ENOMEM, "gw_host_call_prepare: too many prepared calls");
  h->host_calls.push_back(gw_host_call{*io, host_actions, host_reward, host_shaped, host_ended, zero_copy});
  *token = (int)h->host_calls.size() - 1;
  return GW_OK;
}

int gw_host_call_reset(gw_handle* h) {            // forget every prepared call (tokens become invalid)
  if (!h) return GW_EINVAL;
  h->host_calls.clear();
  return GW_OK;
}

int gw_host_call_run(gw_handle* h, int token, void* stream) {
  if (!h) return GW_EINVAL;
  if (token < 0 || (size_t)token >= h->host_calls.size()) return fail(h, GW_EINVAL, "gw_host_call_run: unknown token");
  const gw_host_call& c = h->host_calls[(size_t)token];
  return gw_step_host(h, &c.io, c.actions, c.reward, c.shaped, c.ended, c.mode, stream);
}

int gw_sync(gw_handle* h, void* stream) {
  if (!h) return GW_EINVAL;
  if (int rc = server_stop(h)) return rc;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  GW_CUDA(h, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  GW_CUDA(h, cudaGetLastError());
  return GW_OK;
}

size_t gw_state_bytes(const gw_handle* h) { return h ? sizeof(uint4) * (size_t)h->cfg.num_envs : 0; }

int gw_get_state(gw_handle* h, void* dst, int dst_is_device, void* stream) {
  if (!h || !dst) return fail(h, GW_EINVAL, "gw_get_state: null argument");
  if (int rc = server_stop(h)) return rc;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GW_CUDA(h, cudaMemcpyAsync(dst, h->d_state, gw_state_bytes(h), dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  if (!dst_is_device) GW_CUDA(h, cudaStreamSynchronize(s));
  return GW_OK;
}

int gw_set_state(gw_handle* h, const void* src, int src_is_device, void* stream) {
  if (!h || !src) return fail(h, GW_EINVAL, "gw_set_state: null argument");
  if (int rc = server_stop(h)) return rc;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GW_CUDA(h, cudaMemcpyAsync(h->d_state, src, gw_state_bytes(h), src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  if (!src_is_device) GW_CUDA(h, cudaStreamSynchronize(s));
  h->reset_done = true;
  return GW_OK;
}

int gw_get_stats(gw_handle* h, gw_stats* out, void* stream) {
  if (!h || !out) return fail(h, GW_EINVAL, "gw_get_stats: null argument");
  if (int rc = server_stop(h)) return rc;
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
    out->fear_tasks += r[gw::ST_TASKS];
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
  if (int rc = server_stop(h)) return rc;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  GW_CUDA(h, cudaMemsetAsync(h->d_stats, 0, sizeof(unsigned long long) * gw::STAT_SLOTS * 8, static_cast<cudaStream_t>(stream)));
  h->env_steps = 0;
  return GW_OK;
}

// dev only: GW_TRACE=1 makes gw_create allocate a per-CTA stamp buffer; gw_debug_trace copies it out (16 x uint64 per CTA)
int gw_debug_trace(gw_handle* h, unsigned long long* host_out, int max_ctas) {
  if (!h || !h->d_trace || !host_out) return GW_EINVAL;
  if (int rc = server_stop(h)) return rc;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  GW_CUDA(h, cudaDeviceSynchronize());
  GW_CUDA(h, cudaMemcpy(host_out, h->d_trace, sizeof(unsigned long long) * 16 * (size_t)max_ctas, cudaMemcpyDeviceToHost));
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
  if (int rc = server_stop(h)) return rc;
  if (n_cases < 0 || !positions || !actions || !new_positions || !crash || !restricted)
    return fail(h, GW_EINVAL, "gw_update_world: null/invalid argument");
  if (n_cases == 0) return GW_OK;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  const int threads = 128;
  const long long blocks = (n_cases + threads - 1) / threads;
  gw::gw_update_world_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d_tables, h->cfg.n_agents, n_cases, n_per, positions, actions, apples, new_positions, crash, restricted, caught);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_fear_matrix(gw_handle* h, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
                   const int8_t* mdr, const uint8_t* in_list, double* resp, int8_t* n_mdr, int8_t* n_act, void* stream) {
  if (!h) return GW_EINVAL;
  if (int rc = server_stop(h)) return rc;
  if (n_cases < 0 || !positions || !actions || !mdr || !resp || !n_mdr || !n_act)
    return fail(h, GW_EINVAL, "gw_fear_matrix: null/invalid argument");
  if (n_cases == 0) return GW_OK;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  const long long blocks = (n_cases * 4 * 32 + 127) / 128;
  gw::gw_fear_matrix_kernel<<<(unsigned)blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d_tables, h->cfg.n_agents, n_cases, n_per, positions, actions, mdr, in_list, resp, n_mdr, n_act);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_feal(gw_handle* h, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
            const int8_t* mdr, const uint8_t* in_list, double* feal, int8_t* n_mdr, int8_t* n_act, void* stream) {
  if (!h) return GW_EINVAL;
  if (int rc = server_stop(h)) return rc;
  if (n_cases < 0 || !positions || !actions || !mdr || !feal || !n_mdr || !n_act)
    return fail(h, GW_EINVAL, "gw_feal: null/invalid argument");
  if (n_cases == 0) return GW_OK;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  const long long blocks = (n_cases * 32 + 127) / 128;
  gw::gw_feal_kernel<<<(unsigned)blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d_tables, h->cfg.n_agents, n_cases, n_per, positions, actions, mdr, in_list, feal, n_mdr, n_act);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_fear_one_actor(gw_handle* h, int64_t n_cases, const int8_t* n_per, const int8_t* positions, const int8_t* actions,
                      const int8_t* mdr, const int8_t* actor, const uint8_t* in_list, double* resp, int8_t* n_mdr,
                      int8_t* n_act, void* stream) {
  if (!h) return GW_EINVAL;
  if (int rc = server_stop(h)) return rc;
  if (n_cases < 0 || !positions || !actions || !mdr || !actor || !resp)
    return fail(h, GW_EINVAL, "gw_fear_one_actor: null/invalid argument");
  if (n_cases == 0) return GW_OK;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  const int threads = 128;
  const long long blocks = (n_cases * 32 + threads - 1) / threads;
  gw::gw_fear_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d_tables, h->cfg.n_agents, n_cases, n_per, positions, actions, mdr, actor, in_list, resp, n_mdr, n_act);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

}  // extern "C"
