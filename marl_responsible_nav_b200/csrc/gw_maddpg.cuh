// Declarations shared by the two implementations of the fused MADDPG update (gw_maddpg.cu: grid-phase kernel, every
// supported shape; gw_maddpg_cluster.cu: row-block clusters, the fast path): argument block, scratch layout, network
// layout inside the flat parameter vectors, row-wise device helpers.  Not part of the C-ABI.
#pragma once
#include <cstring>
#include <map>
#include <string>

#include "gw_replay_dev.cuh"

namespace gwl {

constexpr int HID = 128;
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int MAXN = GW_MAX_LEARNERS;
constexpr int NA = GW_N_ACTIONS;           // action_dim: compile-time, the row phases keep per-action values in registers

struct NetLayout {                         // offsets (floats) inside one network's block; ln gamma = b + HID, beta = b + 2 HID
  int in, out, w1, b1, w2, b2, w3, b3, total, small;
};
__host__ __device__ inline NetLayout make_layout(int in, int out) {
  NetLayout L;
  L.in = in; L.out = out;
  L.w1 = 0; L.b1 = HID * in;
  L.w2 = L.b1 + 3 * HID; L.b2 = L.w2 + HID * HID;
  L.w3 = L.b2 + 3 * HID; L.b3 = L.w3 + out * HID;
  L.total = L.b3 + out;
  L.small = 6 * HID + out * (HID + 1);     // compact index space of the vector parameters: [b1 g1 be1 | b2 g2 be2 | w3 b3]
  return L;
}
__device__ __forceinline__ int compact_index(const NetLayout& L, int idx) {   // -1: a matrix element (w1 / w2)
  if (idx >= L.w3) return 6 * HID + (idx - L.w3);
  if (idx >= L.b2) return 3 * HID + (idx - L.b2);
  if (idx >= L.w2) return -1;
  if (idx >= L.b1) return idx - L.b1;
  return -1;
}

struct Pass { float *z1, *h1, *st1, *z2, *h2, *st2; };   // one forward pass of one network: [B,H] x4, statistics [B,2] x2

struct Scratch {
  float *S, *S2, *ACT, *R, *D;               // the gathered batch (fused sampling)
  Pass ta[MAXN], ct[MAXN], c[MAXN], ac[MAXN], c2[MAXN];
  float* a2;                                 // [B, n*A] target actors' actions on next_state
  float *anew[MAXN], *ax[MAXN];              // [B, A] actor i on state; [B, n*A] batch actions with block i replaced
  float *q[MAXN], *y[MAXN], *dq[MAXN];       // [B]
  float *dz2[MAXN], *dh1[MAXN], *dz1[MAXN];  // critic backward (TD pass, then the actor-loss pass)
  float *adz2[MAXN], *adh1[MAXN], *adz1[MAXN];
  float* pb[2 * MAXN];                       // [B / RB][small] partial sums per network
  float* lp;                                 // [2n][B / RB] loss partial sums (critics, then actors)
  unsigned* bar;                             // grid barrier: arrivals, generation
  unsigned long long* trace;                 // [PH_COUNT + 1] clock64 of CTA 0 at the start of every phase of the last update, and at the end
};

enum { ADAM_FROM_G = 1, ADAM_WRITE_G = 2, ADAM_APPLY = 4 };   // vector gradients already in G / store gradients in G / step

enum Phase {
  PH_GATHER = 0, PH_L1, PH_L2, PH_HEADS, PH_CT_L2, PH_TD, PH_C_BWD2, PH_C_LN1, PH_C_DW1, PH_ADAM_C,
  PH_C2_L1, PH_C2_L2, PH_ALOSS, PH_C2_DH1, PH_ACT_BWD, PH_A_BWD2, PH_A_LN1, PH_A_DW1, PH_ADAM_A, PH_COUNT
};

struct LearnArgs {
  int n, O, A, B, CI, SO, SA;                // agents, obs_len, action_dim, batch, critic input, n*O, n*A
  NetLayout la, lc;
  long long net_off[2 * MAXN];
  float *P, *T, *M, *V, *G, *steps;
  Scratch s;
  const float *bS, *bS2, *bACT, *bR, *bD;    // the batch the update reads (the staging above or the caller's tensors)
  const float *gum_next, *gum_cur;
  gw_replay_view ring;
  int sample;
  long long t_now, n_valid;
  uint32_t rk0, rk1, gk0, gk1;
  unsigned long long draw_base, upd_base;
  int updates, ph_begin, ph_end;
  int adam_mode[2];                          // [critics, actors]: ADAM_FROM_G | ADAM_WRITE_G | ADAM_APPLY
  float grad_scale, lr_a, lr_c, gamma, tau, beta1, beta2, eps, ln_eps;
  float* losses;
};

// ------------------------------------------------------------------------------------------------ small device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long phase_clock() {
  return (unsigned long long)clock64();           // CTA 0 stays on one SM: its cycle counter orders and times the phases
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float sum4(const float4 v) { return (v.x + v.y) + (v.z + v.w); }
__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
__device__ __forceinline__ float4 mul4(const float4 a, const float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 scale4(const float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ void st4(float* p, const float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Grid barrier (all CTAs are co-resident: cooperative launch).  Everything a CTA wrote before it is visible to every CTA
// after it: the arriving thread fences (cumulative over the CTA's writes through bar.sync), the last arrival bumps the
// generation, waiters spin on it with volatile loads.  Data produced inside the launch is only ever read with .cg loads
// (L2), so no stale L1 line can be hit.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned gen = *reinterpret_cast<volatile unsigned*>(bar + 1);
    __threadfence();
    if (atomicAdd(bar, 1u) == n_ctas - 1) {
      *reinterpret_cast<volatile unsigned*>(bar) = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*reinterpret_cast<volatile unsigned*>(bar + 1) == gen) {}
    }
    __threadfence();
  }
  __syncthreads();
}

// this lane's 4 columns of a LayerNorm'd row: xhat, and the statistics (computed here)
__device__ __forceinline__ float4 ln_row(const float4 z, float eps, float& mu, float& rs) {
  mu = warp_sum(sum4(z)) * (1.0f / HID);
  const float4 d = make_float4(z.x - mu, z.y - mu, z.z - mu, z.w - mu);
  const float var = warp_sum(dot4(d, d)) * (1.0f / HID);
  rs = rsqrtf(var + eps);
  return scale4(d, rs);
}
__device__ __forceinline__ float4 affine_relu(const float4 xh, const float4 g, const float4 b) {
  return make_float4(fmaxf(fmaf(xh.x, g.x, b.x), 0.f), fmaxf(fmaf(xh.y, g.y, b.y), 0.f), fmaxf(fmaf(xh.z, g.z, b.z), 0.f),
                     fmaxf(fmaf(xh.w, g.w, b.w), 0.f));
}
// LayerNorm + ReLU backward for one row: dh -> dz; dy (= dh where the ReLU passed) and xhat are returned for dgamma / dbeta
__device__ __forceinline__ float4 ln_relu_bwd_row(const float4 dh, const float4 h, const float4 xh, const float4 g, float rs,
                                                  float4& dy) {
  dy = make_float4(h.x > 0.f ? dh.x : 0.f, h.y > 0.f ? dh.y : 0.f, h.z > 0.f ? dh.z : 0.f, h.w > 0.f ? dh.w : 0.f);
  const float4 dxh = mul4(dy, g);
  const float m1 = warp_sum(sum4(dxh)) * (1.0f / HID);
  const float m2 = warp_sum(dot4(dxh, xh)) * (1.0f / HID);
  return make_float4(rs * (dxh.x - m1 - xh.x * m2), rs * (dxh.y - m1 - xh.y * m2), rs * (dxh.z - m1 - xh.z * m2),
                     rs * (dxh.w - m1 - xh.w * m2));
}
__device__ __forceinline__ float4 xhat_of(const float4 z, float mu, float rs) {
  return make_float4((z.x - mu) * rs, (z.y - mu) * rs, (z.z - mu) * rs, (z.w - mu) * rs);
}

// Gumbel noise of one (update, row, agent, which) as 9 values: g = -log(-log(u) + 1e-20), u uniform in (0, 1)
__device__ __forceinline__ float gumbel_of(uint32_t w) {
  const float u = ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return -logf(-logf(u) + 1e-20f);
}
__device__ __forceinline__ void gumbel_row(const LearnArgs& a, unsigned long long upd, int row, int agent, int which, float (&g)[NA]) {
#pragma unroll
  for (int c = 0; 4 * c < NA; ++c) {
    uint32_t w[4] = {(uint32_t)row, (uint32_t)agent | ((uint32_t)which << 8) | ((uint32_t)c << 16), (uint32_t)upd,
                     (uint32_t)(upd >> 32) ^ 0x6C6561u};
    gw::philox4x32(w, a.gk0, a.gk1);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * c + i < NA) g[4 * c + i] = gumbel_of(w[i]);
  }
}

}  // namespace gwl

// host side of a learner (gw_maddpg.cu); the cluster implementation adds its own launch state
struct gw_learner {
  gw_handle* h = nullptr;
  gw_learner_config cfg;
  gw_learner_layout lay;
  gw_learner_buffers buf;
  gwl::LearnArgs args;
  std::map<std::string, std::pair<float*, int64_t>> dbg;
  size_t smem = 0;
  int grid = 0;
  unsigned long long updates_done = 0;
  float* cluster_scratch = nullptr;          // gradient slabs of the cluster kernel (inside the caller's scratch block)
  int cluster_grid = 0;                      // CTAs of the cluster kernel, 0: shape unsupported / clusters not co-resident
  int kernel_kind = GW_LEARN_KERNEL_AUTO;
  int cluster_max_active = -1;
  // gradient exchange over peer memory (gw_learner_peer_export / _connect): exchange block of every rank, mapped here
  int peer_world = 1, peer_rank = 0;
  void* peer_base[GW_MAX_PEERS] = {};         // [rank] = this process's own cudaMalloc block; the others cudaIpcOpenMemHandle
  size_t peer_flag_off = 0;                   // bytes of the flag-in-data slots (GW_MAX_PEERS of them); the error word follows
  unsigned long long peer_epochs = 0;         // exchanges done so far (two per update)               // cudaOccupancyMaxActiveClusters of the cluster kernel on this device
};

// gw_maddpg_cluster.cu
bool gwc_supported(const gw_learner_config& c);
int64_t gwc_scratch_floats(const gw_learner_config& c);
int gwc_prepare(gw_learner* l);
int gwc_launch(gw_learner* l, const gwl::LearnArgs& a, int segment, cudaStream_t stream);
