// Fused MADDPG update: `agent.learn(experiences)` of the reference's trainer (maddpg/agent.py:209-213 and
// :218-224, AgileRL's MADDPG.learn behind it) as ONE persistent cooperative kernel that runs whole updates -- batch draw
// and gather from the device replay ring, target actors, TD target, critic forward / backward, Adam, actor loss through
// the updated critic, actor backward, Adam, soft target update -- for all agents, `updates` times per launch.
//
// Why one kernel: at BATCH_SIZE 128 an update is 0.23 GFLOP spread over ~20 dependent steps; as ~85 dependent library
// kernels (round 1) it took 321 us.  Here every step is a PHASE of a grid-resident kernel (one CTA per SM), phases are
// separated by a grid barrier (one atomic arrive + a spin on a generation word, ~0.5 us), and all intermediate tensors
// stay in L2.  Arithmetic is fp32 FMA like the reference's (PyTorch fp32): the update is latency-bound, tensor cores
// would buy nothing at these sizes and cost the 1e-4 agreement with fp32 autograd the tests ask for.
//
// Work decomposition
//   GEMM phases   32 x 32 output tiles, one tile per CTA per phase; the CTA stages both operand panels over the whole K
//                 in shared memory (cp.async, L2 -> smem, bypassing L1 because the operands change inside the launch), its
//                 8 warps split K (each warp a [32 x 32] partial tile as 4 x 8 accumulators per thread, 12 LDS.128 per
//                 128 FMA, conflict-free by construction: see tile_mma), and a shared-memory reduction joins them.
//                 Three operand layouts cover forward (x W^T), input gradients (dy W) and weight gradients (dy^T x).
//                 LayerNorm + ReLU of the previous layer is the PROLOGUE of the consuming GEMM (rows normalised in
//                 shared memory; the first column tile also stores h and the row statistics for the backward pass).
//   row phases    everything that is row-wise -- output layers (128 -> 9 / 1), Gumbel-softmax, TD target, loss gradients,
//                 LayerNorm backward, the 9-column slice of the critic's input gradient -- runs warp-per-row, 8 rows per
//                 CTA; what must be summed over the batch (bias / gamma / beta / output-layer gradients, losses) leaves
//                 as per-CTA partial sums that the Adam phase adds in a fixed order (deterministic, no atomics).
//   Adam phases   element-wise over the flat parameter vectors: gradient (matrix gradients from the GEMM phases, vector
//                 gradients from the partial sums), Adam step (torch.optim.Adam arithmetic), soft update of the target.
#include "gw_maddpg.cuh"

namespace gwl {

constexpr int TM = 32, TN = 32;
constexpr int RED_LD = 36;                 // reduction tile row stride: conflict-free for all three thread mappings
constexpr int RB = WARPS;                  // rows per row-phase task (one per warp)
constexpr int MAX_SLOTS = 12;              // 128-wide partial-sum vectors per row task (actor backward: 9 + 3)
static_assert(GW_N_ACTIONS + 3 <= MAX_SLOTS && GW_N_ACTIONS <= 16, "partial-sum slots");


// `rows` rows of `len` floats, row r at src + r * ld -> dst + r * dld (shared); columns [len, fill) are zero-filled.
// 16-byte cp.async where alignment allows, else 8-/4-byte L2 loads.
__device__ void stage_rows(float* dst, int dld, const float* src, long long ld, int rows, int len, int fill) {
  const int tid = threadIdx.x;
  const uintptr_t sa = reinterpret_cast<uintptr_t>(src), da = reinterpret_cast<uintptr_t>(dst);
  if (((sa | da) & 15) == 0 && (ld & 3) == 0 && (len & 3) == 0) {
    const int per = len >> 2;
    for (int i = tid; i < rows * per; i += THREADS) {
      const int r = i / per, c = (i - r * per) << 2;
      cp_async16(dst + r * dld + c, src + r * ld + c);
    }
  } else if (((sa | da) & 7) == 0 && (ld & 1) == 0 && (len & 1) == 0 && (dld & 1) == 0) {
    const int per = len >> 1;
#pragma unroll 4
    for (int i = tid; i < rows * per; i += THREADS) {
      const int r = i / per, c = (i - r * per) << 1;
      *reinterpret_cast<float2*>(dst + r * dld + c) = __ldcg(reinterpret_cast<const float2*>(src + r * ld + c));
    }
  } else {
#pragma unroll 4
    for (int i = tid; i < rows * len; i += THREADS) {
      const int r = i / len, c = i - r * len;
      dst[r * dld + c] = __ldcg(src + r * ld + c);
    }
  }
  if (fill > len) {
    const int w = fill - len;
    for (int i = tid; i < rows * w; i += THREADS) {
      const int r = i / w;
      dst[r * dld + len + (i - r * w)] = 0.f;
    }
  }
}

// shared row stride of a K-major operand panel: K rounded up to 4, quarter count odd => the 8 (or 4) rows a warp reads
// in one LDS.128 fall into disjoint bank quads
__host__ __device__ inline int pad_k(int K) {
  int k4 = (K + 3) >> 2;
  if ((k4 & 1) == 0) ++k4;
  return k4 << 2;
}

// ------------------------------------------------------------------------------------------------ GEMM tile
// C[m0.., n0..] (32 x 32) = sum_k A(m, k) B(n, k) (+ bias[n]).  var 0: A [m][k], B [n][k] in global memory (forward:
// x W^T); var 1: A [m][k], B [k][n] (input gradient: dy W); var 2: A [k][m], B [k][n] (weight gradient: dy^T x).
struct GemmJob {
  int var, M, N, K;
  const float* A; long long lda;
  const float* A2; long long lda2; int ksplit;      // var 0: columns k >= ksplit of the A rows come from A2 (critic input = [obs | actions])
  const float* Bp; long long ldb;
  const float* B2; long long ldb2; int nsplit;      // var 2: columns n >= nsplit of the B rows come from B2
  float* C; long long ldc;
  const float* bias;
  int pro;                                          // 0: none; 1: LayerNorm + ReLU on the A rows (K = HID); 2: first add act W_act^T
  const float *gamma, *beta;
  float *h_out, *st_out;                            // written by the first column tile (pro >= 1), nullable
  const float* act; int n_act; const float* w_act; long long ldw;   // pro 2: act [M][n_act], w_act[o * ldw + k]
};

__device__ __forceinline__ void fma44(float (&acc)[4][8], const float4 (&a)[4], const float4 (&b)[8]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      acc[r][c] = fmaf(a[r].x, b[c].x, acc[r][c]);
      acc[r][c] = fmaf(a[r].y, b[c].y, acc[r][c]);
      acc[r][c] = fmaf(a[r].z, b[c].z, acc[r][c]);
      acc[r][c] = fmaf(a[r].w, b[c].w, acc[r][c]);
    }
}
__device__ __forceinline__ void fma_outer(float (&acc)[4][8], const float (&a)[4], const float4 b0, const float4 b1) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    acc[r][0] = fmaf(a[r], b0.x, acc[r][0]); acc[r][1] = fmaf(a[r], b0.y, acc[r][1]);
    acc[r][2] = fmaf(a[r], b0.z, acc[r][2]); acc[r][3] = fmaf(a[r], b0.w, acc[r][3]);
    acc[r][4] = fmaf(a[r], b1.x, acc[r][4]); acc[r][5] = fmaf(a[r], b1.y, acc[r][5]);
    acc[r][6] = fmaf(a[r], b1.z, acc[r][6]); acc[r][7] = fmaf(a[r], b1.w, acc[r][7]);
  }
}

// LayerNorm + ReLU of 32 rows of HID floats held in shared memory (row stride ld), in place; warp w owns rows 4w..4w+3.
__device__ void ln_relu_rows_smem(float* As, int ld, int m0, const float* gamma, const float* beta, float eps, float* h_out,
                                  float* st_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 g = ldcg4(gamma + 4 * lane), be = ldcg4(beta + 4 * lane);
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp * 4 + rr;
    float* row = As + r * ld + 4 * lane;
    const float4 v = *reinterpret_cast<const float4*>(row);
    const float mu = warp_sum(sum4(v)) * (1.0f / HID);
    const float4 d = make_float4(v.x - mu, v.y - mu, v.z - mu, v.w - mu);
    const float var = warp_sum(dot4(d, d)) * (1.0f / HID);
    const float rs = rsqrtf(var + eps);
    float4 o;
    o.x = fmaxf(fmaf(d.x * rs, g.x, be.x), 0.f);
    o.y = fmaxf(fmaf(d.y * rs, g.y, be.y), 0.f);
    o.z = fmaxf(fmaf(d.z * rs, g.z, be.z), 0.f);
    o.w = fmaxf(fmaf(d.w * rs, g.w, be.w), 0.f);
    *reinterpret_cast<float4*>(row) = o;
    if (h_out) st4(h_out + (long long)(m0 + r) * HID + 4 * lane, o);
    if (st_out && lane == 0) *reinterpret_cast<float2*>(st_out + 2 * (m0 + r)) = make_float2(mu, rs);
  }
}

__device__ void gemm_tile(const GemmJob& j, int mt, int nt, float* smem, float ln_eps) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lm = lane >> 2, ln = lane & 3;
  const int m0 = mt * TM, n0 = nt * TN;
  const int nvalid = min(TN, j.N - n0);
  const int K = j.K, kceil = (K + 3) & ~3, ksteps = kceil >> 2;
  const int lda_s = j.var == 2 ? TM : pad_k(K), ldb_s = j.var == 0 ? pad_k(K) : TN;
  float* As = smem;
  float* Bs = As + (j.var == 2 ? K * TM : TM * lda_s);
  float* red = Bs + (j.var == 0 ? TN * ldb_s : K * TN);

  // ---- stage the operand panels (the whole K)
  if (j.var != 2) {
    const int k1 = j.A2 ? j.ksplit : K;
    stage_rows(As, lda_s, j.A + (long long)m0 * j.lda, j.lda, TM, k1, j.A2 ? k1 : kceil);
    if (j.A2) stage_rows(As + k1, lda_s, j.A2 + (long long)m0 * j.lda2, j.lda2, TM, K - k1, kceil - k1);
  } else {
    stage_rows(As, TM, j.A + m0, j.lda, K, TM, TM);
  }
  if (j.var == 0) {
    stage_rows(Bs, ldb_s, j.Bp + (long long)n0 * j.ldb, j.ldb, TN, K, kceil);
  } else if (j.B2 && n0 >= j.nsplit) {
    stage_rows(Bs, TN, j.B2 + (n0 - j.nsplit), j.ldb2, K, nvalid, TN);
  } else {
    stage_rows(Bs, TN, j.Bp + n0, j.ldb, K, nvalid, TN);
  }
  if (j.pro == 2) {                                 // action columns of the first layer: act rows and the weight slice
    stage_rows(red, j.n_act, j.act + (long long)m0 * j.n_act, j.n_act, TM, j.n_act, j.n_act);
    stage_rows(red + TM * j.n_act, j.n_act, j.w_act, j.ldw, HID, j.n_act, j.n_act);
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- prologue on the A rows
  if (j.pro == 2) {
    const float* sa = red;
    const float* sw = red + TM * j.n_act;
    for (int i = tid; i < TM * HID; i += THREADS) {
      const int r = i >> 7, o = i & (HID - 1);
      float acc = As[r * lda_s + o];
      for (int k = 0; k < j.n_act; ++k) acc = fmaf(sa[r * j.n_act + k], sw[o * j.n_act + k], acc);
      As[r * lda_s + o] = acc;
    }
    __syncthreads();
  }
  if (j.pro >= 1) {
    ln_relu_rows_smem(As, lda_s, m0, j.gamma, j.beta, ln_eps, nt == 0 ? j.h_out : nullptr, nt == 0 ? j.st_out : nullptr);
    __syncthreads();
  }

  // ---- the warp's partial tile over its share of K
  float acc[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
  if (j.var == 0) {                                 // rows lm + 8r, columns ln + 4c
    for (int k4 = warp; k4 < ksteps; k4 += WARPS) {
      float4 a[4], b[8];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(As + (lm + 8 * r) * lda_s + 4 * k4);
#pragma unroll
      for (int c = 0; c < 8; ++c) b[c] = *reinterpret_cast<const float4*>(Bs + (ln + 4 * c) * ldb_s + 4 * k4);
      fma44(acc, a, b);
    }
  } else if (j.var == 1) {                          // rows lm + 8r, columns 8 ln + c
    for (int k4 = warp; k4 < ksteps; k4 += WARPS) {
      float4 a[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(As + (lm + 8 * r) * lda_s + 4 * k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float* brow = Bs + (4 * k4 + kk) * TN + 8 * ln;
        const float4 b0 = *reinterpret_cast<const float4*>(brow), b1 = *reinterpret_cast<const float4*>(brow + 4);
        const float av[4] = {kk == 0 ? a[0].x : kk == 1 ? a[0].y : kk == 2 ? a[0].z : a[0].w,
                             kk == 0 ? a[1].x : kk == 1 ? a[1].y : kk == 2 ? a[1].z : a[1].w,
                             kk == 0 ? a[2].x : kk == 1 ? a[2].y : kk == 2 ? a[2].z : a[2].w,
                             kk == 0 ? a[3].x : kk == 1 ? a[3].y : kk == 2 ? a[3].z : a[3].w};
        fma_outer(acc, av, b0, b1);
      }
    }
  } else {                                          // rows 4 lm + r, columns 8 ln + c
    for (int k4 = warp; k4 < ksteps; k4 += WARPS) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = 4 * k4 + kk;
        const float4 a = *reinterpret_cast<const float4*>(As + k * TM + 4 * lm);
        const float* brow = Bs + k * TN + 8 * ln;
        const float4 b0 = *reinterpret_cast<const float4*>(brow), b1 = *reinterpret_cast<const float4*>(brow + 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
        fma_outer(acc, av, b0, b1);
      }
    }
  }

  // ---- join the 8 partial tiles, add the bias, store
  float* my = red + warp * (TM * RED_LD);
  if (j.pro == 2) __syncthreads();                  // the prologue's staging lived in `red`
  if (j.var == 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) my[(lm + 8 * r) * RED_LD + ln + 4 * c] = acc[r][c];
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = j.var == 1 ? lm + 8 * r : 4 * lm + r;
      st4(my + m * RED_LD + 8 * ln, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
      st4(my + m * RED_LD + 8 * ln + 4, make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]));
    }
  }
  __syncthreads();
  {
    const int m = tid >> 3, nq = (tid & 7) << 2;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(red + w * (TM * RED_LD) + m * RED_LD + nq);
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    const float o[4] = {sum.x, sum.y, sum.z, sum.w};
    float* crow = j.C + (long long)(m0 + m) * j.ldc + n0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (nq + i < nvalid) crow[nq + i] = o[i] + (j.bias ? __ldcg(j.bias + n0 + nq + i) : 0.f);
  }
  __syncthreads();                                  // shared memory is free for the CTA's next task
}

// ------------------------------------------------------------------------------------------------ row-phase helpers
struct RowSmem {                                    // carved from the dynamic shared memory in row phases
  float* part;                                      // [WARPS][slots][HID]
  float* sc;                                        // [WARPS][16]
  float* wa;                                        // [A][HID] critic first-layer action columns (actor backward)
};


// partial sums of one row task -> global [rb][small]: vectors by compact offset, scalars likewise
__device__ void flush_partials(const RowSmem& rs, int slots, const int* slot_off, float* pb_rb, int n_sc, const int* sc_off) {
  __syncthreads();
  for (int i = threadIdx.x; i < slots * HID; i += THREADS) {
    const int sl = i >> 7, col = i & (HID - 1);
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) v += rs.part[(w * slots + sl) * HID + col];
    pb_rb[slot_off[sl] + col] = v;
  }
  if (threadIdx.x < n_sc) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) v += rs.sc[w * 16 + threadIdx.x];
    pb_rb[sc_off[threadIdx.x]] = v;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(THREADS, 1) gw_learn_kernel(const LearnArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_off[MAX_SLOTS + 16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = a.n, B = a.B, O = a.O, SO = a.SO, SA = a.SA, CI = a.CI;
  constexpr int A = NA;
  const int MT = B / TM, NRB = B / RB;
  const NetLayout la = a.la, lc = a.lc;
  const unsigned n_ctas = gridDim.x;
  float step0[2 * MAXN];
  for (int k = 0; k < 2 * n; ++k) step0[k] = __ldcg(a.steps + k);
  RowSmem rs;
  rs.part = smem;
  rs.sc = smem + WARPS * MAX_SLOTS * HID;
  rs.wa = rs.sc + WARPS * 16;

  auto actor_p = [&](const float* base, int i) { return base + a.net_off[i]; };
  auto critic_p = [&](const float* base, int i) { return base + a.net_off[n + i]; };

  for (int u = 0; u < a.updates; ++u) {
    const unsigned long long upd = a.upd_base + (unsigned long long)u;
    for (int ph = a.ph_begin; ph < a.ph_end; ++ph) {
      if (blockIdx.x == 0 && tid == 0 && u == a.updates - 1) a.s.trace[ph] = phase_clock();
      switch (ph) {
        // ============================================================ batch draw + gather (maddpg/agent.py:209-211)
        case PH_GATHER: {
          if (!a.sample) break;
          const gw_replay_view& r = a.ring;
          const int n_obs = n * O;
          for (long long b = (long long)blockIdx.x * WARPS + warp; b < B; b += (long long)n_ctas * WARPS) {
            long long t_abs = 0, e = 0;
            if (lane == 0) gwr::draw_index(b, a.draw_base + (unsigned long long)u, a.rk0, a.rk1, a.t_now, a.n_valid, r.num_envs, t_abs, e);
            t_abs = __shfl_sync(0xffffffffu, t_abs, 0);
            e = __shfl_sync(0xffffffffu, e, 0);
            long long sl = t_abs % r.slots;
            if (sl < 0) sl += r.slots;
            const long long sl1 = sl + 1 == r.slots ? 0 : sl + 1;
            const long long row = sl * r.num_envs + e, row1 = sl1 * r.num_envs + e;
            const bool ended = r.ended[row] != 0;
            if (r.obs_dtype == GW_OBS_F32) {
              const float* obs = static_cast<const float*>(r.obs);
              const float* fin = static_cast<const float*>(r.final_obs);
              gwr::copy_row<float>(obs + row * n_obs, a.s.S + b * n_obs, n_obs, lane, true);
              gwr::copy_row<float>(ended ? fin + row * n_obs : obs + row1 * n_obs, a.s.S2 + b * n_obs, n_obs, lane, true);
            } else {
              const __nv_bfloat16* obs = static_cast<const __nv_bfloat16*>(r.obs);
              const __nv_bfloat16* fin = static_cast<const __nv_bfloat16*>(r.final_obs);
              gwr::copy_row<__nv_bfloat16>(obs + row * n_obs, a.s.S + b * n_obs, n_obs, lane, true);
              gwr::copy_row<__nv_bfloat16>(ended ? fin + row * n_obs : obs + row1 * n_obs, a.s.S2 + b * n_obs, n_obs, lane, true);
            }
            for (int i = lane; i < SA; i += 32) a.s.ACT[b * SA + i] = r.action[row * SA + i];
            if (lane < n) {
              a.s.R[b * n + lane] = r.reward[row * n + lane];
              a.s.D[b * n + lane] = r.terminated[row * n + lane] ? 1.0f : 0.0f;
            }
          }
        } break;

        // ============================================================ first layers of everything that only needs the batch
        case PH_L1: {
          const int per = MT * (HID / TN);
          for (int t = blockIdx.x; t < 4 * n * per; t += n_ctas) {
            const int g = t / per, tile = t - g * per, mt = tile / (HID / TN), nt = tile % (HID / TN);
            const int kind = g / n, i = g - kind * n;
            GemmJob j;
            memset(&j, 0, sizeof(j));
            j.var = 0; j.M = B; j.N = HID; j.ldc = HID;
            if (kind == 0) {                        // target actor i on next_state_i
              const float* w = actor_p(a.T, i);
              j.A = a.bS2 + i * O; j.lda = SO; j.K = O; j.Bp = w + la.w1; j.ldb = O; j.bias = w + la.b1; j.C = a.s.ta[i].z1;
            } else if (kind == 1) {                 // critic target i: the observation columns (the actions follow in PH_CT_L2)
              const float* w = critic_p(a.T, i);
              j.A = a.bS2; j.lda = SO; j.K = SO; j.Bp = w + lc.w1; j.ldb = CI; j.bias = w + lc.b1; j.C = a.s.ct[i].z1;
            } else if (kind == 2) {                 // critic i on [state | action]
              const float* w = critic_p(a.P, i);
              j.A = a.bS; j.lda = SO; j.A2 = a.bACT; j.lda2 = SA; j.ksplit = SO; j.K = CI;
              j.Bp = w + lc.w1; j.ldb = CI; j.bias = w + lc.b1; j.C = a.s.c[i].z1;
            } else {                                // actor i on state_i
              const float* w = actor_p(a.P, i);
              j.A = a.bS + i * O; j.lda = SO; j.K = O; j.Bp = w + la.w1; j.ldb = O; j.bias = w + la.b1; j.C = a.s.ac[i].z1;
            }
            gemm_tile(j, mt, nt, smem, a.ln_eps);
          }
        } break;

        // ============================================================ second layers (LayerNorm + ReLU of the first as prologue)
        case PH_L2: case PH_C2_L2: case PH_CT_L2: {
          const int per = MT * (HID / TN);
          const int kinds = ph == PH_L2 ? 3 : 1;
          for (int t = blockIdx.x; t < kinds * n * per; t += n_ctas) {
            const int g = t / per, tile = t - g * per, mt = tile / (HID / TN), nt = tile % (HID / TN);
            const int kind = g / n, i = g - kind * n;
            GemmJob j;
            memset(&j, 0, sizeof(j));
            j.var = 0; j.M = B; j.N = HID; j.K = HID; j.lda = HID; j.ldb = HID; j.ldc = HID; j.pro = 1;
            const float* w;
            const NetLayout* L;
            const Pass* p;
            if (ph == PH_CT_L2) { w = critic_p(a.T, i); L = &lc; p = &a.s.ct[i];
              j.pro = 2; j.act = a.s.a2; j.n_act = SA; j.w_act = w + lc.w1 + SO; j.ldw = CI;
            } else if (ph == PH_C2_L2) { w = critic_p(a.P, i); L = &lc; p = &a.s.c2[i];
            } else if (kind == 0) { w = actor_p(a.T, i); L = &la; p = &a.s.ta[i];
            } else if (kind == 1) { w = critic_p(a.P, i); L = &lc; p = &a.s.c[i];
            } else { w = actor_p(a.P, i); L = &la; p = &a.s.ac[i]; }
            j.A = p->z1; j.Bp = w + L->w2; j.bias = w + L->b2; j.C = p->z2;
            j.gamma = w + L->b1 + HID; j.beta = w + L->b1 + 2 * HID; j.h_out = p->h1; j.st_out = p->st1;
            gemm_tile(j, mt, nt, smem, a.ln_eps);
          }
        } break;

        // ============================================================ output layers: target actions, Q, the actors' actions
        case PH_HEADS: {
          for (int t = blockIdx.x; t < 3 * n * NRB; t += n_ctas) {
            const int g = t / NRB, rb = t - g * NRB, kind = g / n, i = g - kind * n;
            const int row = rb * RB + warp;
            const bool is_critic = kind == 1;
            const float* w = kind == 0 ? actor_p(a.T, i) : kind == 1 ? critic_p(a.P, i) : actor_p(a.P, i);
            const NetLayout& L = is_critic ? lc : la;
            const Pass& p = kind == 0 ? a.s.ta[i] : kind == 1 ? a.s.c[i] : a.s.ac[i];
            float mu, rstd;
            const float4 xh = ln_row(ldcg4(p.z2 + (long long)row * HID + 4 * lane), a.ln_eps, mu, rstd);
            const float4 h = affine_relu(xh, ldcg4(w + L.b2 + HID + 4 * lane), ldcg4(w + L.b2 + 2 * HID + 4 * lane));
            st4(p.h2 + (long long)row * HID + 4 * lane, h);
            if (lane == 0) *reinterpret_cast<float2*>(p.st2 + 2 * row) = make_float2(mu, rstd);
            if (is_critic) {
              const float qv = warp_sum(dot4(h, ldcg4(w + L.w3 + 4 * lane))) + __ldcg(w + L.b3);
              if (lane == 0) a.s.q[i][row] = qv;
            } else {
              float logit[A], gn[A];
#pragma unroll
              for (int k = 0; k < A; ++k) logit[k] = warp_sum(dot4(h, ldcg4(w + L.w3 + k * HID + 4 * lane))) + __ldcg(w + L.b3 + k);
              const float* ext = kind == 0 ? a.gum_next : a.gum_cur;
              if (ext) {
#pragma unroll
                for (int k = 0; k < A; ++k) gn[k] = __ldg(ext + ((long long)row * n + i) * A + k);
              } else {
                gumbel_row(a, upd, row, i, kind == 0 ? 0 : 1, gn);
              }
              float mx = -3.4e38f;
#pragma unroll
              for (int k = 0; k < A; ++k) { logit[k] += gn[k]; mx = fmaxf(mx, logit[k]); }
              float den = 0.f;
#pragma unroll
              for (int k = 0; k < A; ++k) { logit[k] = expf(logit[k] - mx); den += logit[k]; }
              const float inv = 1.0f / den;
              float pl = 0.f;                           // lane k holds probability k
#pragma unroll
              for (int k = 0; k < A; ++k) pl = lane == k ? logit[k] * inv : pl;
              if (kind == 0) {
                if (lane < A) a.s.a2[(long long)row * SA + i * A + lane] = pl;
              } else {
                if (lane < A) a.s.anew[i][(long long)row * A + lane] = pl;
                // the critic input of the actor loss: the batch's actions with agent i's replaced by its actor's output
                for (int c0 = 0; c0 < SA; c0 += 32) {
                  const int c = c0 + lane, k = c - i * A;
                  const bool mine = k >= 0 && k < A;
                  const float pv = __shfl_sync(0xffffffffu, pl, mine ? k : 0);
                  if (c < SA) a.s.ax[i][(long long)row * SA + c] = mine ? pv : __ldcg(a.bACT + (long long)row * SA + c);
                }
              }
            }
          }
        } break;

        // ============================================================ TD target, critic loss, backward through layer 3 and LayerNorm 2
        case PH_TD: case PH_ALOSS: {
          const bool td = ph == PH_TD;
          if (tid == 0) {
            s_off[0] = 3 * HID; s_off[1] = 4 * HID; s_off[2] = 5 * HID; s_off[3] = 6 * HID;     // db2, dgamma2, dbeta2, dW3
            s_off[MAX_SLOTS] = 6 * HID + HID;                                                    // db3
          }
          for (int t = blockIdx.x; t < n * NRB; t += n_ctas) {
            const int i = t / NRB, rb = t - i * NRB, row = rb * RB + warp;
            const float* w = critic_p(a.P, i);
            const Pass& p = td ? a.s.c[i] : a.s.c2[i];
            float4 h, xh;
            float rstd, dqv, loss_part;
            const float4 w3 = ldcg4(w + lc.w3 + 4 * lane), g2 = ldcg4(w + lc.b2 + HID + 4 * lane);
            if (td) {
              // Q_target(next_state, target actions): the target critic's output layer
              const float* wt = critic_p(a.T, i);
              float mu_t, rs_t;
              const float4 xt = ln_row(ldcg4(a.s.ct[i].z2 + (long long)row * HID + 4 * lane), a.ln_eps, mu_t, rs_t);
              const float4 ht = affine_relu(xt, ldcg4(wt + lc.b2 + HID + 4 * lane), ldcg4(wt + lc.b2 + 2 * HID + 4 * lane));
              const float qn = warp_sum(dot4(ht, ldcg4(wt + lc.w3 + 4 * lane))) + __ldcg(wt + lc.b3);
              const float yv = __ldcg(a.bR + (long long)row * n + i) + a.gamma * (1.0f - __ldcg(a.bD + (long long)row * n + i)) * qn;
              const float diff = __ldcg(a.s.q[i] + row) - yv;
              dqv = 2.0f * diff / (float)B;                   // d mse / dq
              loss_part = diff * diff / (float)B;
              if (lane == 0) { a.s.y[i][row] = yv; a.s.dq[i][row] = dqv; }
              h = ldcg4(p.h2 + (long long)row * HID + 4 * lane);
              const float2 st = __ldcg(reinterpret_cast<const float2*>(p.st2 + 2 * row));
              rstd = st.y;
              xh = xhat_of(ldcg4(p.z2 + (long long)row * HID + 4 * lane), st.x, st.y);
            } else {
              // actor loss -mean Q(state, [actor_i(state_i), other actions]) through the UPDATED critic
              float mu;
              xh = ln_row(ldcg4(p.z2 + (long long)row * HID + 4 * lane), a.ln_eps, mu, rstd);
              h = affine_relu(xh, g2, ldcg4(w + lc.b2 + 2 * HID + 4 * lane));
              const float qv = warp_sum(dot4(h, w3)) + __ldcg(w + lc.b3);
              dqv = -1.0f / (float)B;
              loss_part = -qv / (float)B;
            }
            float4 dy;
            const float4 dz = ln_relu_bwd_row(scale4(w3, dqv), h, xh, g2, rstd, dy);
            st4(a.s.dz2[i] + (long long)row * HID + 4 * lane, dz);
            if (td) {
              float* pt = rs.part + (warp * 4) * HID + 4 * lane;
              st4(pt, dz); st4(pt + HID, mul4(dy, xh)); st4(pt + 2 * HID, dy); st4(pt + 3 * HID, scale4(h, dqv));
              if (lane == 0) { rs.sc[warp * 16] = dqv; rs.sc[warp * 16 + 1] = loss_part; }
            } else if (lane == 0) rs.sc[warp * 16] = loss_part;
            // losses go to lp (critics first, then actors); the critic's vector gradients to its partial-sum block
            __syncthreads();
            if (td) {
              flush_partials(rs, 4, s_off, a.s.pb[n + i] + (long long)rb * lc.small, 1, s_off + MAX_SLOTS);
              if (tid == 0) { float v = 0.f; for (int wv = 0; wv < WARPS; ++wv) v += rs.sc[wv * 16 + 1]; a.s.lp[i * NRB + rb] = v; }
            } else {
              if (tid == 0) { float v = 0.f; for (int wv = 0; wv < WARPS; ++wv) v += rs.sc[wv * 16]; a.s.lp[(n + i) * NRB + rb] = v; }
            }
            __syncthreads();
          }
        } break;

        // ============================================================ second-layer gradients: dW2 = dz2^T h1, dh1 = dz2 W2
        case PH_C_BWD2: case PH_A_BWD2: case PH_C2_DH1: {
          const int per = (HID / TM) * (HID / TN), per_x = MT * (HID / TN);
          const bool critic = ph != PH_A_BWD2;
          const int n_w = ph == PH_C2_DH1 ? 0 : n * per;           // weight-gradient tiles first, then input-gradient tiles
          for (int t = blockIdx.x; t < n_w + n * per_x; t += n_ctas) {
            GemmJob j;
            memset(&j, 0, sizeof(j));
            int i, mt, nt;
            const NetLayout& L = critic ? lc : la;
            if (t < n_w) {
              i = t / per; const int tile = t - i * per; mt = tile / (HID / TN); nt = tile % (HID / TN);
              const Pass& p = critic ? a.s.c[i] : a.s.ac[i];
              j.var = 2; j.M = HID; j.N = HID; j.K = B;
              j.A = critic ? a.s.dz2[i] : a.s.adz2[i]; j.lda = HID; j.Bp = p.h1; j.ldb = HID;
              j.C = a.G + a.net_off[critic ? n + i : i] + L.w2; j.ldc = HID;
            } else {
              const int tt = t - n_w; i = tt / per_x; const int tile = tt - i * per_x; mt = tile / (HID / TN); nt = tile % (HID / TN);
              const float* w = critic ? critic_p(a.P, i) : actor_p(a.P, i);
              j.var = 1; j.M = B; j.N = HID; j.K = HID;
              j.A = critic ? a.s.dz2[i] : a.s.adz2[i]; j.lda = HID; j.Bp = w + L.w2; j.ldb = HID;
              j.C = critic ? a.s.dh1[i] : a.s.adh1[i]; j.ldc = HID;
            }
            gemm_tile(j, mt, nt, smem, a.ln_eps);
          }
        } break;

        // ============================================================ LayerNorm 1 backward (+ its vector gradients)
        case PH_C_LN1: case PH_A_LN1: {
          const bool critic = ph == PH_C_LN1;
          if (tid == 0) { s_off[0] = 0; s_off[1] = HID; s_off[2] = 2 * HID; }                   // db1, dgamma1, dbeta1
          for (int t = blockIdx.x; t < n * NRB; t += n_ctas) {
            const int i = t / NRB, rb = t - i * NRB, row = rb * RB + warp;
            const NetLayout& L = critic ? lc : la;
            const float* w = critic ? critic_p(a.P, i) : actor_p(a.P, i);
            const Pass& p = critic ? a.s.c[i] : a.s.ac[i];
            const float2 st = __ldcg(reinterpret_cast<const float2*>(p.st1 + 2 * row));
            const float4 xh = xhat_of(ldcg4(p.z1 + (long long)row * HID + 4 * lane), st.x, st.y);
            const float4 h = ldcg4(p.h1 + (long long)row * HID + 4 * lane);
            const float4 dh = ldcg4((critic ? a.s.dh1[i] : a.s.adh1[i]) + (long long)row * HID + 4 * lane);
            float4 dy;
            const float4 dz = ln_relu_bwd_row(dh, h, xh, ldcg4(w + L.b1 + HID + 4 * lane), st.y, dy);
            st4((critic ? a.s.dz1[i] : a.s.adz1[i]) + (long long)row * HID + 4 * lane, dz);
            float* pt = rs.part + (warp * 3) * HID + 4 * lane;
            st4(pt, dz); st4(pt + HID, mul4(dy, xh)); st4(pt + 2 * HID, dy);
            flush_partials(rs, 3, s_off, a.s.pb[critic ? n + i : i] + (long long)rb * L.small, 0, nullptr);
          }
        } break;

        // ============================================================ first-layer weight gradients: dW1 = dz1^T x
        case PH_C_DW1: case PH_A_DW1: {
          const bool critic = ph == PH_C_DW1;
          const int N = critic ? CI : O, ntn = (N + TN - 1) / TN, per = (HID / TM) * ntn;
          for (int t = blockIdx.x; t < n * per; t += n_ctas) {
            const int i = t / per, tile = t - i * per, mt = tile / ntn, nt = tile % ntn;
            GemmJob j;
            memset(&j, 0, sizeof(j));
            j.var = 2; j.M = HID; j.N = N; j.K = B; j.lda = HID;
            if (critic) {
              j.A = a.s.dz1[i]; j.Bp = a.bS; j.ldb = SO; j.B2 = a.bACT; j.ldb2 = SA; j.nsplit = SO;
              j.C = a.G + a.net_off[n + i] + lc.w1; j.ldc = CI;
            } else {
              j.A = a.s.adz1[i]; j.Bp = a.bS + i * O; j.ldb = SO;
              j.C = a.G + a.net_off[i] + la.w1; j.ldc = O;
            }
            gemm_tile(j, mt, nt, smem, a.ln_eps);
          }
        } break;

        // ============================================================ Adam + soft update (torch.optim.Adam arithmetic; TAU)
        case PH_ADAM_C: case PH_ADAM_A: {
          const bool critic = ph == PH_ADAM_C;
          const NetLayout& L = critic ? lc : la;
          const float lr = critic ? a.lr_c : a.lr_a;
          const int mode = a.adam_mode[critic ? 0 : 1];
          for (int i = 0; i < n; ++i) {
            const int net = critic ? n + i : i;
            const long long off = a.net_off[net];
            const float stepf = step0[net] + (float)(u + 1);
            const float bc1 = 1.0f - powf(a.beta1, stepf), bc2s = sqrtf(1.0f - powf(a.beta2, stepf));
            const float step_size = lr / bc1;
            const float* pb = a.s.pb[net];
            for (int idx = blockIdx.x * THREADS + tid; idx < L.total; idx += n_ctas * THREADS) {
              const int ci = compact_index(L, idx);
              float g;
              if (ci < 0 || (mode & ADAM_FROM_G)) {           // a matrix gradient, or vector gradients finalised by an earlier launch
                g = __ldcg(a.G + off + idx);
              } else {
                g = 0.f;
                for (int rb = 0; rb < NRB; ++rb) g += __ldcg(pb + (long long)rb * L.small + ci);
              }
              if (mode & ADAM_WRITE_G) a.G[off + idx] = g;
              if (mode & ADAM_APPLY) {
                g *= a.grad_scale;
                float m = __ldcg(a.M + off + idx), v = __ldcg(a.V + off + idx), p = __ldcg(a.P + off + idx);
                m = m + (1.0f - a.beta1) * (g - m);
                v = a.beta2 * v + (1.0f - a.beta2) * g * g;
                const float denom = sqrtf(v) / bc2s + a.eps;
                p -= step_size * (m / denom);
                a.M[off + idx] = m; a.V[off + idx] = v; a.P[off + idx] = p;
                const float tg = __ldcg(a.T + off + idx);
                a.T[off + idx] = tg + a.tau * (p - tg);
              }
            }
            if (blockIdx.x == 0 && tid == 0 && a.losses) {
              float v = 0.f;
              const float* lp = a.s.lp + (critic ? i : n + i) * NRB;
              for (int rb = 0; rb < NRB; ++rb) v += __ldcg(lp + rb);
              a.losses[((long long)u * 2 + (critic ? 1 : 0)) * n + i] = v;
            }
          }
        } break;

        // ============================================================ actor loss: the updated critic on [state | ax_i]
        case PH_C2_L1: {
          const int per = MT * (HID / TN);
          for (int t = blockIdx.x; t < n * per; t += n_ctas) {
            const int i = t / per, tile = t - i * per, mt = tile / (HID / TN), nt = tile % (HID / TN);
            const float* w = critic_p(a.P, i);
            GemmJob j;
            memset(&j, 0, sizeof(j));
            j.var = 0; j.M = B; j.N = HID; j.K = CI; j.ldc = HID;
            j.A = a.bS; j.lda = SO; j.A2 = a.s.ax[i]; j.lda2 = SA; j.ksplit = SO;
            j.Bp = w + lc.w1; j.ldb = CI; j.bias = w + lc.b1; j.C = a.s.c2[i].z1;
            gemm_tile(j, mt, nt, smem, a.ln_eps);
          }
        } break;

        // ============================================================ critic LayerNorm 1 backward -> d action_i -> softmax -> actor layer 3 + LayerNorm 2 backward
        case PH_ACT_BWD: {
          if (tid == 0) {
            s_off[0] = 3 * HID; s_off[1] = 4 * HID; s_off[2] = 5 * HID;                          // db2, dgamma2, dbeta2 (actor)
            for (int k = 0; k < A; ++k) { s_off[3 + k] = 6 * HID + k * HID; s_off[MAX_SLOTS + k] = 6 * HID + A * HID + k; }
          }
          for (int t = blockIdx.x; t < n * NRB; t += n_ctas) {
            const int i = t / NRB, rb = t - i * NRB, row = rb * RB + warp;
            const float* wc = critic_p(a.P, i);
            const float* wa = actor_p(a.P, i);
            // the critic's first-layer columns of agent i's action, transposed: wa_s[k][o] = W1c[o][SO + i A + k]
            __syncthreads();
            for (int e = tid; e < A * HID; e += THREADS) {
              const int o = e / A, k = e - o * A;
              rs.wa[k * HID + o] = __ldcg(wc + lc.w1 + (long long)o * CI + SO + i * A + k);
            }
            __syncthreads();
            const Pass& pc = a.s.c2[i];
            const float2 st1 = __ldcg(reinterpret_cast<const float2*>(pc.st1 + 2 * row));
            const float4 xh1 = xhat_of(ldcg4(pc.z1 + (long long)row * HID + 4 * lane), st1.x, st1.y);
            float4 dy;
            const float4 dz1 = ln_relu_bwd_row(ldcg4(a.s.dh1[i] + (long long)row * HID + 4 * lane),
                                               ldcg4(pc.h1 + (long long)row * HID + 4 * lane), xh1,
                                               ldcg4(wc + lc.b1 + HID + 4 * lane), st1.y, dy);
            float da[A], an[A], dl[A];
            float sdot = 0.f;
#pragma unroll
            for (int k = 0; k < A; ++k) {
              da[k] = warp_sum(dot4(dz1, *reinterpret_cast<const float4*>(rs.wa + k * HID + 4 * lane)));
              an[k] = __ldcg(a.s.anew[i] + (long long)row * A + k);
              sdot += an[k] * da[k];
            }
#pragma unroll
            for (int k = 0; k < A; ++k) dl[k] = an[k] * (da[k] - sdot);      // softmax backward (the Gumbel noise is a constant)
            // actor layer 3: dW3 += dl^T h2, db3 += dl, dh2 = dl W3
            const Pass& pa = a.s.ac[i];
            const float4 h2 = ldcg4(pa.h2 + (long long)row * HID + 4 * lane);
            float4 dh2 = make_float4(0.f, 0.f, 0.f, 0.f);
            float* pt = rs.part + (warp * (3 + A)) * HID + 4 * lane;
#pragma unroll
            for (int k = 0; k < A; ++k) {
              const float4 w3 = ldcg4(wa + la.w3 + k * HID + 4 * lane);
              dh2.x = fmaf(dl[k], w3.x, dh2.x); dh2.y = fmaf(dl[k], w3.y, dh2.y);
              dh2.z = fmaf(dl[k], w3.z, dh2.z); dh2.w = fmaf(dl[k], w3.w, dh2.w);
              st4(pt + (3 + k) * HID, scale4(h2, dl[k]));
              if (lane == 0) rs.sc[warp * 16 + k] = dl[k];
            }
            const float2 st2 = __ldcg(reinterpret_cast<const float2*>(pa.st2 + 2 * row));
            const float4 xh2 = xhat_of(ldcg4(pa.z2 + (long long)row * HID + 4 * lane), st2.x, st2.y);
            const float4 dz2 = ln_relu_bwd_row(dh2, h2, xh2, ldcg4(wa + la.b2 + HID + 4 * lane), st2.y, dy);
            st4(a.s.adz2[i] + (long long)row * HID + 4 * lane, dz2);
            st4(pt, dz2); st4(pt + HID, mul4(dy, xh2)); st4(pt + 2 * HID, dy);
            flush_partials(rs, 3 + A, s_off, a.s.pb[i] + (long long)rb * la.small, A, s_off + MAX_SLOTS);
          }
        } break;
        default: break;
      }
      if (!(u == a.updates - 1 && ph == a.ph_end - 1)) grid_barrier(a.s.bar, n_ctas);
    }
  }
  // step counters of the optimisers whose Adam phase ran in this launch
  if (blockIdx.x == 0 && tid == 0) {
    a.s.trace[a.ph_end] = phase_clock();
    for (int k = 0; k < 2 * n; ++k) {
      const int ph = k < n ? PH_ADAM_A : PH_ADAM_C;
      if (ph >= a.ph_begin && ph < a.ph_end && (a.adam_mode[k < n ? 1 : 0] & ADAM_APPLY)) a.steps[k] = step0[k] + (float)a.updates;
    }
  }
}

}  // namespace gwl


namespace {

int check_cfg(const gw_learner_config* c) {
  if (!c || c->struct_size != sizeof(gw_learner_config)) return GW_EINVAL;
  if (c->n_agents < 1 || c->n_agents > GW_MAX_LEARNERS) return GW_EINVAL;
  if (c->obs_len < 16 || c->obs_len % 16 != 0 || (c->n_agents * c->obs_len) % 32 != 0) return GW_EINVAL;
  if (c->action_dim != gwl::NA) return GW_EINVAL;
  if (c->batch < 32 || c->batch % 32 != 0 || c->batch > 512) return GW_EINVAL;
  return GW_OK;
}

inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// carve the scratch block; base == nullptr only measures
int64_t carve(const gw_learner_config& c, char* base, gwl::Scratch* s, std::map<std::string, std::pair<float*, int64_t>>* dbg,
              float** cluster = nullptr) {
  const int n = c.n_agents, B = c.batch, O = c.obs_len, A = c.action_dim, H = gwl::HID;
  const gwl::NetLayout la = gwl::make_layout(O, A), lc = gwl::make_layout(n * (O + A), 1);
  int64_t off = 0;
  auto take = [&](int64_t floats, const char* name, int idx) -> float* {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += round_up(floats * 4, 256);
    if (dbg && name) (*dbg)[std::string(name) + "#" + std::to_string(idx)] = {p, floats};
    return p;
  };
  gwl::Scratch t;
  std::memset(&t, 0, sizeof(t));
  t.bar = reinterpret_cast<unsigned*>(take(64, nullptr, 0));
  t.trace = reinterpret_cast<unsigned long long*>(take(2 * 64, "trace", 0));
  t.S = take((int64_t)B * n * O, "S", 0); t.S2 = take((int64_t)B * n * O, "S2", 0);
  t.ACT = take((int64_t)B * n * A, "ACT", 0); t.R = take((int64_t)B * n, "R", 0); t.D = take((int64_t)B * n, "D", 0);
  const char* names[5] = {"ta", "ct", "c", "ac", "c2"};
  gwl::Pass* passes[5] = {t.ta, t.ct, t.c, t.ac, t.c2};
  for (int k = 0; k < 5; ++k)
    for (int i = 0; i < n; ++i) {
      gwl::Pass& p = passes[k][i];
      const std::string nm = names[k];
      p.z1 = take((int64_t)B * H, (nm + ".z1").c_str(), i); p.h1 = take((int64_t)B * H, (nm + ".h1").c_str(), i);
      p.st1 = take((int64_t)B * 2, (nm + ".st1").c_str(), i);
      p.z2 = take((int64_t)B * H, (nm + ".z2").c_str(), i); p.h2 = take((int64_t)B * H, (nm + ".h2").c_str(), i);
      p.st2 = take((int64_t)B * 2, (nm + ".st2").c_str(), i);
    }
  t.a2 = take((int64_t)B * n * A, "a2", 0);
  for (int i = 0; i < n; ++i) {
    t.anew[i] = take((int64_t)B * A, "anew", i); t.ax[i] = take((int64_t)B * n * A, "ax", i);
    t.q[i] = take(B, "q", i); t.y[i] = take(B, "y", i); t.dq[i] = take(B, "dq", i);
    t.dz2[i] = take((int64_t)B * H, "dz2", i); t.dh1[i] = take((int64_t)B * H, "dh1", i); t.dz1[i] = take((int64_t)B * H, "dz1", i);
    t.adz2[i] = take((int64_t)B * H, "adz2", i); t.adh1[i] = take((int64_t)B * H, "adh1", i); t.adz1[i] = take((int64_t)B * H, "adz1", i);
    t.pb[i] = take((int64_t)(B / gwl::RB) * la.small, nullptr, 0);
    t.pb[n + i] = take((int64_t)(B / gwl::RB) * lc.small, nullptr, 0);
  }
  t.lp = take((int64_t)2 * n * (B / gwl::RB), nullptr, 0);
  float* cl = take(gwc_scratch_floats(c) + 4, nullptr, 0);
  if (cluster) *cluster = cl;
  if (s) *s = t;
  return off;
}

size_t smem_need(const gw_learner_config& c) {
  const int n = c.n_agents, B = c.batch, CI = n * (c.obs_len + c.action_dim);
  const size_t red = (size_t)gwl::WARPS * gwl::TM * gwl::RED_LD;
  const size_t fwd = 2 * (size_t)gwl::TM * gwl::pad_k(CI) + red;                       // var 0, K = CI
  const size_t dx = (size_t)gwl::TM * gwl::pad_k(gwl::HID) + (size_t)gwl::HID * gwl::TN + red;
  const size_t dw = 2 * (size_t)B * gwl::TM + red;                                     // var 2, K = B
  const size_t row = (size_t)gwl::WARPS * gwl::MAX_SLOTS * gwl::HID + gwl::WARPS * 16 + (size_t)c.action_dim * gwl::HID;
  size_t m = fwd;
  if (dx > m) m = dx;
  if (dw > m) m = dw;
  if (row > m) m = row;
  return m * sizeof(float);
}

}  // namespace

extern "C" int gw_learner_layout_of(const gw_learner_config* cfg, gw_learner_layout* out) {
  if (!out || out->struct_size != sizeof(gw_learner_layout)) return GW_EINVAL;
  if (int rc = check_cfg(cfg)) return rc;
  if (smem_need(*cfg) > (size_t)227 * 1024) return GW_EINVAL;      // the critic's input row (n * (obs_len + 9) floats) must fit the tiles' shared-memory staging
  const int n = cfg->n_agents;
  const gwl::NetLayout la = gwl::make_layout(cfg->obs_len, cfg->action_dim);
  const gwl::NetLayout lc = gwl::make_layout(n * (cfg->obs_len + cfg->action_dim), 1);
  std::memset(out, 0, sizeof(*out));
  out->struct_size = sizeof(gw_learner_layout);
  out->n_nets = 2 * n;
  int64_t off = 0;
  for (int k = 0; k < 2 * n; ++k) {
    out->net_offset[k] = off;
    out->net_params[k] = k < n ? la.total : lc.total;
    off += round_up(out->net_params[k], 4);
  }
  out->param_floats = off;
  out->scratch_bytes = carve(*cfg, nullptr, nullptr, nullptr);
  return GW_OK;
}

extern "C" int gw_learner_create(gw_handle* h, const gw_learner_config* cfg, const gw_learner_buffers* buf, gw_learner** out) {
  if (h == nullptr || out == nullptr) return GW_EINVAL;
  if (check_cfg(cfg) != GW_OK) return gw_fail(h, GW_EINVAL, "gw_learner_create: bad config (n_agents 1..2, obs_len % 16, action_dim 9, batch % 32, batch <= 512)");
  if (!buf || !buf->params || !buf->targets || !buf->adam_m || !buf->adam_v || !buf->grads || !buf->adam_steps || !buf->scratch)
    return gw_fail(h, GW_EINVAL, "gw_learner_create: buffer missing");
  for (const void* p : {(const void*)buf->params, (const void*)buf->targets, (const void*)buf->adam_m, (const void*)buf->adam_v,
                        (const void*)buf->grads})
    if (reinterpret_cast<uintptr_t>(p) & 15) return gw_fail(h, GW_EINVAL, "gw_learner_create: parameter vectors must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(buf->scratch) & 255) return gw_fail(h, GW_EINVAL, "gw_learner_create: scratch must be 256-byte aligned");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  gw_learner* l = new gw_learner();
  l->h = h; l->cfg = *cfg; l->buf = *buf;
  l->lay.struct_size = sizeof(gw_learner_layout);
  gw_learner_layout_of(cfg, &l->lay);
  gwl::LearnArgs& a = l->args;
  std::memset(&a, 0, sizeof(a));
  a.n = cfg->n_agents; a.O = cfg->obs_len; a.A = cfg->action_dim; a.B = cfg->batch;
  a.SO = a.n * a.O; a.SA = a.n * a.A; a.CI = a.SO + a.SA;
  a.la = gwl::make_layout(a.O, a.A); a.lc = gwl::make_layout(a.CI, 1);
  for (int k = 0; k < 2 * a.n; ++k) a.net_off[k] = l->lay.net_offset[k];
  a.P = buf->params; a.T = buf->targets; a.M = buf->adam_m; a.V = buf->adam_v; a.G = buf->grads; a.steps = buf->adam_steps;
  carve(*cfg, static_cast<char*>(buf->scratch), &a.s, &l->dbg, &l->cluster_scratch);
  a.gk0 = (uint32_t)cfg->seed; a.gk1 = (uint32_t)(cfg->seed >> 32) ^ 0x47554D42u;     // "GUMB"
  a.lr_a = cfg->lr_actor; a.lr_c = cfg->lr_critic; a.gamma = cfg->gamma; a.tau = cfg->tau;
  a.beta1 = cfg->beta1; a.beta2 = cfg->beta2; a.eps = cfg->adam_eps; a.ln_eps = cfg->ln_eps;
  l->smem = smem_need(*cfg);
  cudaError_t e = cudaFuncSetAttribute(gwl::gw_learn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l->smem);
  if (e != cudaSuccess) { delete l; return gw_cuda_fail(h, e, "cudaFuncSetAttribute(gw_learn_kernel)"); }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gwl::gw_learn_kernel, gwl::THREADS, l->smem);
  if (e != cudaSuccess || per_sm < 1) { delete l; return gw_fail(h, GW_ECUDA, "gw_learner_create: the update kernel does not fit on an SM"); }
  l->grid = h->sm_count;                             // one CTA per SM, all co-resident (cooperative launch)
  l->cluster_grid = gwc_prepare(l);                  // the fast path, where the shape and the device allow it
  *out = l;
  return GW_OK;
}

extern "C" int gw_learner_destroy(gw_learner* l) {
  if (l) {
    for (int q = 0; q < GW_MAX_PEERS; ++q) {
      if (!l->peer_base[q]) continue;
      if (q == l->peer_rank) cudaFree(l->peer_base[q]); else cudaIpcCloseMemHandle(l->peer_base[q]);
    }
  }
  delete l;
  return GW_OK;
}

// ---- gradient exchange of a data-parallel run over NVLink peer memory (one process per GPU, CUDA IPC)
extern "C" int gw_learner_peer_export(gw_learner* l, gw_peer_handle* out) {
  if (!l || !out) return GW_EINVAL;
  gw_handle* h = l->h;
  static_assert(sizeof(gw_peer_handle) == sizeof(cudaIpcMemHandle_t), "gw_peer_handle carries a cudaIpcMemHandle_t");
  if (gw_learner_kernel(l) != GW_LEARN_KERNEL_CLUSTER) return gw_fail(h, GW_ESTATE, "gw_learner_peer_export: the exchange lives in the cluster kernel");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  if (!l->peer_base[l->peer_rank]) {
    l->peer_flag_off = (size_t)round_up(l->lay.param_floats * 8, 256) * (GW_MAX_PEERS + 1);      // one slot of flag-in-data lines per sending rank + one for returned sums ({value, epoch} pairs: 8 bytes per float), then the error word
    void* p = nullptr;
    const size_t bytes = l->peer_flag_off + 8 * GW_MAX_PEERS + 256;
    GW_CUDA(h, cudaMalloc(&p, bytes));                 // a whole allocation of its own: what CUDA IPC can share
    GW_CUDA(h, cudaMemset(p, 0, bytes));
    l->peer_base[l->peer_rank] = p;
  }
  cudaIpcMemHandle_t ipc;
  GW_CUDA(h, cudaIpcGetMemHandle(&ipc, l->peer_base[l->peer_rank]));
  std::memcpy(out->bytes, &ipc, sizeof(ipc));
  return GW_OK;
}

extern "C" int gw_learner_peer_connect(gw_learner* l, int32_t rank, int32_t world, const gw_peer_handle* handles) {
  if (!l || !handles) return GW_EINVAL;
  gw_handle* h = l->h;
  if (world < 2 || world > GW_MAX_PEERS || rank < 0 || rank >= world) return gw_fail(h, GW_EINVAL, "gw_learner_peer_connect: world 2..8, 0 <= rank < world");
  if (!l->peer_base[l->peer_rank]) return gw_fail(h, GW_ESTATE, "gw_learner_peer_connect: call gw_learner_peer_export first");
  if (l->peer_world > 1) return gw_fail(h, GW_ESTATE, "gw_learner_peer_connect: already connected");
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  void* own = l->peer_base[l->peer_rank];
  l->peer_base[l->peer_rank] = nullptr;
  l->peer_rank = rank;
  l->peer_base[rank] = own;
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, handles[q].bytes, sizeof(ipc));
    void* p = nullptr;
    GW_CUDA(h, cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    l->peer_base[q] = p;
  }
  l->peer_world = world;
  return GW_OK;
}

extern "C" int gw_learner_peer_disable(gw_learner* l) {      // back to exchanges done by the caller (a peer failed to connect)
  if (!l) return GW_EINVAL;
  l->peer_world = 1;
  return GW_OK;
}

extern "C" int gw_learner_peer_status(gw_learner* l, int32_t* world, int32_t* timed_out) {
  if (!l) return GW_EINVAL;
  if (world) *world = l->peer_world;
  if (timed_out) {
    *timed_out = 0;
    if (l->peer_world > 1) {
      unsigned int e = 0;
      GW_CUDA(l->h, cudaMemcpy(&e, static_cast<char*>(l->peer_base[l->peer_rank]) + l->peer_flag_off, 4, cudaMemcpyDeviceToHost));
      *timed_out = (int32_t)e;
    }
  }
  return GW_OK;
}

extern "C" int gw_learner_set_kernel(gw_learner* l, int32_t kind) {
  if (!l) return GW_EINVAL;
  if (kind < GW_LEARN_KERNEL_AUTO || kind > GW_LEARN_KERNEL_CLUSTER) return gw_fail(l->h, GW_EINVAL, "gw_learner_set_kernel: bad kind");
  if (kind == GW_LEARN_KERNEL_CLUSTER && l->cluster_grid == 0)
    return gw_fail(l->h, GW_EINVAL, "gw_learner_set_kernel: the cluster kernel needs 2 learners, obs_len 160, batch % 16 == 0 and 2 * batch / 16 co-resident "
                   "clusters of 4 CTAs (this device: " + std::to_string(l->cluster_max_active) + ")");
  l->kernel_kind = kind;
  return GW_OK;
}

extern "C" int gw_learner_kernel(const gw_learner* l) {
  if (!l) return GW_EINVAL;
  if (l->kernel_kind == GW_LEARN_KERNEL_AUTO) return l->cluster_grid > 0 ? GW_LEARN_KERNEL_CLUSTER : GW_LEARN_KERNEL_PHASE;
  return l->kernel_kind;
}

extern "C" int gw_learner_debug_ptr(gw_learner* l, const char* name, int index, float** ptr, int64_t* floats) {
  if (!l || !name || !ptr) return GW_EINVAL;
  auto it = l->dbg.find(std::string(name) + "#" + std::to_string(index));
  if (it == l->dbg.end()) return gw_fail(l->h, GW_EINVAL, std::string("gw_learner_debug_ptr: no tensor named ") + name);
  *ptr = it->second.first;
  if (floats) *floats = it->second.second;
  return GW_OK;
}

extern "C" int gw_learner_update(gw_learner* l, const gw_learn_batch* batch, const gw_replay_view* ring, int64_t t_now,
                                 uint64_t sample_seed, uint64_t draw_base, int32_t updates, int32_t segment, float grad_scale,
                                 float* losses, void* stream) {
  if (l == nullptr) return GW_EINVAL;
  gw_handle* h = l->h;
  if ((batch == nullptr) == (ring == nullptr)) return gw_fail(h, GW_EINVAL, "gw_learner_update: give either a batch or a ring");
  if (updates < 1) return gw_fail(h, GW_EINVAL, "gw_learner_update: updates < 1");
  if (segment < GW_LEARN_ALL || segment > GW_LEARN_FINISH) return gw_fail(h, GW_EINVAL, "gw_learner_update: bad segment");
  if ((batch != nullptr || segment != GW_LEARN_ALL) && updates != 1)
    return gw_fail(h, GW_EINVAL, "gw_learner_update: an explicit batch or a segment is one update per call");
  if (int rc = gw_server_stop(h)) return rc;         // a resident step kernel would hold this stream
  gwl::LearnArgs a = l->args;
  if (batch) {
    if (!batch->state || !batch->action || !batch->reward || !batch->next_state || !batch->done)
      return gw_fail(h, GW_EINVAL, "gw_learner_update: batch pointer missing");
    for (const void* p : {(const void*)batch->state, (const void*)batch->next_state})
      if (reinterpret_cast<uintptr_t>(p) & 15) return gw_fail(h, GW_EINVAL, "gw_learner_update: state tensors must be 16-byte aligned");
    a.bS = batch->state; a.bS2 = batch->next_state; a.bACT = batch->action; a.bR = batch->reward; a.bD = batch->done;
    a.gum_next = batch->gumbel_next; a.gum_cur = batch->gumbel_cur;
    a.sample = 0;
  } else {
    const gw_replay_view& r = *ring;
    if (r.struct_size != sizeof(gw_replay_view)) return gw_fail(h, GW_EINVAL, "gw_learner_update: ring view of another size");
    if (r.n_learners != a.n || r.obs_len != a.O || r.action_dim != a.A || r.slots < 3 || r.num_envs < 1)
      return gw_fail(h, GW_EINVAL, "gw_learner_update: ring shape does not match the learner");
    if (r.obs_dtype != GW_OBS_F32 && r.obs_dtype != GW_OBS_BF16) return gw_fail(h, GW_EINVAL, "gw_learner_update: bad obs_dtype");
    if (!r.obs || !r.final_obs || !r.action || !r.reward || !r.terminated || !r.ended)
      return gw_fail(h, GW_EINVAL, "gw_learner_update: ring pointer missing");
    if ((reinterpret_cast<uintptr_t>(r.obs) | reinterpret_cast<uintptr_t>(r.final_obs)) & 15)
      return gw_fail(h, GW_EINVAL, "gw_learner_update: ring observations must be 16-byte aligned");
    a.ring = r; a.sample = 1; a.t_now = t_now;
    a.n_valid = t_now < r.slots - 1 ? t_now : r.slots - 1;
    if (a.n_valid < 1) return gw_fail(h, GW_ESTATE, "gw_learner_update: the ring is empty");
    gwr::sample_key(sample_seed, a.rk0, a.rk1);
    a.draw_base = draw_base;
    a.bS = a.s.S; a.bS2 = a.s.S2; a.bACT = a.s.ACT; a.bR = a.s.R; a.bD = a.s.D;
    a.gum_next = a.gum_cur = nullptr;
  }
  a.updates = updates;
  a.upd_base = l->updates_done;
  a.grad_scale = grad_scale;
  a.losses = losses;
  switch (segment) {
    case GW_LEARN_ALL:
      a.ph_begin = 0; a.ph_end = gwl::PH_COUNT;
      a.adam_mode[0] = a.adam_mode[1] = gwl::ADAM_WRITE_G | gwl::ADAM_APPLY;
      break;
    case GW_LEARN_CRITIC_GRADS:
      a.ph_begin = 0; a.ph_end = gwl::PH_ADAM_C + 1;
      a.adam_mode[0] = gwl::ADAM_WRITE_G; a.adam_mode[1] = 0;
      break;
    case GW_LEARN_ACTOR_GRADS:
      a.ph_begin = gwl::PH_ADAM_C; a.ph_end = gwl::PH_ADAM_A + 1;
      a.adam_mode[0] = gwl::ADAM_FROM_G | gwl::ADAM_APPLY; a.adam_mode[1] = gwl::ADAM_WRITE_G;
      break;
    default:
      a.ph_begin = gwl::PH_ADAM_A; a.ph_end = gwl::PH_ADAM_A + 1;
      a.adam_mode[0] = 0; a.adam_mode[1] = gwl::ADAM_FROM_G | gwl::ADAM_APPLY;
      break;
  }
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  if (gw_learner_kernel(l) == GW_LEARN_KERNEL_CLUSTER) {
    if (int rc = gwc_launch(l, a, segment, static_cast<cudaStream_t>(stream))) return rc;
    h->launches += 1;
    if (segment == GW_LEARN_ALL || segment == GW_LEARN_FINISH) l->updates_done += (unsigned long long)updates;
    return GW_OK;
  }
  void* params[] = {&a};
  GW_CUDA(h, cudaLaunchCooperativeKernel((const void*)gwl::gw_learn_kernel, dim3((unsigned)l->grid), dim3(gwl::THREADS), params,
                                         l->smem, static_cast<cudaStream_t>(stream)));
  h->launches += 1;
  if (segment == GW_LEARN_ALL || segment == GW_LEARN_FINISH) l->updates_done += (unsigned long long)updates;
  return GW_OK;
}
