// Fused MADDPG update, fast path (maddpg/agent.py:209-224): the same update as gw_maddpg.cu, reorganised so
// that a whole forward / backward chain needs NO grid-wide synchronisation.
//
// Decomposition.  Every step of the update except the weight-gradient sums is row-local (a batch row never meets another
// row), so a CLUSTER of 4 CTAs owns 16 batch rows of one agent and carries them through the entire chain:
//     phase A   target actors -> target critic -> TD target ; critic forward -> critic backward          (per cluster)
//     Adam C    sum the row blocks' partial gradients in a fixed order, Adam, soft update                 (whole grid)
//     phase B   actor forward -> updated critic forward -> -Q backward to the action -> actor backward   (per cluster)
//     Adam A                                                                                             (whole grid)
// Four grid barriers per update (the phase kernel needs nineteen).  Inside a cluster, CTA c computes output columns
// [32c, 32c+32) of every 128-wide layer -- it only ever loads ITS 32-row slice of each weight matrix (prefetched with
// cp.async through a 3-slot ring, two packages ahead of their use) -- and the full rows the next layer needs are
// rebuilt by an all-gather through distributed shared memory (st.shared::cluster into the 3 peers, one cluster barrier
// per layer).  LayerNorm, ReLU, the 9-/1-wide output layers, Gumbel-softmax, TD target and the LayerNorm backward passes
// run redundantly in every CTA on the full rows (a few thousand flops).  Weight gradients are rank-16 updates per row
// block; each cluster writes its partial gradient to its own slab, the Adam phase adds the slabs in order: deterministic.
//
// Arithmetic.  The 16 x 16 x K tiles run on the tensor cores as 3xTF32 (mma.sync.m16n8k8, each fp32 operand split into
// a tf32 head and a tf32 tail, three MMAs, fp32 accumulation): fp32-level accuracy (the tests hold gradients to 1e-4 of
// fp32 autograd), a fraction of the shared-memory traffic of an FMA tile.  tcgen05 needs M >= 64 rows per tile, which
// would put a whole layer on four SMs; the update is latency-bound, not FLOP-bound, so the 16-row tile is the better fit.
//
// Supported shape: the reference's (2 learners, obs_len 160, 9 actions, hidden 128), BATCH_SIZE a multiple of 16 with
// 2 * BATCH_SIZE / 16 clusters of 4 co-resident (a B200 holds 33: BATCH_SIZE <= 256; 8-CTA clusters would halve the
// work per CTA, but only 15 of them fit at once and BATCH_SIZE 128 needs 16).  Anything else runs on the phase kernel.
#include "gw_maddpg.cuh"

namespace gwc {
using namespace gwl;

constexpr int CL = 4;                       // CTAs per cluster
constexpr int RR = 16;                      // batch rows per cluster
constexpr int CW = HID / CL;                // output columns per CTA
constexpr int N2 = 2, O = 160, SO = N2 * O, SA = N2 * NA, CI = SO + SA;
constexpr int ALD = 136;                    // activation row stride (== 8 mod 32: conflict-free 64-bit fragment loads)
constexpr int XLD = 344;                    // critic input row stride (== 24 mod 32)
constexpr int SLOT_W = 5456, SLOT_V = 432, SLOT = SLOT_W + SLOT_V, NSLOT = 3, WLEAD = 4;
constexpr int RLD = 24, OLD = 36, TLD = 36; // reduction tile / reduced tile / transposed weight slice strides
constexpr int AXLD = 24;
// dynamic shared memory (floats)
constexpr int SM_X = 0, SM_X2 = SM_X + RR * XLD, SM_EX = SM_X2 + RR * XLD, SM_PV = SM_EX + 2 * 3 * RR * ALD;
constexpr int SM_RING = SM_PV + 3 * RR * ALD, SM_RED = SM_RING + NSLOT * SLOT, SM_OUT = SM_RED + WARPS * RR * RLD;
constexpr int SM_PERS = SM_OUT + RR * OLD, PERS_FLOATS = 3360, SM_MISC = SM_PERS + PERS_FLOATS, MISC_FLOATS = 2432;
constexpr int SM_TOTAL = SM_MISC + MISC_FLOATS;
static_assert(SM_TOTAL * 4 <= 227 * 1024, "shared memory plan");
static_assert(2 * RR * ALD <= RR * XLD, "two activation buffers live where next_state was staged");
static_assert(HID * TLD <= SLOT_W && 32 * 170 + 16 <= SLOT_W, "ring slot");
// MISC sub-offsets
constexpr int MI_Q = 0, MI_Y = 16, MI_DQ = 32, MI_RW = 48, MI_DN = 64, MI_RS = 80 /* [6][16] */, MI_ANEW = 176 /* [16][12] */;
constexpr int MI_AXA = 368 /* [16][24]: phase A target actions, phase B actions with the actor's */, MI_DL = 752 /* [16][12] */;
constexpr int MI_W1ACT = 944 /* [9][128] */, MI_LOSS = 2096;
static_assert(MI_LOSS + 16 <= MISC_FLOATS, "misc plan");
// PERS sub-offsets (vectors that must outlive their ring slot)
constexpr int PE_G1 = 0, PE_BE1 = 128, PE_G2 = 256, PE_BE2 = 384, PE_W3 = 512;         // critic (phase A: c, phase B: c2)
constexpr int PE_AG1 = 640, PE_ABE1 = 768, PE_AG2 = 896, PE_ABE2 = 1024, PE_AW3 = 1152; // actor (phase B); W3 as [16][ALD] (rows >= 9 zero), then b3
constexpr int PE_AB3 = PE_AW3 + 16 * ALD;
constexpr int W3P = 16 * ALD + 16;          // a padded head: W3 [16][ALD] K-major, rows >= NA zero, then the bias
constexpr int PE_T0 = 640, PE_T1 = 896;                                               // phase A: gamma | beta of target actor 0 / 1 (layer 1, then layer 2)
static_assert(PE_AB3 + 16 <= PERS_FLOATS, "pers plan");

struct ClusterArgs {
  LearnArgs a;
  float* gpart[2 * MAXN];                   // [B / 16][round4(net params)] partial gradients of every row block
  float* lpart;                             // [2n][B / 16] loss partial sums (critics, then actors)
  long long gstride[2 * MAXN];
  int cp_begin, cp_end;                     // cluster-kernel phases: 0 A, 1 Adam C, 2 B, 3 Adam A
  // data-parallel gradient exchange over NVLink peer memory (world > 1): every rank's exchange block, mapped into this process
  int world, rank;
  uint4* peer_ll[GW_MAX_PEERS];             // [world] rank p's exchange block: one slot of ll_stride lines per sending rank q
  long long ll_stride;                      // lines per slot: two 16-byte lines {x, epoch, y, epoch} {z, epoch, w, epoch} per float4 of gradient
  int split_roles;                          // actor-side clusters next to the row-block owners (when enough clusters are resident)
  int reduce_scatter;                       // 0: all-gather (every rank sums everything); 1: the owner of a float4 (index mod world) sums and returns it
  unsigned long long epoch0;                // exchanges completed before this launch
  unsigned long long timeout_ns;
  unsigned int* peer_err;                   // set to 1 if a peer did not arrive in time (the launch then finishes without it)
};

// ------------------------------------------------------------------------------------------------ primitives
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Flag-in-data lines (the LL protocol of NCCL): a 16-byte store carries two {value, epoch} pairs, each 8-byte half validates
// itself, so the receiver needs neither a fence nor a separate arrival word -- it polls the line until both epochs match.
__device__ __forceinline__ void st_ll(uint4* p, float x, float y, uint32_t ep) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(x)), "r"(ep), "r"(__float_as_uint(y)), "r"(ep) : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldcv4(const float* p) { return __ldcv(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_peer(const void* p, uint32_t peer) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(peer));
  return r;
}
__device__ __forceinline__ void st_peer4(uint32_t addr, const float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* g) {          // read-only data of this launch only (.ca)
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc (16 x 16, two n8 blocks in mma C-fragment layout) += A B over k-steps k0 = kbeg, kbeg + kstride, ... < kend (8 wide).
// AL 0: A[m][k] at A[m * lda + k]; AL 1: A[m][k] at A[k * lda + m].  BL 0: B[n][k] at B[n * ldb + k]; BL 1: at B[k * ldb + n].
// With a K-major A the two k of a thread are adjacent (one 64-bit load); the same pairing is applied to B.
template <int AL, int BL>
__device__ __forceinline__ void mma_tile(float (&acc)[2][4], const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                         int kbeg, int kend, int kstride) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int k0 = kbeg; k0 < kend; k0 += kstride) {
    float av[4], bv[2][2];
    if (AL == 0) {
      const float2 r0 = *reinterpret_cast<const float2*>(A + g * lda + k0 + 2 * t);
      const float2 r1 = *reinterpret_cast<const float2*>(A + (g + 8) * lda + k0 + 2 * t);
      av[0] = r0.x; av[2] = r0.y; av[1] = r1.x; av[3] = r1.y;
    } else {
      av[0] = A[(k0 + t) * lda + g]; av[1] = A[(k0 + t) * lda + g + 8];
      av[2] = A[(k0 + t + 4) * lda + g]; av[3] = A[(k0 + t + 4) * lda + g + 8];
    }
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      if (BL == 0) {
        const float2 w = *reinterpret_cast<const float2*>(B + (nb * 8 + g) * ldb + k0 + 2 * t);
        bv[nb][0] = w.x; bv[nb][1] = w.y;
      } else if (AL == 0) {
        bv[nb][0] = B[(k0 + 2 * t) * ldb + nb * 8 + g]; bv[nb][1] = B[(k0 + 2 * t + 1) * ldb + nb * 8 + g];
      } else {
        bv[nb][0] = B[(k0 + t) * ldb + nb * 8 + g]; bv[nb][1] = B[(k0 + t + 4) * ldb + nb * 8 + g];
      }
    }
    uint32_t ah[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(av[i], ah[i], al[i]);
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      uint32_t bh0, bl0, bh1, bl1;
      split_tf32(bv[nb][0], bh0, bl0);
      split_tf32(bv[nb][1], bh1, bl1);
      mma8(acc[nb], al, bh0, bh1);                  // small terms first
      mma8(acc[nb], ah, bl0, bl1);
      mma8(acc[nb], ah, bh0, bh1);
    }
  }
}
// Gumbel noise of action q for (update, row, agent, which): the same Philox words gumbel_row hands out, one action per lane
__device__ __forceinline__ float gumbel_lane(const LearnArgs& a, unsigned long long upd, int row, int agent, int which, int q) {
  uint32_t w[4] = {(uint32_t)row, (uint32_t)agent | ((uint32_t)which << 8) | ((uint32_t)(q >> 2) << 16), (uint32_t)upd,
                   (uint32_t)(upd >> 32) ^ 0x6C6561u};
  gw::philox4x32(w, a.gk0, a.gk1);
  const int i = q & 3;
  return gumbel_of(i == 0 ? w[0] : i == 1 ? w[1] : i == 2 ? w[2] : w[3]);
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// softmax over the NA values held by lanes 0 .. NA-1 (one per lane); lane q returns probability q, the other lanes 0
__device__ __forceinline__ float softmax_lanes(float v, int lane) {
  v = lane < NA ? v : -3.4e38f;
  const float mx = warp_max(v);
  const float e = lane < NA ? expf(v - mx) : 0.f;
  const float den = warp_sum(e);
  return e * (1.0f / den);
}
// softmax over the NA values held by lanes c = lane & 15 < NA of each half warp; lane c returns probability c, the others 0
__device__ __forceinline__ float softmax_half(float v, int c) {
  v = c < NA ? v : -3.4e38f;
  float mx = v;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float e = c < NA ? expf(v - mx) : 0.f;
  float den = e;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
  return e * (1.0f / den);
}
// logits of one row: lane q < NA gets h . W3[q] + b3[q]   (h: this lane's 4 columns; W3 [NA][HID] then b3 in shared memory)
__device__ __forceinline__ float head_logit_lane(const float4 h, const float* w3, int lane) {
  float mine = 0.f;
#pragma unroll
  for (int q = 0; q < NA; ++q) {
    const float s = warp_sum(dot4(h, *reinterpret_cast<const float4*>(w3 + q * HID + 4 * lane)));
    mine = lane == q ? s : mine;
  }
  return lane < NA ? mine + w3[NA * HID + lane] : 0.f;
}

__device__ __forceinline__ void zero_acc(float (&acc)[2][4]) {
#pragma unroll
  for (int nb = 0; nb < 2; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nb][i] = 0.f;
}

// ------------------------------------------------------------------------------------------------ package loads (cp.async)
// CW rows [o][k0 .. k0 + K) of a K-major weight matrix (row pitch `pitch` floats) -> slot[WLEAD + o * ldw + k]; the
// source rows are only 8-byte aligned when the pitch is 338, so whole aligned 16-byte chunks are copied and land at the
// same phase in shared memory (ldw == pitch mod 4): a row's chunks may spill up to 3 floats into its neighbours' padding.
__device__ __forceinline__ void load_w(float* slot, int ldw, const float* base, long long g0, int pitch, int K) {
  const int nch = (K >> 2) + 2;
  for (int i = threadIdx.x; i < CW * nch; i += THREADS) {
    const int o = i / nch, j = i - o * nch;
    const long long gi0 = g0 + (long long)o * pitch;
    const int m = (int)(gi0 & 3);
    const long long c = gi0 - m + 4 * j;
    if (c < gi0 + K) cp_async16(slot + WLEAD + o * ldw - m + 4 * j, base + c);
  }
}
// `rows` rows of CW floats: W[o][col0 .. col0 + CW) -> dst[o * TLD + j]   (the column slice the input-gradient GEMM needs)
__device__ __forceinline__ void load_wt(float* dst, const float* w, int col0, int rows) {
  for (int i = threadIdx.x; i < rows * (CW / 4); i += THREADS) {
    const int o = i / (CW / 4), q = i - o * (CW / 4);
    cp_async16(dst + o * TLD + 4 * q, w + (long long)o * HID + col0 + 4 * q);
  }
}
__device__ __forceinline__ void load_vec(float* dst, const float* src, int n4) {       // n4 16-byte chunks
  for (int i = threadIdx.x; i < n4; i += THREADS) cp_async16(dst + 4 * i, src + 4 * i);
}

// slot vector area: [0,32) bias slice | [32,160) gamma | [160,288) beta | [288,416) w3 | [416,420) b3
constexpr int V_B = 0, V_G = 32, V_BE = 160, V_W3 = 288, V_B3 = 416;
__device__ __forceinline__ void load_layer_vecs(float* v, const float* net, int b_off, int col0, bool head_w3, int w3_off, int b3_off) {
  load_vec(v + V_B, net + b_off + col0, CW / 4);
  load_vec(v + V_G, net + b_off + HID, 32);
  load_vec(v + V_BE, net + b_off + 2 * HID, 32);
  if (head_w3) { load_vec(v + V_W3, net + w3_off, 32); load_vec(v + V_B3, net + b3_off, 1); }
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(THREADS, 1) gw_learn_cluster_kernel(const ClusterArgs ca) {
  extern __shared__ __align__(16) float sm[];
  __shared__ long long s_row[RR], s_row1[RR];
  __shared__ int s_ended[RR];
  const LearnArgs& a = ca.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_rank(), col0 = rank * CW;
  const int B = a.B, NRB = B / RR;
  // clusters beyond the 2 * B / 16 that own row blocks are helpers: they only take part in the Adam phases and the barriers
  const int cid = blockIdx.x / CL;
  // Roles.  NW = 2 * B / 16 clusters own a row block of one agent.  When twice as many are resident and the launch runs whole
  // updates, a second set of NW clusters owns the ACTOR side of the same row blocks: it runs the actor's forward pass (which
  // reads nothing the critic's update writes) while the first set is in phase A, and the rest of phase B afterwards, so the
  // actor's forward leaves the critical path.  Everything else only joins the Adam phases and the barriers.
  const int NW = N2 * NRB;
  const bool split = ca.cp_begin == 0 && ca.cp_end == 4 && (int)(gridDim.x / CL) >= 2 * NW && ca.split_roles;
  const int role = cid < NW ? 0 : (split && cid < 2 * NW ? 1 : 2);      // 0: row-block owner (critic side when split), 1: actor side, 2: helper
  const bool worker = role != 2;
  const int wid = role == 1 ? cid - NW : cid;
  const int ag = worker ? wid / NRB : 0, rb = worker ? wid - ag * NRB : 0, row0 = rb * RR;
  const unsigned n_ctas = gridDim.x;
  const NetLayout la = a.la, lc = a.lc;
  float* X = sm + SM_X; float* X2 = sm + SM_X2; float* EX = sm + SM_EX; float* PV = sm + SM_PV; float* RING = sm + SM_RING;
  float* RED = sm + SM_RED; float* OUT = sm + SM_OUT; float* PERS = sm + SM_PERS; float* MISC = sm + SM_MISC;
  auto ex = [&](int set, int slot) { return EX + (set * 3 + slot) * RR * ALD; };
  // private activation buffers: 0..2 own storage; 3, 4 where next_state was staged (dead after the first layers); 5, 6 the
  // exchange buffers only phase A's three-network rounds use
  auto pv = [&](int k) { return k < 3 ? PV + k * RR * ALD : k < 5 ? X2 + (k - 3) * RR * ALD : EX + (k - 4) * RR * ALD; };
  auto slot_w = [&](int p) { return RING + (p % NSLOT) * SLOT; };
  auto slot_v = [&](int p) { return RING + (p % NSLOT) * SLOT + SLOT_W; };
  float step0[2 * MAXN];
  for (int k = 0; k < 2 * N2; ++k) step0[k] = __ldcg(a.steps + k);
  if (worker) for (int i = tid; i < SM_TOTAL; i += THREADS) sm[i] = 0.f;      // every later content is a finite float (padding is multiplied by 0)
  __syncthreads();
  cluster_sync();                                                 // peers' shared memory exists and is zeroed before anyone writes into it
  int round = 0;                                                  // all-gather rounds done (parity selects the exchange set)
  bool x_valid = false;

  // ---- the cluster's 16 batch rows: draw (ring mode) and stage [state | action] into X, next_state into X2
  auto stage_rows = [&](bool want_x2, int u) {
    if (a.sample) {
      const gw_replay_view& r = a.ring;
      if (tid < RR) {
        long long t_abs, e;
        gwr::draw_index(row0 + tid, a.draw_base + (unsigned long long)u, a.rk0, a.rk1, a.t_now, a.n_valid,
                        r.num_envs, t_abs, e);
        long long sl = t_abs % r.slots;
        if (sl < 0) sl += r.slots;
        const long long sl1 = sl + 1 == r.slots ? 0 : sl + 1;
        s_row[tid] = sl * r.num_envs + e;
        s_row1[tid] = sl1 * r.num_envs + e;
        s_ended[tid] = r.ended[sl * r.num_envs + e] != 0;
      }
      __syncthreads();
      if (tid < RR) {
        MISC[MI_RW + tid] = r.reward[s_row[tid] * N2 + ag];
        MISC[MI_DN + tid] = r.terminated[s_row[tid] * N2 + ag] ? 1.0f : 0.0f;
      }
      if (r.obs_dtype == GW_OBS_F32) {
        const float* obs = static_cast<const float*>(r.obs);
        const float* fin = static_cast<const float*>(r.final_obs);
        for (int i = tid; i < RR * (SO / 4); i += THREADS) {
          const int j = i / (SO / 4), q = i - j * (SO / 4);
          cp_async16(X + j * XLD + 4 * q, obs + s_row[j] * SO + 4 * q);
          if (want_x2) cp_async16(X2 + j * XLD + 4 * q, (s_ended[j] ? fin + s_row[j] * SO : obs + s_row1[j] * SO) + 4 * q);
        }
      } else {
        const __nv_bfloat16* obs = static_cast<const __nv_bfloat16*>(r.obs);
        const __nv_bfloat16* fin = static_cast<const __nv_bfloat16*>(r.final_obs);
        for (int j = warp; j < RR; j += WARPS) {
          gwr::copy_row<__nv_bfloat16>(obs + s_row[j] * SO, X + j * XLD, SO, lane, true);
          if (want_x2) gwr::copy_row<__nv_bfloat16>(s_ended[j] ? fin + s_row[j] * SO : obs + s_row1[j] * SO, X2 + j * XLD, SO, lane, true);
        }
      }
      for (int i = tid; i < RR * SA; i += THREADS) {
        const int j = i / SA, q = i - j * SA;
        cp_async4(X + j * XLD + SO + q, r.action + s_row[j] * SA + q);
      }
    } else {
      if (tid < RR) {
        MISC[MI_RW + tid] = __ldg(a.bR + (long long)(row0 + tid) * N2 + ag);
        MISC[MI_DN + tid] = __ldg(a.bD + (long long)(row0 + tid) * N2 + ag);
      }
      for (int i = tid; i < RR * (SO / 4); i += THREADS) {
        const int j = i / (SO / 4), q = i - j * (SO / 4);
        cp_async16(X + j * XLD + 4 * q, a.bS + (long long)(row0 + j) * SO + 4 * q);
        if (want_x2) cp_async16(X2 + j * XLD + 4 * q, a.bS2 + (long long)(row0 + j) * SO + 4 * q);
      }
      for (int i = tid; i < RR * SA; i += THREADS) {
        const int j = i / SA, q = i - j * SA;
        cp_async4(X + j * XLD + SO + q, a.bACT + (long long)(row0 + j) * SA + q);
      }
    }
  };

  // ---- 8 per-warp partial tiles -> this CTA's 16 x 16 slice (+ bias) -> the same columns of every peer's exchange buffer
  auto reduce_exchange = [&](float (&acc)[2][4], const float* bias, float* exbuf) {
    const int g = lane >> 2, t = lane & 3;
    float* my = RED + warp * RR * RLD;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      *reinterpret_cast<float2*>(my + g * RLD + nb * 8 + 2 * t) = make_float2(acc[nb][0], acc[nb][1]);
      *reinterpret_cast<float2*>(my + (g + 8) * RLD + nb * 8 + 2 * t) = make_float2(acc[nb][2], acc[nb][3]);
    }
    __syncthreads();
    {
      const int r = tid >> 4, c = tid & 15;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {                           // warp w holds columns 16 (w & 1) .. + 16 over k-part w >> 1
        float v = bias ? bias[16 * hf + c] : 0.f;
#pragma unroll
        for (int kp = 0; kp < WARPS / 2; ++kp) v += RED[(2 * kp + hf) * RR * RLD + r * RLD + c];
        OUT[r * OLD + 16 * hf + c] = v;
      }
    }
    __syncthreads();
    {
      const int peer = tid >> 6, rem = tid & 63, r = rem >> 2, q = rem & 3;
      const float4 v0 = *reinterpret_cast<const float4*>(OUT + r * OLD + 4 * q);
      const float4 v1 = *reinterpret_cast<const float4*>(OUT + r * OLD + 16 + 4 * q);
      const uint32_t dst = map_peer(exbuf + r * ALD + col0 + 4 * q, (uint32_t)peer);
      st_peer4(dst, v0);
      st_peer4(dst + 64, v1);
    }
  };

  // ---- LayerNorm + ReLU of the 16 gathered rows (warp w: rows 2w, 2w+1): xhat / h / rstd to private buffers (nullable)
  auto ln_rows = [&](const float* z, const float* gam, const float* bet, float* xh_out, float* h_out, float* rs_out, float4 (&hreg)[2]) {
    const float4 g = *reinterpret_cast<const float4*>(gam + 4 * lane), be = *reinterpret_cast<const float4*>(bet + 4 * lane);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = 2 * warp + rr;
      float mu, rs;
      const float4 xh = ln_row(*reinterpret_cast<const float4*>(z + r * ALD + 4 * lane), a.ln_eps, mu, rs);
      const float4 h = affine_relu(xh, g, be);
      if (xh_out) st4(xh_out + r * ALD + 4 * lane, xh);
      if (h_out) st4(h_out + r * ALD + 4 * lane, h);
      if (rs_out && lane == 0) rs_out[r] = rs;
      hreg[rr] = h;
    }
  };
  // LayerNorm + ReLU backward of the rows of this warp: dh (in registers) -> dz into `dz_out`; xhat from `xh`, mask recomputed
  auto ln_bwd_rows = [&](const float4 (&dh)[2], const float* xh, const float* gam, const float* bet, const float* rs, float* dz_out) {
    const float4 g = *reinterpret_cast<const float4*>(gam + 4 * lane), be = *reinterpret_cast<const float4*>(bet + 4 * lane);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = 2 * warp + rr;
      const float4 x = *reinterpret_cast<const float4*>(xh + r * ALD + 4 * lane);
      const float4 h = affine_relu(x, g, be);
      float4 dy;
      const float4 dz = ln_relu_bwd_row(dh[rr], h, x, g, rs[r], dy);
      st4(dz_out + r * ALD + 4 * lane, dz);
    }
  };
  // column-slice sums over the 16 rows of LayerNorm-layer vector gradients: db = sum dz, dgamma = sum dy * xhat, dbeta = sum dy
  // (dy = dh where the ReLU passed; dh given per row by `dh_of(row, col)`), written to the row block's gradient slab
  auto ln_vec_grads = [&](float* slab, int b_off, const float* dz, const float* xh, const float* gam, const float* bet, auto dh_of) {
    if (tid < 3 * CW) {
      const int v = tid / CW, c = col0 + (tid - v * CW);   // 96 threads
      const float g = gam[c], be = bet[c];
      float s = 0.f;
      for (int r = 0; r < RR; ++r) {
        if (v == 0) { s += dz[r * ALD + c]; continue; }
        const float x = xh[r * ALD + c];
        const float dy = fmaf(x, g, be) > 0.f ? dh_of(r, c) : 0.f;
        s += v == 1 ? dy * x : dy;
      }
      slab[b_off + v * HID + c] = s;
    }
  };
  // weight-gradient slice dW[col0 + m][n] = sum_rows dz[row][col0 + m] * in[row][n], n < N: warp w takes n-blocks w, w + 8, ...
  auto dw_slice = [&](float* slab_w, int pitch, const float* dz, const float* in, int ldin, int N) {
    const int g = lane >> 2, t = lane & 3;
    for (int nb0 = warp * 16; nb0 < N; nb0 += WARPS * 16) {
#pragma unroll
      for (int mt = 0; mt < CW / 16; ++mt) {
        float acc[2][4];
        zero_acc(acc);
        mma_tile<1, 1>(acc, dz + col0 + 16 * mt, ALD, in + nb0, ldin, 0, RR, 8);
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          const int n = nb0 + nb * 8 + 2 * t;
          if (n < N) {
            *reinterpret_cast<float2*>(slab_w + (long long)(col0 + 16 * mt + g) * pitch + n) = make_float2(acc[nb][0], acc[nb][1]);
            *reinterpret_cast<float2*>(slab_w + (long long)(col0 + 16 * mt + g + 8) * pitch + n) = make_float2(acc[nb][2], acc[nb][3]);
          }
        }
      }
    }
  };

  // this warp's share of a layer GEMM: output columns 16 (warp & 1) .. + 16 of the CTA's 32, k-steps (warp >> 1) mod 4
  auto gemm = [&](float (&acc)[2][4], const float* A, int lda, const float* wslot, int ldw, int K8) {
    mma_tile<0, 0>(acc, A, lda, wslot + WLEAD + (warp & 1) * 16 * ldw, ldw, 8 * (warp >> 1), K8, 32);
  };
  auto gemm_dh = [&](float (&acc)[2][4], const float* A, const float* wt) {      // dz W, W as [k][CW] slice
    mma_tile<0, 1>(acc, A, ALD, wt + (warp & 1) * 16, TLD, 8 * (warp >> 1), HID, 32);
  };

  // output layer of an actor on the tensor cores: logits[16 rows][16] = h [16][HID] . w3p^T over the 8 warps' k-slices, joined
  // through RED; thread (r = tid >> 4, c = tid & 15) returns logit c of row r (bias added for c < NA)
  auto head_logits = [&](const float* h, const float* w3p) -> float {
    float acc[2][4];
    zero_acc(acc);
    mma_tile<0, 0>(acc, h, ALD, w3p, ALD, 8 * warp, HID, 64);
    const int g = lane >> 2, t = lane & 3;
    float* my = RED + warp * RR * RLD;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      *reinterpret_cast<float2*>(my + g * RLD + nb * 8 + 2 * t) = make_float2(acc[nb][0], acc[nb][1]);
      *reinterpret_cast<float2*>(my + (g + 8) * RLD + nb * 8 + 2 * t) = make_float2(acc[nb][2], acc[nb][3]);
    }
    __syncthreads();
    const int r = tid >> 4, c = tid & 15;
    float v = c < NA ? w3p[16 * ALD + c] : 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) v += RED[w * RR * RLD + r * RLD + c];
    __syncthreads();
    return v;
  };
#ifdef GW_LEARN_TRACE
  int tpk = 0;
#define TP() do { if (blockIdx.x == 0 && tid == 0 && u == a.updates - 1 && tpk < 100) a.s.trace[8 + tpk] = phase_clock(); ++tpk; } while (0)
#else
#define TP() do { } while (0)
#endif
  for (int u = 0; u < a.updates; ++u) {
    const unsigned long long upd = a.upd_base + (unsigned long long)u;
#ifdef GW_LEARN_TRACE
    tpk = 0;
#endif
    for (int ph = ca.cp_begin; ph < ca.cp_end; ++ph) {
      if (blockIdx.x == 0 && tid == 0 && u == a.updates - 1) a.s.trace[ph] = phase_clock();
      const bool run_a = ph == 0 && role == 0;
      const bool run_b1 = split ? (ph == 0 && role == 1) : (ph == 2 && role == 0);     // actor forward
      const bool run_b2 = ph == 2 && (split ? role == 1 : role == 0);                  // the rest of phase B
      if (run_a) {
        // ======================================================================================== phase A: critic gradients
        const float* Tc = a.T + a.net_off[N2 + ag];           // target critic
        const float* Pc = a.P + a.net_off[N2 + ag];           // critic
        float* slab = ca.gpart[N2 + ag] + (long long)rb * ca.gstride[N2 + ag];
        auto issue = [&](int p) {                              // package p of phase A -> ring slot p % 3 (always commits a group)
          float* w = slot_w(p); float* v = slot_v(p);
          const long long c1 = a.net_off[N2 + ag] + lc.w1 + (long long)col0 * CI;      // this CTA's rows of the critics' W1
          switch (p) {
            case 0: case 1: { const float* net = a.T + a.net_off[p];
              load_w(w, 168, a.T, a.net_off[p] + la.w1 + (long long)col0 * O, O, O); load_layer_vecs(v, net, la.b1, col0, false, 0, 0); } break;
            case 2: load_w(w, 170, a.P, c1, CI, O); break;
            case 3: load_w(w, 170, a.P, c1 + O, CI, SO - O); break;
            case 4: load_w(w, 26, a.P, c1 + SO, CI, SA); load_layer_vecs(v, Pc, lc.b1, col0, false, 0, 0); break;
            case 5: load_w(w, 170, a.T, c1, CI, O); break;
            case 6: load_w(w, 170, a.T, c1 + O, CI, SO - O); break;
            case 7: case 8: { const float* net = a.T + a.net_off[p - 7];
              load_w(w, ALD, a.T, a.net_off[p - 7] + la.w2 + (long long)col0 * HID, HID, HID); load_layer_vecs(v, net, la.b2, col0, false, 0, 0); } break;
            case 9: load_w(w, ALD, a.P, a.net_off[N2 + ag] + lc.w2 + (long long)col0 * HID, HID, HID); load_layer_vecs(v, Pc, lc.b2, col0, true, lc.w3, lc.b3); break;
            case 10: for (int k = 0; k < N2; ++k) { const float* net = a.T + a.net_off[k];   // the target actors' heads, padded layout
              for (int i = tid; i < NA * (HID / 4); i += THREADS) cp_async16(w + k * W3P + (i >> 5) * ALD + 4 * (i & 31), net + la.w3 + 4 * i);
              load_vec(w + k * W3P + 16 * ALD, net + la.b3, 3); } break;
            case 11: load_w(w, 26, a.T, c1 + SO, CI, SA); load_layer_vecs(v, Tc, lc.b1, col0, false, 0, 0); break;
            case 12: load_w(w, ALD, a.T, a.net_off[N2 + ag] + lc.w2 + (long long)col0 * HID, HID, HID); load_layer_vecs(v, Tc, lc.b2, col0, true, lc.w3, lc.b3); break;
            case 13: load_wt(w, Pc + lc.w2, col0, HID); break;
            default: break;
          }
          cp_commit();
        };
        auto ready = [&](int p, bool first = false) {         // package p landed and visible; then prefetch package p + 2
          if (first) cp_wait<0>(); else cp_wait<1>();
          __syncthreads();
          issue(p + 2);
        };
        stage_rows(true, u);
        TP();
        issue(0); issue(1);
        float acc_t[N2][2][4], acc_c[2][4], acc_ct[2][4];
        zero_acc(acc_t[0]); zero_acc(acc_t[1]); zero_acc(acc_c); zero_acc(acc_ct);
        // ---- round 1: first layers
        ready(0, true);  gemm(acc_t[0], X2, XLD, slot_w(0), 168, O);
        TP();
        float* TAh[N2] = {pv(0), pv(1)};
        float* XH2 = pv(2); float* XH1 = pv(3); float* H1 = pv(4);   // 3, 4: where next_state is staged, dead after this round's GEMMs
        float* A2 = MISC + MI_AXA;                                  // target actions [16][24], zero padded
        float* rs1 = MISC + MI_RS; float* rs2 = MISC + MI_RS + 16;
        {
          float* e0 = ex(round & 1, 0);
          reduce_exchange(acc_t[0], slot_v(0) + V_B, e0);
          for (int i = tid; i < 2 * HID; i += THREADS) PERS[PE_T0 + i] = slot_v(0)[V_G + i];
        }
        ready(1);  gemm(acc_t[1], X2 + O, XLD, slot_w(1), 168, O);
        TP();
        reduce_exchange(acc_t[1], slot_v(1) + V_B, ex(round & 1, 1));
        for (int i = tid; i < 2 * HID; i += THREADS) PERS[PE_T1 + i] = slot_v(1)[V_G + i];
        ready(2);  gemm(acc_c, X, XLD, slot_w(2), 170, O);
        TP();
        ready(3);  gemm(acc_c, X + O, XLD, slot_w(3), 170, O);
        TP();
        ready(4);  gemm(acc_c, X + SO, XLD, slot_w(4), 26, 24);
        TP();
        reduce_exchange(acc_c, slot_v(4) + V_B, ex(round & 1, 2));
        for (int i = tid; i < 2 * HID; i += THREADS) PERS[PE_G1 + i] = slot_v(4)[V_G + i];     // gamma1 | beta1 of the critic
        ready(5);  gemm(acc_ct, X2, XLD, slot_w(5), 170, O);
        TP();
        ready(6);  gemm(acc_ct, X2 + O, XLD, slot_w(6), 170, O);
        TP();
        cluster_sync();
        TP();
        float4 hreg[2];
        {
          const int set = round & 1;
          ln_rows(ex(set, 0), PERS + PE_T0, PERS + PE_T0 + HID, nullptr, TAh[0], nullptr, hreg);
          ln_rows(ex(set, 1), PERS + PE_T1, PERS + PE_T1 + HID, nullptr, TAh[1], nullptr, hreg);
          ln_rows(ex(set, 2), PERS + PE_G1, PERS + PE_BE1, XH1, H1, rs1, hreg);
          ++round;
        }
        // ---- round 2: second layers
        float acc2[2][4];
        for (int k = 0; k < N2; ++k) {
          ready(7 + k);
          TP();
          zero_acc(acc2);
          gemm(acc2, TAh[k], ALD, slot_w(7 + k), ALD, HID);
          reduce_exchange(acc2, slot_v(7 + k) + V_B, ex(round & 1, k));
          for (int i = tid; i < 2 * HID; i += THREADS) PERS[(k ? PE_T1 : PE_T0) + i] = slot_v(7 + k)[V_G + i];   // layer-1 values were consumed before round 2
        }
        ready(9);
        TP();
        zero_acc(acc2);
        gemm(acc2, H1, ALD, slot_w(9), ALD, HID);
        reduce_exchange(acc2, slot_v(9) + V_B, ex(round & 1, 2));
        for (int i = tid; i < 3 * HID; i += THREADS) PERS[PE_G2 + i] = slot_v(9)[V_G + i];     // gamma2 | beta2 | w3 of the critic
        const float b3c = slot_v(9)[V_B3];
        ready(10);
        TP();
        for (int i = tid; i < N2 * 7 * ALD; i += THREADS) slot_w(10)[(i / (7 * ALD)) * W3P + NA * ALD + i % (7 * ALD)] = 0.f;   // rows 9..15 of the padded heads
        cluster_sync();
        TP();
        {
          const int set = round & 1;
          // target actors' heads: Gumbel-softmax actions on next_state -> A2[:, k * 9 ...]
          for (int k = 0; k < N2; ++k) ln_rows(ex(set, k), PERS + (k ? PE_T1 : PE_T0), PERS + (k ? PE_T1 : PE_T0) + HID, nullptr, TAh[k], nullptr, hreg);
          __syncthreads();
          for (int k = 0; k < N2; ++k) {
            float lg = head_logits(TAh[k], slot_w(10) + k * W3P);
            const int r = tid >> 4, c = tid & 15, row = row0 + r;
            if (c < NA) lg += a.gum_next ? __ldg(a.gum_next + ((long long)row * N2 + k) * NA + c) : gumbel_lane(a, upd, row, k, 0, c);
            const float pl = softmax_half(lg, c);
            if (c < NA) {
              A2[r * AXLD + k * NA + c] = pl;
              if (rank == 0 && ag == 0) a.s.a2[(long long)row * SA + k * NA + c] = pl;
            } else if (k == 0 && c - NA < AXLD - SA) {
              A2[r * AXLD + SA + c - NA] = 0.f;
            }
          }
          // critic head: Q(state, action)
          ln_rows(ex(set, 2), PERS + PE_G2, PERS + PE_BE2, XH2, nullptr, rs2, hreg);
          const float4 w3c = *reinterpret_cast<const float4*>(PERS + PE_W3 + 4 * lane);
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const float qv = warp_sum(dot4(hreg[rr], w3c)) + b3c;
            if (lane == 0) { MISC[MI_Q + 2 * warp + rr] = qv; if (rank == 0) a.s.q[ag][row0 + 2 * warp + rr] = qv; }
          }
          ++round;
        }
        // ---- round 3: target critic, first layer completed with the target actions
        ready(11);
        TP();
        gemm(acc_ct, A2, AXLD, slot_w(11), 26, 24);
        reduce_exchange(acc_ct, slot_v(11) + V_B, ex(round & 1, 0));
        cluster_sync();
        TP();
        float* CTh = TAh[0];
        ln_rows(ex(round & 1, 0), slot_v(11) + V_G, slot_v(11) + V_BE, nullptr, CTh, nullptr, hreg);
        ++round;
        // ---- round 4: target critic, second layer -> TD target -> critic loss -> backward through layer 3 / LayerNorm 2
        ready(12);
        TP();
        zero_acc(acc2);
        gemm(acc2, CTh, ALD, slot_w(12), ALD, HID);
        reduce_exchange(acc2, slot_v(12) + V_B, ex(round & 1, 0));
        cluster_sync();
        TP();
        float* DZ2 = TAh[1]; float* DZ1 = TAh[0];
        {
          ln_rows(ex(round & 1, 0), slot_v(12) + V_G, slot_v(12) + V_BE, nullptr, nullptr, nullptr, hreg);
          const float4 w3t = *reinterpret_cast<const float4*>(slot_v(12) + V_W3 + 4 * lane);
          const float b3t = slot_v(12)[V_B3];
          const float4 w3c = *reinterpret_cast<const float4*>(PERS + PE_W3 + 4 * lane);
          float4 dh[2];
          float lsum = 0.f;
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * warp + rr;
            const float qn = warp_sum(dot4(hreg[rr], w3t)) + b3t;
            const float yv = MISC[MI_RW + r] + a.gamma * (1.0f - MISC[MI_DN + r]) * qn;
            const float diff = MISC[MI_Q + r] - yv;
            const float dqv = 2.0f * diff / (float)B;
            lsum += diff * diff / (float)B;
            dh[rr] = scale4(w3c, dqv);
            if (lane == 0) {
              MISC[MI_DQ + r] = dqv;
              if (rank == 0) { a.s.y[ag][row0 + r] = yv; a.s.dq[ag][row0 + r] = dqv; }
            }
          }
          if (lane == 0) MISC[MI_LOSS + warp] = lsum;
          ln_bwd_rows(dh, XH2, PERS + PE_G2, PERS + PE_BE2, rs2, DZ2);
          ++round;
        }
        __syncthreads();
        if (rank == 0 && tid == 0) {
          float v = 0.f;
          for (int w = 0; w < WARPS; ++w) v += MISC[MI_LOSS + w];
          ca.lpart[ag * NRB + rb] = v;
        }
        // vector gradients of layer 3 / LayerNorm 2 (column slice), db3
        ln_vec_grads(slab, lc.b2, DZ2, XH2, PERS + PE_G2, PERS + PE_BE2, [&](int r, int c) { return MISC[MI_DQ + r] * PERS[PE_W3 + c]; });
        if (tid >= 96 && tid < 96 + CW) {                       // dW3[c] = sum_r dq[r] * h2[r][c]
          const int c = col0 + tid - 96;
          float s = 0.f;
          for (int r = 0; r < RR; ++r) s += MISC[MI_DQ + r] * fmaxf(fmaf(XH2[r * ALD + c], PERS[PE_G2 + c], PERS[PE_BE2 + c]), 0.f);
          slab[lc.w3 + c] = s;
        }
        if (rank == 0 && tid == 128) {
          float s = 0.f;
          for (int r = 0; r < RR; ++r) s += MISC[MI_DQ + r];
          slab[lc.b3] = s;
        }
        // ---- round 5: dh1 = dz2 W2 (this CTA's 16 columns) -> LayerNorm 1 backward
        ready(13);
        TP();
        zero_acc(acc2);
        gemm_dh(acc2, DZ2, slot_w(13));
        reduce_exchange(acc2, nullptr, ex(round & 1, 0));
        dw_slice(slab + lc.w2, HID, DZ2, H1, ALD, HID);         // dW2 rows of this CTA (independent of the exchange)
        TP();
        cluster_sync();
        TP();
        {
          float4 dh[2];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) dh[rr] = *reinterpret_cast<const float4*>(ex(round & 1, 0) + (2 * warp + rr) * ALD + 4 * lane);
          ln_bwd_rows(dh, XH1, PERS + PE_G1, PERS + PE_BE1, rs1, DZ1);
          __syncthreads();
          const float* dhx = ex(round & 1, 0);
          ln_vec_grads(slab, lc.b1, DZ1, XH1, PERS + PE_G1, PERS + PE_BE1, [&](int r, int c) { return dhx[r * ALD + c]; });
          ++round;
        }
        dw_slice(slab + lc.w1, CI, DZ1, X, XLD, CI);            // dW1 rows of this CTA: dz1^T [state | action]
        TP();
        cp_wait<0>();
        TP();
        x_valid = true;
      }
      if (run_b1 || run_b2) {
        // ======================================================================================== phase B: actor gradients
        const float* Pa = a.P + a.net_off[ag];
        const float* Pc = a.P + a.net_off[N2 + ag];
        float* slab = ca.gpart[ag] + (long long)rb * ca.gstride[ag];
        // weight packages in the order of their use: the actor's three first (they do not depend on the critic's Adam step and
        // may be loaded while it runs), then the updated critic's
        auto issue = [&](int pos) {
          float* w = slot_w(pos); float* v = slot_v(pos);
          switch (pos) {
            case 0: load_w(w, 168, a.P, a.net_off[ag] + la.w1 + (long long)col0 * O, O, O); load_layer_vecs(v, Pa, la.b1, col0, false, 0, 0); break;
            case 1: load_w(w, ALD, a.P, a.net_off[ag] + la.w2 + (long long)col0 * HID, HID, HID); load_layer_vecs(v, Pa, la.b2, col0, false, 0, 0); break;
            case 2: load_vec(w, Pa + la.w3, NA * HID / 4); load_vec(w + NA * HID, Pa + la.b3, 3); break;
            case 3: load_w(w, 170, a.P, a.net_off[N2 + ag] + lc.w1 + (long long)col0 * CI, CI, O); break;
            case 4: load_w(w, 170, a.P, a.net_off[N2 + ag] + lc.w1 + (long long)col0 * CI + O, CI, SO - O); break;
            case 5: load_w(w, 26, a.P, a.net_off[N2 + ag] + lc.w1 + (long long)col0 * CI + SO, CI, SA); load_layer_vecs(v, Pc, lc.b1, col0, false, 0, 0); break;
            case 6: load_w(w, ALD, a.P, a.net_off[N2 + ag] + lc.w2 + (long long)col0 * HID, HID, HID); load_layer_vecs(v, Pc, lc.b2, col0, true, lc.w3, lc.b3); break;
            case 7: load_wt(w, Pc + lc.w2, col0, HID); break;
            case 8: load_wt(w, Pa + la.w2, col0, HID); break;
            default: break;
          }
          cp_commit();
        };
        int issue_limit = 3;                                    // the actor's packages only until the critic has been stepped
        auto ready = [&](int pos, bool first = false) {
          if (first) cp_wait<0>(); else cp_wait<1>();
          __syncthreads();
          if (pos + 2 < issue_limit) issue(pos + 2); else cp_commit();
        };
        float* XH1a = pv(0); float* H1a = pv(1); float* XH2a = pv(2); float* XH1c = pv(3); float* XH2c = pv(4);
        float* D1 = pv(5); float* D2 = pv(6);                   // D1: h1 of the critic pass, then dz2 (critic), then dz2 (actor); 5, 6 = exchange
                                                                // buffers (set 0, slots 1 / 2) that this phase's one-network rounds never use
        float* rs1a = MISC + MI_RS + 32; float* rs2a = MISC + MI_RS + 48; float* rs1c = MISC + MI_RS + 64; float* rs2c = MISC + MI_RS + 80;
        float acc_a[2][4], acc_c[2][4], acc2[2][4];
        float4 hreg[2];
        if (run_b1) {
          // ============ the actor's forward pass (actor-side clusters run it while the critic side is in phase A)
          if (!x_valid) stage_rows(false, u);
          issue(0); issue(1);
          zero_acc(acc_a);
          // ---- round 1: actor layer 1
          ready(0, true);  gemm(acc_a, X + ag * O, XLD, slot_w(0), 168, O);
          TP();
          reduce_exchange(acc_a, slot_v(0) + V_B, ex(round & 1, 0));
          for (int i = tid; i < 2 * HID; i += THREADS) PERS[PE_AG1 + i] = slot_v(0)[V_G + i];
          cluster_sync();
          TP();
          ln_rows(ex(round & 1, 0), PERS + PE_AG1, PERS + PE_ABE1, XH1a, H1a, rs1a, hreg);
          ++round;
          // ---- round 2: actor layer 2 -> head -> Gumbel-softmax action
          ready(1);
          TP();
          zero_acc(acc2);
          gemm(acc2, H1a, ALD, slot_w(1), ALD, HID);
          reduce_exchange(acc2, slot_v(1) + V_B, ex(round & 1, 0));
          for (int i = tid; i < 2 * HID; i += THREADS) PERS[PE_AG2 + i] = slot_v(1)[V_G + i];
          ready(2);
          TP();
          for (int i = tid; i < NA * HID + 12; i += THREADS) {  // padded layout [16][ALD] + bias (rows 9..15 stay zero)
            if (i < NA * HID) PERS[PE_AW3 + (i >> 7) * ALD + (i & (HID - 1))] = slot_w(2)[i];
            else PERS[PE_AB3 + i - NA * HID] = slot_w(2)[i];
          }
          cluster_sync();
          TP();
          ln_rows(ex(round & 1, 0), PERS + PE_AG2, PERS + PE_ABE2, XH2a, D2, rs2a, hreg);   // h2 into D2 (free until round 5)
          ++round;
          __syncthreads();                                      // D2 and PERS (actor W3) written by all threads above
          {
            float lg = head_logits(D2, PERS + PE_AW3);
            const int r = tid >> 4, c = tid & 15, row = row0 + r;
            if (c < NA) lg += a.gum_cur ? __ldg(a.gum_cur + ((long long)row * N2 + ag) * NA + c) : gumbel_lane(a, upd, row, ag, 1, c);
            const float pl = softmax_half(lg, c);
            if (c < NA) {
              MISC[MI_ANEW + r * 12 + c] = pl;
              if (rank == 0) a.s.anew[ag][(long long)row * NA + c] = pl;
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                    // the batch's actions with this agent's replaced (zero padded): columns c, c + 16
              const int col = c + 16 * hf, k = col - ag * NA;
              const bool mine = k >= 0 && k < NA;
              const float pv_ = __shfl_sync(0xffffffffu, pl, (lane & 16) + (mine ? k : 0));
              if (col < AXLD) MISC[MI_AXA + r * AXLD + col] = col >= SA ? 0.f : (mine ? pv_ : X[r * XLD + SO + col]);
            }
          }
          cp_wait<0>();
          __syncthreads();
          x_valid = true;                                       // this cluster's rows are staged (its phase B continues with them)
        }
        if (run_b2) {
        // ============ the updated critic on [state | .. actor's action ..], -Q backward to the action, the actor's backward pass
        issue_limit = 9;
        issue(3); issue(4);
        // the critic's first-layer columns of this agent's action, transposed: W1ACT[k][o] (read after the critic's Adam step)
        for (int e = tid; e < NA * HID; e += THREADS) {
          const int o = e / NA, k = e - o * NA;
          MISC[MI_W1ACT + k * HID + o] = __ldcg(Pc + lc.w1 + (long long)o * CI + SO + ag * NA + k);
        }
        zero_acc(acc_c);
        // ---- round 3: the updated critic's layer 1 on [state | actions with the actor's]
        ready(3, true);  gemm(acc_c, X, XLD, slot_w(3), 170, O);
        TP();
        ready(4);  gemm(acc_c, X + O, XLD, slot_w(4), 170, O);
        TP();
        ready(5);
        TP();
        gemm(acc_c, MISC + MI_AXA, AXLD, slot_w(5), 26, 24);
        reduce_exchange(acc_c, slot_v(5) + V_B, ex(round & 1, 0));
        for (int i = tid; i < 2 * HID; i += THREADS) PERS[PE_G1 + i] = slot_v(5)[V_G + i];
        cluster_sync();
        TP();
        ln_rows(ex(round & 1, 0), PERS + PE_G1, PERS + PE_BE1, XH1c, D1, rs1c, hreg);
        ++round;
        // ---- round 4: critic layer 2 -> Q -> actor loss -> backward through layer 3 / LayerNorm 2
        ready(6);
        TP();
        zero_acc(acc2);
        gemm(acc2, D1, ALD, slot_w(6), ALD, HID);
        reduce_exchange(acc2, slot_v(6) + V_B, ex(round & 1, 0));
        for (int i = tid; i < 3 * HID; i += THREADS) PERS[PE_G2 + i] = slot_v(6)[V_G + i];
        const float b3c = slot_v(6)[V_B3];
        cluster_sync();
        TP();
        {
          ln_rows(ex(round & 1, 0), PERS + PE_G2, PERS + PE_BE2, XH2c, nullptr, rs2c, hreg);
          const float4 w3c = *reinterpret_cast<const float4*>(PERS + PE_W3 + 4 * lane);
          const float dqv = -1.0f / (float)B;
          float lsum = 0.f;
          float4 dh[2];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            lsum += -(warp_sum(dot4(hreg[rr], w3c)) + b3c) / (float)B;
            dh[rr] = scale4(w3c, dqv);
          }
          if (lane == 0) MISC[MI_LOSS + warp] = lsum;
          ln_bwd_rows(dh, XH2c, PERS + PE_G2, PERS + PE_BE2, rs2c, D1);     // D1 (h1 of the critic) was consumed by the GEMM above
          ++round;
        }
        __syncthreads();
        if (rank == 0 && tid == 0) {
          float v = 0.f;
          for (int w = 0; w < WARPS; ++w) v += MISC[MI_LOSS + w];
          ca.lpart[(N2 + ag) * NRB + rb] = v;
        }
        // ---- round 5: dh1 (critic) -> LayerNorm 1 backward -> gradient w.r.t. the action -> softmax -> actor layer 3
        ready(7);
        TP();
        zero_acc(acc2);
        gemm_dh(acc2, D1, slot_w(7));
        reduce_exchange(acc2, nullptr, ex(round & 1, 0));
        cluster_sync();
        TP();
        {
          float4 dh[2];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) dh[rr] = *reinterpret_cast<const float4*>(ex(round & 1, 0) + (2 * warp + rr) * ALD + 4 * lane);
          ln_bwd_rows(dh, XH1c, PERS + PE_G1, PERS + PE_BE1, rs1c, D2);
          ++round;
          __syncwarp();
          const float4 g2 = *reinterpret_cast<const float4*>(PERS + PE_AG2 + 4 * lane), be2 = *reinterpret_cast<const float4*>(PERS + PE_ABE2 + 4 * lane);
          float4 dh2[2];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int r = 2 * warp + rr;
            const float4 dz1 = *reinterpret_cast<const float4*>(D2 + r * ALD + 4 * lane);
            float da[NA], an[NA], dl[NA];
            float sdot = 0.f;
#pragma unroll
            for (int k = 0; k < NA; ++k) {
              da[k] = warp_sum(dot4(dz1, *reinterpret_cast<const float4*>(MISC + MI_W1ACT + k * HID + 4 * lane)));
              an[k] = MISC[MI_ANEW + r * 12 + k];
              sdot += an[k] * da[k];
            }
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NA; ++k) {
              dl[k] = an[k] * (da[k] - sdot);                    // softmax backward (the Gumbel noise is a constant)
              const float4 w3 = *reinterpret_cast<const float4*>(PERS + PE_AW3 + k * ALD + 4 * lane);
              d.x = fmaf(dl[k], w3.x, d.x); d.y = fmaf(dl[k], w3.y, d.y); d.z = fmaf(dl[k], w3.z, d.z); d.w = fmaf(dl[k], w3.w, d.w);
              if (lane == 0) MISC[MI_DL + r * 12 + k] = dl[k];
            }
            dh2[rr] = d;
          }
          __syncthreads();                                      // every warp has read its D2 rows ... D1 is rewritten below
          ln_bwd_rows(dh2, XH2a, PERS + PE_AG2, PERS + PE_ABE2, rs2a, D1);   // dz2 of the actor
        }
        __syncthreads();
        // vector gradients of the actor's layer 3 / LayerNorm 2
        ln_vec_grads(slab, la.b2, D1, XH2a, PERS + PE_AG2, PERS + PE_ABE2, [&](int r, int c) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < NA; ++k) s = fmaf(MISC[MI_DL + r * 12 + k], PERS[PE_AW3 + k * ALD + c], s);
          return s;
        });
        for (int e = tid; e < NA * CW + NA; e += THREADS) {      // dW3[k][c] = sum_r dl[r][k] * h2[r][c]; db3[k] = sum_r dl[r][k]
          float s = 0.f;
          if (e < NA * CW) {
            const int k = e / CW, c = col0 + (e - k * CW);
            for (int r = 0; r < RR; ++r) s += MISC[MI_DL + r * 12 + k] * fmaxf(fmaf(XH2a[r * ALD + c], PERS[PE_AG2 + c], PERS[PE_ABE2 + c]), 0.f);
            slab[la.w3 + k * HID + c] = s;
          } else if (rank == 0) {
            for (int r = 0; r < RR; ++r) s += MISC[MI_DL + r * 12 + e - NA * CW];
            slab[la.b3 + e - NA * CW] = s;
          }
        }
        // ---- round 6: dh1 (actor) -> LayerNorm 1 backward
        ready(8);
        TP();
        zero_acc(acc2);
        gemm_dh(acc2, D1, slot_w(8));
        reduce_exchange(acc2, nullptr, ex(round & 1, 0));
        dw_slice(slab + la.w2, HID, D1, H1a, ALD, HID);
        TP();
        cluster_sync();
        TP();
        {
          float4 dh[2];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) dh[rr] = *reinterpret_cast<const float4*>(ex(round & 1, 0) + (2 * warp + rr) * ALD + 4 * lane);
          ln_bwd_rows(dh, XH1a, PERS + PE_AG1, PERS + PE_ABE1, rs1a, D2);
          __syncthreads();
          const float* dhx = ex(round & 1, 0);
          ln_vec_grads(slab, la.b1, D2, XH1a, PERS + PE_AG1, PERS + PE_ABE1, [&](int r, int c) { return dhx[r * ALD + c]; });
          ++round;
        }
        dw_slice(slab + la.w1, O, D2, X + ag * O, XLD, O);
        TP();
        cp_wait<0>();
        TP();
        x_valid = false;                                        // the next update draws new rows
        }
      }
      if (ph == 1 || ph == 3) {
        // ======================================================================================== Adam + soft update
        const bool critic = ph == 1;
        const NetLayout& L = critic ? lc : la;
        const float lr = critic ? a.lr_c : a.lr_a;
        const int mode = a.adam_mode[critic ? 0 : 1];
        const int n4 = (L.total + 3) >> 2;
        float step_size[N2], bc2s[N2];
        for (int i = 0; i < N2; ++i) {
          const float stepf = step0[critic ? N2 + i : i] + (float)(u + 1);
          step_size[i] = lr / (1.0f - powf(a.beta1, stepf));
          bc2s[i] = sqrtf(1.0f - powf(a.beta2, stepf));
        }
        // four parameters per thread and step (every vector is 16-byte aligned, every network block padded to 4 floats; the
        // padding carries zero gradients and stays zero); both networks of the kind share the index space
        auto local_gradient = [&](int net, int i4) {             // the row blocks' slabs, added in row-block order
          const float* gp = ca.gpart[net] + 4 * i4;
          const long long gs = ca.gstride[net];
          float4 g = make_float4(0.f, 0.f, 0.f, 0.f), part[16];
          for (int r0 = 0; r0 < NRB; r0 += 16) {                 // the loads are issued together
#pragma unroll
            for (int r = 0; r < 16; ++r) if (r0 + r < NRB) part[r] = ldcg4(gp + (long long)(r0 + r) * gs);
#pragma unroll
            for (int r = 0; r < 16; ++r) if (r0 + r < NRB) { g.x += part[r].x; g.y += part[r].y; g.z += part[r].z; g.w += part[r].w; }
          }
          return g;
        };
        const bool peers = ca.world > 1;
        // Several ranks, exchange inside the kernel and inside this loop: the thread that owns four parameters adds its own
        // rank's slabs, PUSHES the result as two flag-in-data lines into its slot of every other rank's exchange block (posted
        // stores over NVLink), then polls ITS OWN block until the other ranks' lines of this exchange have arrived and adds the
        // world's gradients in rank order -- an all-gather + local sum with the same arithmetic on every rank (parameters stay
        // bit-identical), no barrier across GPUs, no fence, no NCCL call.  A slot is rewritten two exchanges later, which the
        // sender reaches only through grid barriers that every reader of the old value has passed.  From five ranks on the
        // owner of a float4 (its index mod world) receives the contributions, sums them in rank order and pushes the sum back
        // (reduce-scatter + all-gather: two hops, 2 (world - 1) / world gradients of traffic per rank instead of world - 1).
        const uint32_t ep = (uint32_t)(ca.epoch0 + 2ull * (unsigned long long)u + (critic ? 1ull : 2ull));
        const float gscale = peers ? 1.0f / (float)ca.world : a.grad_scale;
        for (int t4 = blockIdx.x * THREADS + tid; t4 < N2 * n4; t4 += n_ctas * THREADS) {
          const int i = t4 >= n4 ? 1 : 0, i4 = t4 - i * n4, net = critic ? N2 + i : i;
          const long long e = a.net_off[net] + 4 * i4;
          float4 g = make_float4(0.f, 0.f, 0.f, 0.f), m4, v4, p4, t4v;
          if (peers) {
            const float4 own = local_gradient(net, i4);
            const long long line = e >> 1;
            const unsigned long long t0 = global_ns();
            auto poll = [&](const uint4* src) {                     // two lines of this exchange -> four values
              uint4 l0, l1;
              unsigned spins = 0;
              for (;;) {
                l0 = ld_ll(src); l1 = ld_ll(src + 1);
                if (l0.y == ep && l0.w == ep && l1.y == ep && l1.w == ep) break;
                if ((++spins & 1023u) == 0 && global_ns() - t0 > ca.timeout_ns) { *ca.peer_err = 1u; break; }
              }
              return make_float4(__uint_as_float(l0.x), __uint_as_float(l0.z), __uint_as_float(l1.x), __uint_as_float(l1.z));
            };
            const int owner = ca.reduce_scatter ? t4 % ca.world : ca.rank;      // all-gather: every rank is the "owner" of everything
            if (!ca.reduce_scatter) {
#pragma unroll
              for (int p = 0; p < GW_MAX_PEERS; ++p)
                if (p < ca.world && p != ca.rank) {
                  st_ll(ca.peer_ll[p] + (long long)ca.rank * ca.ll_stride + line, own.x, own.y, ep);
                  st_ll(ca.peer_ll[p] + (long long)ca.rank * ca.ll_stride + line + 1, own.z, own.w, ep);
                }
            } else if (owner != ca.rank) {                          // this rank's contribution to the owner's slot
              st_ll(ca.peer_ll[owner] + (long long)ca.rank * ca.ll_stride + line, own.x, own.y, ep);
              st_ll(ca.peer_ll[owner] + (long long)ca.rank * ca.ll_stride + line + 1, own.z, own.w, ep);
            }
            if (mode & ADAM_APPLY) { m4 = ldcg4(a.M + e); v4 = ldcg4(a.V + e); p4 = ldcg4(a.P + e); t4v = ldcg4(a.T + e); }
            if (owner == ca.rank) {
              for (int q = 0; q < ca.world; ++q) {                  // the world's gradients in rank order
                const float4 x = q == ca.rank ? own : poll(ca.peer_ll[ca.rank] + (long long)q * ca.ll_stride + line);
                g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
              }
              if (ca.reduce_scatter) {                              // the sum goes back to everybody: result slot GW_MAX_PEERS
#pragma unroll
                for (int p = 0; p < GW_MAX_PEERS; ++p)
                  if (p < ca.world && p != ca.rank) {
                    st_ll(ca.peer_ll[p] + (long long)GW_MAX_PEERS * ca.ll_stride + line, g.x, g.y, ep);
                    st_ll(ca.peer_ll[p] + (long long)GW_MAX_PEERS * ca.ll_stride + line + 1, g.z, g.w, ep);
                  }
              }
            } else {
              g = poll(ca.peer_ll[ca.rank] + (long long)GW_MAX_PEERS * ca.ll_stride + line);
            }
          } else {
            if (mode & ADAM_APPLY) { m4 = ldcg4(a.M + e); v4 = ldcg4(a.V + e); p4 = ldcg4(a.P + e); t4v = ldcg4(a.T + e); }
            g = (mode & ADAM_FROM_G) ? ldcg4(a.G + e) : local_gradient(net, i4);
          }
          if (mode & ADAM_WRITE_G) st4(a.G + e, g);
          if (mode & ADAM_APPLY) {
            float gg[4] = {g.x, g.y, g.z, g.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
            float pp[4] = {p4.x, p4.y, p4.z, p4.w}, tt[4] = {t4v.x, t4v.y, t4v.z, t4v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float gk = gg[k] * gscale;
              mm[k] = mm[k] + (1.0f - a.beta1) * (gk - mm[k]);
              vv[k] = a.beta2 * vv[k] + (1.0f - a.beta2) * gk * gk;
              const float denom = sqrtf(vv[k]) / bc2s[i] + a.eps;
              pp[k] -= step_size[i] * (mm[k] / denom);
              tt[k] = tt[k] + a.tau * (pp[k] - tt[k]);
            }
            st4(a.M + e, make_float4(mm[0], mm[1], mm[2], mm[3])); st4(a.V + e, make_float4(vv[0], vv[1], vv[2], vv[3]));
            st4(a.P + e, make_float4(pp[0], pp[1], pp[2], pp[3])); st4(a.T + e, make_float4(tt[0], tt[1], tt[2], tt[3]));
          }
        }
        TP();
        if (blockIdx.x == 0 && tid < N2 && a.losses) {
          float v = 0.f;
          const float* lp = ca.lpart + (critic ? tid : N2 + tid) * NRB;
          for (int r = 0; r < NRB; ++r) v += __ldcg(lp + r);
          a.losses[((long long)u * 2 + (critic ? 1 : 0)) * N2 + tid] = v;
        }
      }
      if (!(u == a.updates - 1 && ph == ca.cp_end - 1)) grid_barrier(a.s.bar, n_ctas);
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    a.s.trace[ca.cp_end] = phase_clock();
    for (int k = 0; k < 2 * N2; ++k) {
      const int ph = k < N2 ? 3 : 1;
      if (ph >= ca.cp_begin && ph < ca.cp_end && (a.adam_mode[k < N2 ? 1 : 0] & ADAM_APPLY)) a.steps[k] = step0[k] + (float)a.updates;
    }
  }
  cluster_sync();                                                 // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace gwc

// ---------------------------------------------------------------------------------------------------- host side
bool gwc_supported(const gw_learner_config& c) {
  return c.n_agents == gwc::N2 && c.obs_len == gwc::O && c.action_dim == gwl::NA && c.batch % gwc::RR == 0;
}
int64_t gwc_scratch_floats(const gw_learner_config& c) {          // gradient slabs + loss partial sums
  if (!gwc_supported(c)) return 0;
  const int nrb = c.batch / gwc::RR;
  const gwl::NetLayout la = gwl::make_layout(c.obs_len, c.action_dim), lc = gwl::make_layout(c.n_agents * (c.obs_len + c.action_dim), 1);
  const int64_t sa = (la.total + 63) / 64 * 64, sc = (lc.total + 63) / 64 * 64;
  return (int64_t)c.n_agents * nrb * (sa + sc) + 64;
}

// grid size of the cluster kernel if all its clusters can be co-resident, else 0
int gwc_prepare(gw_learner* l) {
  const gw_learner_config& c = l->cfg;
  if (!gwc_supported(c)) return 0;
  const size_t smem = (size_t)gwc::SM_TOTAL * sizeof(float);
  if (cudaFuncSetAttribute(gwc::gw_learn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  const int clusters = c.n_agents * (c.batch / gwc::RR);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * gwc::CL));
  cfg.blockDim = dim3(gwl::THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = gwc::CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, gwc::gw_learn_cluster_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  l->cluster_max_active = max_clusters;
  if (max_clusters < clusters) return 0;
  // every co-resident cluster is launched: those without a row block spread the Adam phases over more SMs
  int launch = max_clusters;
  if (const char* e = getenv("GW_LEARN_HELPERS")) launch = clusters + atoi(e);
  if (launch > max_clusters) launch = max_clusters;
  if (launch < clusters) launch = clusters;
  return launch * gwc::CL;
}

int gwc_launch(gw_learner* l, const gwl::LearnArgs& a, int segment, cudaStream_t stream) {
  gwc::ClusterArgs ca;
  ca.a = a;
  const gw_learner_config& c = l->cfg;
  const int nrb = c.batch / gwc::RR;
  const int64_t sa = (a.la.total + 63) / 64 * 64, sc = (a.lc.total + 63) / 64 * 64;
  float* p = l->cluster_scratch;
  for (int k = 0; k < 2 * c.n_agents; ++k) {
    ca.gpart[k] = p;
    ca.gstride[k] = k < c.n_agents ? sa : sc;
    p += (int64_t)nrb * ca.gstride[k];
  }
  ca.lpart = p;
  ca.world = 1; ca.rank = 0; ca.epoch0 = 0; ca.timeout_ns = 5000000000ull; ca.peer_err = nullptr;
  for (int q = 0; q < GW_MAX_PEERS; ++q) ca.peer_ll[q] = nullptr;
  ca.ll_stride = 0; ca.reduce_scatter = 0;
  ca.split_roles = getenv("GW_LEARN_NO_SPLIT") == nullptr;
  if (l->peer_world > 1 && segment == GW_LEARN_ALL) {            // gradients exchanged inside the kernel over NVLink peer memory
    ca.world = l->peer_world; ca.rank = l->peer_rank; ca.epoch0 = l->peer_epochs;
    ca.ll_stride = (long long)(l->peer_flag_off / ((GW_MAX_PEERS + 1) * sizeof(uint4)));
    // all-gather moves (world - 1) gradients per rank in one hop; from five ranks on the owner-sums form (two hops, a quarter
    // of the traffic at eight ranks) is the cheaper one.  GW_PEER_PROTOCOL=ag|rs overrides (tests run both on two ranks).
    ca.reduce_scatter = l->peer_world > 4;
    if (const char* e = getenv("GW_PEER_PROTOCOL")) ca.reduce_scatter = e[0] == 'r';
    for (int q = 0; q < l->peer_world; ++q) ca.peer_ll[q] = static_cast<uint4*>(l->peer_base[q]);
    ca.peer_err = reinterpret_cast<unsigned int*>(static_cast<char*>(l->peer_base[l->peer_rank]) + l->peer_flag_off);
    if (const char* e = getenv("GW_PEER_TIMEOUT_MS")) ca.timeout_ns = 1000000ull * (unsigned long long)atoll(e);
    l->peer_epochs += 2ull * (unsigned long long)a.updates;
  }
  switch (segment) {
    case GW_LEARN_ALL: ca.cp_begin = 0; ca.cp_end = 4; break;
    case GW_LEARN_CRITIC_GRADS: ca.cp_begin = 0; ca.cp_end = 2; break;
    case GW_LEARN_ACTOR_GRADS: ca.cp_begin = 1; ca.cp_end = 4; break;
    default: ca.cp_begin = 3; ca.cp_end = 4; break;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)l->cluster_grid);
  cfg.blockDim = dim3(gwl::THREADS);
  cfg.dynamicSmemBytes = (size_t)gwc::SM_TOTAL * sizeof(float);
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = gwc::CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  // Profilers refuse the cooperative launch of a cluster kernel (ncu: LaunchFailed).  GW_LEARN_NO_COOP=1 drops the attribute:
  // the grid (<= 33 clusters, one CTA per SM) is co-resident on an otherwise idle GPU anyway, which is all the barrier needs.
  static const bool no_coop = getenv("GW_LEARN_NO_COOP") != nullptr;
  if (no_coop) cfg.numAttrs = 1;
  GW_CUDA(l->h, cudaLaunchKernelEx(&cfg, gwc::gw_learn_cluster_kernel, ca));
  return GW_OK;
}
