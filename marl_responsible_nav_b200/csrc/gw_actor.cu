// K5: actor forward for the batched rollout (AgileRL MADDPG.get_action as called at maddpg/agent.py:109-113).
//
// Per learner k: obs(160) -> Linear 128 -> LayerNorm -> ReLU -> Linear 128 -> LayerNorm -> ReLU -> Linear 9 ->
// GumbelSoftmax (+ Gaussian exploration noise, clip [0,1], action mask, arg-max).
//
// One CTA = 512 threads on tiles of up to 128 rows (envs) of one learner; all three layers run on the tensor cores:
//   layer 1  An observation is the constant map template plus <= 5 special cells, so W1*obs + b1 = c1 + W1 * delta with
//            c1 = b1 + W1 * template (fp32, exact) and delta non-zero in <= 5 cells.  The A operand is an all-zero
//            [128 x (cells + 16)] bf16 tile in shared memory in which each row's thread drops its <= 5 values (0.5, 1..5,
//            9.5, 10..14: exact in bf16) -- built from the 8-byte obs_code gw_step wrote; the 640-byte observation is
//            never read -- and takes them back one tile later.  cells / 16 + 1 tcgen05.mma (M=128, N=128, K=16, bf16 ->
//            fp32) against W1 (bf16, UMMA layout, one TMA bulk copy) accumulate in TMEM columns 0..127; the last K step
//            carries the bias: A columns of 1.0 against W1 columns holding c1 as a bf16 head + a bf16 remainder.
//   LN 1     warps that share a TMEM lane quarter (a warp's lanes are those of its id mod 4) split the 128 columns
//            (tcgen05.ld, thread = row) and make two passes over the accumulator, both in packed fp32 pairs (FADD2 /
//            FFMA2): the statistics (partial sums joined through shared memory by the warps that share the rows), then
//            normalise, affine, ReLU inside the bf16 conversion, and tcgen05.st of the bf16 pairs into TMEM columns
//            128..199: the A operand of layer 2 lives in tensor memory (lane = row, one 32-bit column per K pair), its
//            bias K step (1, 1, 0 ...) included.  A chunk's TMEM load is in flight while the previous one is processed.
//   layer 2  A2[128 x 144] (TMEM) x W2^T (shared memory): 9 tcgen05.mma with the A operand in TMEM, into columns 0..127
//            again (LayerNorm 1 has consumed the layer-1 accumulator).
//   LN 2     the same epilogue: LayerNorm, ReLU, bf16 pairs into TMEM columns 128..191 = the A operand of layer 3.
//   layer 3  A3[128 x 128] (TMEM) x W3p^T with W3 padded from 9 to 16 outputs: 8 tcgen05.mma with N = 16 into columns 0..15.
//   head     the first column group reads the 9 logits of its row (+ b3) and finishes Gumbel softmax / exploration noise /
//            clip / mask / arg-max.  The noise (9 Gumbel values, 9 normals from 5 Box-Muller pairs: 5 Philox calls per
//            row) is drawn and transformed by all threads while layer 1 runs; the last column group builds the next
//            tile's operand (obs_code fetched one tile ahead), packs its action mask into 9 bits and starts its layer 1
//            while the first one is still in the head.
// Small batches use half-full tiles (64 rows) so that the grid covers the machine; large ones keep two tiles in flight per
// CTA (two 8-warp groups out of phase, see Fixed<GROUPS>).
// sm_100a only (tcgen05 / TMEM); descriptors follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "gw_internal.h"

namespace gwa {

constexpr int HID = 128, NACT = GW_N_ACTIONS, ROWS = 128, MAX_CELLS = GW_MAX_H * GW_W;
// Layers 1 and 2 carry their bias through the GEMM: one more K step (16 columns) whose first two columns of A are 1.0 and
// whose first two columns of W hold the bias split into a bf16 head and a bf16 remainder (hi + lo = the fp32 value to
// 2^-17): the LayerNorm epilogues read the accumulator as it is, no bias vector in shared memory.
constexpr int KB = 16;
constexpr int D3_COL = 200;                        // TMEM column (within a group's 256) of the layer-3 accumulator
constexpr uint32_t ONES2 = 0x3F803F80u;             // bf16 (1.0, 1.0)

struct ActorParams {                     // device-resident, per learner
  alignas(16) __nv_bfloat16 w1_umma[HID * (MAX_CELLS + KB)];   // W1 [out n][in cell] in the canonical K-major core-matrix layout; columns cells, cells + 1: c1 (hi, lo)
  float c1[HID];                         // b1 + W1 * template row (fp32)
  float ln1_g[HID], ln1_b[HID];
  alignas(16) __nv_bfloat16 w2_umma[HID * (HID + KB)];   // W2 [out n][in k] in the canonical K-major core-matrix layout; columns 128, 129: b2 (hi, lo)
  float b2[HID], ln2_g[HID], ln2_b[HID];
  alignas(16) __nv_bfloat16 w3_umma[16 * HID];   // W3 [out o, padded to 16][in j] in the canonical K-major core-matrix layout
  float b3[NACT];
};

struct FwdArgs {
  const ActorParams* params;             // [n_learners]
  const unsigned long long* obs_code;    // [E]
  const int8_t* action_mask;             // [E, L, 9] or null
  float* cont;                           // [E, L, 9]
  int8_t* ids;                           // [E, L]
  long long E;
  long long env_id_base;                 // global id of row 0: the noise is keyed by the GLOBAL env id (independent of the sharding)
  int rows_per_tile;                      // 128, or 64 (half-full tiles) when the batch is too small to fill the machine
  int n, nl, kind, cpo;
  uint32_t apple_cells;
  int gumbel, explore;                   // Gumbel noise of the output activation; Gaussian exploration noise (training)
  float expl_noise, mean_noise;
  uint32_t seed_lo, seed_hi, step_lo, step_hi;
};

// byte offset of element (row r, k) in a [128 x 128] bf16 operand tile laid out as 8x(16 B) core matrices:
// K-chunk major, then 8-row group.  LBO (K-adjacent core matrices) = 2048 B, SBO (adjacent row groups) = 128 B.
__host__ __device__ __forceinline__ uint32_t umma_off(int r, int k) {
  return (uint32_t)(((k >> 3) * 16 + (r >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// the same for a 16-row operand (W3 padded from 9 to 16 outputs): two 8-row groups per K chunk, LBO = 256 B
__host__ __device__ __forceinline__ uint32_t umma_off16(int r, int k) {
  return (uint32_t)(((k >> 3) * 2 + (r >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes = 2048u) {
  // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version = 1 [46,48), SWIZZLE_NONE
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
         (1ull << 46);
}

// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format BF16 (1) [7,10)/[10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
constexpr uint32_t IDESC_N16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);   // layer 3

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}

// CTA size.  (1 024 threads for the two-tile kernel -- 16 warps per tile in flight, 64 registers -- measured 15 % slower
// at 1 M envs: the phases only one column group runs, operand build and head, do not get shorter.)
__host__ __device__ constexpr int threads_of(int /*groups*/) { return 512; }

// GROUPS = 1: the CTA's 16 warps work on one tile (four column groups in the epilogues) -- shortest path for one tile,
// used while the batch does not fill the machine.  GROUPS = 2: two groups of 8 warps work on two tiles out of phase (two
// column groups each, own operand tile, own TMEM columns, own barriers), sharing W1 / W2: while one group waits for its
// MMAs or barriers the other one has the issue slots.
// Dynamic shared memory: [GROUPS x a_tile: 128 x (cells + 16) bf16, the layer-1 operand] [w1: 128 x (cells + 16) bf16]
// [Fixed<GROUPS>].  TMEM per group (256 columns): 0..127 the accumulator of layer 1, then of layer 2 (layer 3: 0..15);
// 128..199 the A operand of layer 2 (144 K values as bf16 pairs), then 128..191 that of layer 3.
template <int GROUPS>
struct Fixed {
  static constexpr int NGRP = threads_of(GROUPS) / GROUPS / 128;         // column groups of the epilogues (warps sharing a TMEM lane quarter)
  alignas(1024) uint8_t w2[HID * (HID + KB) * 2];
  alignas(128) uint8_t w3[16 * HID * 2];             // W3 padded to 16 outputs, bf16, UMMA layout
  float b3[NACT];
  alignas(16) float ln1_g[HID], ln1_b[HID], ln2_g[HID], ln2_b[HID];
  alignas(8) float2 part[GROUPS][NGRP][ROWS];       // LayerNorm partial (sum, sum of squares) per column group and row
  float noise[GROUPS][2][ROWS][19];                 // the head's noise per row: 9 Gumbel values, 9 standard normals (odd stride: no bank
                                                    // conflicts), double-buffered over consecutive tiles like maskbits
  uint16_t maskbits[GROUPS][2][ROWS];               // action-mask bits per row, double-buffered over consecutive tiles
  alignas(8) unsigned long long bar[GROUPS], bar1[GROUPS], bar_w; // completion of layers 2 / 3 and of layer 1 per group; arrival of W1 / W2
  uint32_t tmem_base;
};
// K extent of the layer-1 operand tile (the only operand in shared memory): cells + the bias K step
__host__ __device__ inline int op_k(int cells) { return cells + KB; }
template <int GROUPS>
__host__ __device__ inline size_t smem_bytes(int cells) { return (size_t)(GROUPS + 1) * ROWS * op_k(cells) * 2 + sizeof(Fixed<GROUPS>) + 1024; }

// uniform in (0, 1) from the top 23 bits of a word, without an integer-to-float conversion: [1, 2) - (1 - 2^-24)
__device__ __forceinline__ float unit_from(uint32_t w) { return __uint_as_float(0x3F800000u | (w >> 9)) - 0.99999994f; }
__device__ __forceinline__ float gumbel_from(uint32_t w) {
  return -__logf(-__logf(unit_from(w)) + 1e-20f);
}
__device__ __forceinline__ float2 gauss_pair_from(uint32_t a, uint32_t b) {     // Box-Muller, both outputs
  const float u1 = unit_from(a), u2 = unit_from(b);
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));      // noise: MUFU.SQRT is plenty
  float sn, cs;
  __sincosf(6.283185307f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// k_steps MMAs of K = 16: A is a 128-row operand tile, B has `b_rows` (= N) rows
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Packed fp32 pairs (sm_100: add / fma on two floats in a 64-bit register pair -> FADD2 / FFMA2): the LayerNorm epilogues
// are the kernel's CUDA-core work, this halves their arithmetic instructions.
__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) { return pack2(__float_as_uint(lo), __float_as_uint(hi)); }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  uint32_t a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
  lo = __uint_as_float(a); hi = __uint_as_float(b);
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// (lo, hi) -> max(., 0) -> bf16 pair, lo in the low half: ReLU rides on the conversion
__device__ __forceinline__ uint32_t relu_bf16x2(uint64_t v) {
  uint32_t a, b, d;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
  return d;
}

// tcgen05.ld without the wait, and the wait with the registers as in/out operands (what follows depends on it): a
// chunk's load is in flight while the previous chunk is processed
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

// 8 packed bf16 pairs (16 K columns of this thread's row) into a TMEM-resident A operand: lane = row, one 32-bit
// column per pair
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One LayerNorm + ReLU epilogue of this thread's CPG columns of its row: TMEM accumulator (the layer's bias is already in
// it, see KB) -> statistics (partial sums of the column groups joined through shared memory; `sync` joins the warps that
// share the rows) -> second pass over the accumulator (re-read from TMEM: keeping the row in registers across the barrier
// spills at 128 registers per thread and measured 40 % slower) -> normalise, affine, ReLU, bf16 -> the next layer's A
// operand, which lives in TMEM (bf16 pairs, `a_taddr` = its first column in this warp's lane quarter): no shared-memory
// store, no shared-memory read by the MMA; ONES: this thread also sets the next layer's bias columns (K = 128..143:
// 1, 1, 0 ...) of its row.
// Rows without an env are processed like any other (their accumulator rows are zero: zero operand rows) and their
// results are never read: a row of A only reaches the same row of D.
template <int CPG, int NGRP, class Sync>
__device__ __forceinline__ void ln_relu_epilogue(uint32_t taddr, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 float2 (*part)[ROWS], int cg, int m, uint32_t a_taddr, int col0, bool on,
                                                 bool ones, Sync&& sync) {
  static_assert(CPG % 32 == 0, "two chunks of 16 columns in flight");
  if (on) {
    uint64_t sa2 = 0ull, sb2 = 0ull, qa2 = 0ull, qb2 = 0ull;     // (0.f, 0.f): two chains each for the sum and the sum of squares
    uint32_t ra[16], rb[16];
    auto stats = [&](const uint32_t (&r)[16]) {
#pragma unroll
      for (int u = 0; u < 16; u += 4) {
        const uint64_t x0 = pack2(r[u], r[u + 1]), x1 = pack2(r[u + 2], r[u + 3]);
        sa2 = add2(sa2, x0); qa2 = fma2(x0, x0, qa2);
        sb2 = add2(sb2, x1); qb2 = fma2(x1, x1, qb2);
      }
    };
    tmem_ld16_issue(taddr, ra);
#pragma unroll
    for (int c0 = 0; c0 < CPG; c0 += 32) {
      tmem_wait16(ra);
      tmem_ld16_issue(taddr + (uint32_t)(c0 + 16), rb);
      stats(ra);
      tmem_wait16(rb);
      if (c0 + 32 < CPG) tmem_ld16_issue(taddr + (uint32_t)(c0 + 32), ra);
      stats(rb);
    }
    float s0, s1, q0, q1;
    unpack2(add2(sa2, sb2), s0, s1); unpack2(add2(qa2, qb2), q0, q1);
    part[cg][m] = make_float2(s0 + s1, q0 + q1);
  }
  sync();                                                // partial sums of the column groups
  if (on) {
    uint32_t ra[16], rb[16];
    tmem_ld16_issue(taddr, ra);                          // in flight while the statistics are joined
    if (ones) {                                          // K = 128..143 of the row: (1, 1, 0, ...)
      const uint32_t o8[8] = {ONES2, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st8(a_taddr + (uint32_t)(HID / 2), o8);
    }
    float su = 0.f, sq = 0.f;
#pragma unroll
    for (int g = 0; g < NGRP; ++g) { const float2 pp = part[g][m]; su += pp.x; sq += pp.y; }
    const float mu = su * (1.0f / HID);
    const float rs = rsqrtf(fmaxf(sq * (1.0f / HID) - mu * mu, 0.f) + 1e-5f);
    const float nmr = -mu * rs;
    const uint64_t rs2 = pack2f(rs, rs), nmr2 = pack2f(nmr, nmr);
    auto norm = [&](const uint32_t (&r)[16], int c0) {
      uint32_t pk[8];
#pragma unroll
      for (int u = 0; u < 16; u += 8) {
        const int c8 = c0 + u;
        const float4 g0 = *reinterpret_cast<const float4*>(&gamma[c8]), g1 = *reinterpret_cast<const float4*>(&gamma[c8 + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&beta[c8]), b1 = *reinterpret_cast<const float4*>(&beta[c8 + 4]);
        // (x - mu) * rs as one fma, then the affine part; ReLU inside the conversion
        const uint32_t p0 = relu_bf16x2(fma2(fma2(pack2(r[u], r[u + 1]), rs2, nmr2), pack2f(g0.x, g0.y), pack2f(b0.x, b0.y)));
        const uint32_t p1 = relu_bf16x2(fma2(fma2(pack2(r[u + 2], r[u + 3]), rs2, nmr2), pack2f(g0.z, g0.w), pack2f(b0.z, b0.w)));
        const uint32_t p2 = relu_bf16x2(fma2(fma2(pack2(r[u + 4], r[u + 5]), rs2, nmr2), pack2f(g1.x, g1.y), pack2f(b1.x, b1.y)));
        const uint32_t p3 = relu_bf16x2(fma2(fma2(pack2(r[u + 6], r[u + 7]), rs2, nmr2), pack2f(g1.z, g1.w), pack2f(b1.z, b1.w)));
        pk[u / 2] = p0; pk[u / 2 + 1] = p1; pk[u / 2 + 2] = p2; pk[u / 2 + 3] = p3;
      }
      tmem_st8(a_taddr + (uint32_t)((col0 + c0) / 2), pk);
    };
#pragma unroll
    for (int c0 = 0; c0 < CPG; c0 += 32) {
      tmem_wait16(ra);
      tmem_ld16_issue(taddr + (uint32_t)(c0 + 16), rb);
      norm(ra, c0);
      tmem_wait16(rb);
      if (c0 + 32 < CPG) tmem_ld16_issue(taddr + (uint32_t)(c0 + 32), ra);
      norm(rb, c0 + 16);
    }
    tmem_st_wait();
  }
}

// k_steps MMAs of K = 16 with the A operand in TMEM (bf16 pairs, 8 columns per K step, lane = row) and B in shared memory
__device__ __forceinline__ void mma_k16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_addr, int k_steps, uint32_t idesc = IDESC,
                                           uint32_t b_rows = 128u) {
  const uint32_t b_lbo = b_rows * 16u;
  for (int kk = 0; kk < k_steps; ++kk) {
    const uint64_t db = smem_desc(b_addr + kk * 2 * b_lbo, b_lbo);
    const uint32_t acc = kk > 0 ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a + (uint32_t)(kk * 8)),
        "l"(db), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
  }
}

__device__ __forceinline__ void mma_k16(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, int k_steps, uint32_t idesc = IDESC,
                                        uint32_t b_rows = 128u) {
  const uint32_t b_lbo = b_rows * 16u;                     // bytes between K-adjacent core matrices of B
  for (int kk = 0; kk < k_steps; ++kk) {
    const uint64_t da = smem_desc(a_addr + kk * 4096), db = smem_desc(b_addr + kk * 2 * b_lbo, b_lbo);
    const uint32_t acc = kk > 0 ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  }
}

template <int GROUPS>
__global__ void __launch_bounds__(threads_of(GROUPS), 1) actor_forward_kernel(FwdArgs a) {
  constexpr int THREADS = threads_of(GROUPS), TPG = THREADS / GROUPS, TSH = TPG == 512 ? 2 : 1,   // log2(TPG / ROWS)
                 NGRP = Fixed<GROUPS>::NGRP, CPG = HID / NGRP;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int cells = a.cpo;                                 // K of layer 1: a multiple of 16 (H x 16)
  const uint32_t op_bytes = (uint32_t)ROWS * op_k(cells) * 2;     // operand tile incl. the bias K step
  const uint32_t w1_bytes = (uint32_t)ROWS * (cells + KB) * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gi = tid / TPG, lt = tid % TPG;                // group, thread within the group
  uint8_t* const a_tile = smem_raw + (size_t)gi * op_bytes;
  uint8_t* const w1_tile = smem_raw + (size_t)GROUPS * op_bytes;
  Fixed<GROUPS>& s = *reinterpret_cast<Fixed<GROUPS>*>(smem_raw + (size_t)(GROUPS + 1) * op_bytes);
  auto group_sync = [&]() {                                // barrier of this group's TPG threads (id 0 is __syncthreads)
    if (GROUPS == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + gi), "n"(TPG) : "memory");
  };
  // barrier of the NGRP warps that share this thread's rows (same TMEM lane quarter of the same group)
  auto rows_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(3 + gi * 4 + (warp & 3)), "n"(NGRP * 32) : "memory"); };
  const int k = blockIdx.y;                               // learner
  const ActorParams& P = a.params[k];
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s.bar[gi]), bar1 = (uint32_t)__cvta_generic_to_shared(&s.bar1[gi]),
                 bar_w = (uint32_t)__cvta_generic_to_shared(&s.bar_w);
  const int RT = a.rows_per_tile;

  const int q = warp & 3, cg = (lt >> 5) >> 2;             // TMEM lane quarter (warp id mod 4: TPG is a multiple of 128), column group
  const int m = q * 32 + lane;                             // this thread's row of the tile
  const bool quarter_on = q * 32 < RT;                     // warp-uniform: a half-full tile fills lane quarters 0 and 1
  const long long n_tiles = (a.E + RT - 1) / RT;
  const long long tile0 = (long long)blockIdx.x * GROUPS + gi, tile_step = (long long)gridDim.x * GROUPS;
  // Column group OPG builds the next tile's layer-1 operand and starts layer 1 while column group 0 is still in the head
  // of the previous tile; the row's obs_code is fetched one tile ahead (its DRAM latency hides behind the tile in flight;
  // the first one is issued before the one-time setup)
  constexpr int OPG = NGRP - 1;
  unsigned long long code_next = 0ull;
  if (cg == OPG && quarter_on && m < RT && tile0 < n_tiles && tile0 * RT + m < a.E) code_next = a.obs_code[tile0 * RT + m];

  // ---- one-time setup: 256 TMEM columns per group (accumulator + the operands of layers 2 and 3), mbarriers, W1 / W2 by TMA, the small vectors
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&s.tmem_base)), "n"(GROUPS * 2 * HID));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int g = 0; g < GROUPS; ++g) { mbar_init((uint32_t)__cvta_generic_to_shared(&s.bar[g]), 1); mbar_init((uint32_t)__cvta_generic_to_shared(&s.bar1[g]), 1); }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w), "r"(w1_bytes + (uint32_t)(HID * (HID + KB) * 2 + 16 * HID * 2)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(w1_tile)), "l"(P.w1_umma), "r"(w1_bytes), "r"(bar_w) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(s.w2)), "l"(P.w2_umma), "n"(HID * (HID + KB) * 2), "r"(bar_w) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(s.w3)), "l"(P.w3_umma), "n"(16 * HID * 2), "r"(bar_w) : "memory");
  }
  {
    for (int i = lt; i < (int)(op_bytes / 16); i += TPG) reinterpret_cast<uint4*>(a_tile)[i] = make_uint4(0, 0, 0, 0);
    if (tid < NACT) s.b3[tid] = P.b3[tid];
    if (tid < HID) {
      s.ln1_g[tid] = P.ln1_g[tid]; s.ln1_b[tid] = P.ln1_b[tid];
      s.ln2_g[tid] = P.ln2_g[tid]; s.ln2_b[tid] = P.ln2_b[tid];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_all = s.tmem_base;
  const uint32_t tmem = tmem_all + (uint32_t)(gi * 2 * HID);   // this group's columns
  const uint32_t a_addr = (uint32_t)__cvta_generic_to_shared(a_tile);
  const uint32_t w1_addr = (uint32_t)__cvta_generic_to_shared(w1_tile);
  const uint32_t w2_addr = (uint32_t)__cvta_generic_to_shared(s.w2);
  const uint32_t w3_addr = (uint32_t)__cvta_generic_to_shared(s.w3);
  uint32_t phase = 0;
  bool w_pending = true;

  const int col0 = cg * CPG;
  const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
  auto op_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(11 + gi), "n"(128) : "memory"); };
  int par = 0;                                             // tile parity of this group (maskbits / noise buffer)
  uint32_t prev_cells = 0u;                                // the cells this row's thread set in the operand of the previous tile
  bool prev_live = false;
  // ---- start of a tile (column group OPG only): the layer-1 operand -- the row's special cells, delta against the
  // template, which is 0 on every active cell --, the row's action mask for the head, and layer 1 on the tensor cores.
  // Called for the first tile before the loop and for tile t+1 between LayerNorm 2 and layer 3 of tile t: layer 1 of the
  // next tile runs under layer 3, the head and the noise phase, and is long complete when LayerNorm 1 asks for it.
  auto start_tile = [&](long long tile, int par) {
    if (cg != OPG) return;
    const long long e = tile * RT + m;
    const bool live = quarter_on && m < RT && e < a.E;
    {
    const unsigned long long code = code_next;
    if (quarter_on && m < RT) {
      const long long en = e + tile_step * RT;
      if (tile + tile_step < n_tiles && en < a.E) code_next = a.obs_code[en];
    }
    // ... and the row's action mask for the head (9 bytes -> 9 bits)
    if (live && a.action_mask) {
      const int8_t* mk = a.action_mask + (e * a.nl + k) * NACT;
      uint32_t bits = 0u;
#pragma unroll
      for (int o = 0; o < NACT; ++o) bits |= (mk[o] != 0 ? 1u : 0u) << o;
      s.maskbits[gi][par][m] = (uint16_t)bits;
    }
    // The operand tile belongs to layer 1 alone (the operands of layers 2 and 3 live in TMEM): instead of clearing it,
    // the row's thread takes back the few cells it set for the previous tile (layer 1 of that tile completed long ago)
    const uint32_t apple = (a.kind == GW_ENV_MULTI) ? (a.apple_cells >> (8 * k)) & 0xFFu : a.apple_cells & 0xFFu;
    if (prev_live) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t c = (prev_cells >> (8 * i)) & 0xFFu;
        if (i < a.n && (int)c < cells) *reinterpret_cast<uint16_t*>(a_tile + umma_off(m, (int)c)) = 0;
      }
      if ((int)apple < cells) *reinterpret_cast<uint16_t*>(a_tile + umma_off(m, (int)apple)) = 0;
    }
    prev_live = live;
    prev_cells = (uint32_t)code;
    if (live) {
      const uint32_t cellsw = (uint32_t)code, apples = (uint32_t)(code >> 32) & 3u;
      const bool fresh = (code >> 34) & 1ull;
      const bool apple_on = (a.kind == GW_ENV_MULTI) ? ((apples >> k) & 1u) : (apples & 1u);
      bool covered = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t c = (cellsw >> (8 * i)) & 0xFFu;
        const bool here = i < a.n && apple_on && c == apple;
        covered |= here;
        float v;                                           // same value rules as the renderer (ma_customenv.py:303-322)
        if (fresh) v = 0.5f;
        else if (here || a.kind == GW_ENV_SINGLE) v = (float)(i + 1);
        else v = (i == k) ? 1.0f : 5.0f;
        if (i < a.n && (int)c < cells) *reinterpret_cast<__nv_bfloat16*>(a_tile + umma_off(m, (int)c)) = __float2bfloat16(here ? v + 9.0f : v);
      }
      if (apple_on && !covered && (int)apple < cells) *reinterpret_cast<__nv_bfloat16*>(a_tile + umma_off(m, (int)apple)) = __float2bfloat16(9.0f);
      *reinterpret_cast<uint32_t*>(a_tile + umma_off(m, cells)) = ONES2;       // the bias columns: c1 = hi + lo
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // operand -> visible to the tensor-core proxy
    asm volatile("tcgen05.fence::before_thread_sync;");
    op_sync();                                                       // the four warps of this column group
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- layer 1 on the tensor cores: D0[128x128] (TMEM columns 0..127) = A[128 x cells] * W1^T
    if (lt == OPG * 128) {
      if (w_pending) mbar_wait(bar_w, 0u);
      mma_k16(tmem, a_addr, w1_addr, cells / 16 + 1);
      // completion of all prior MMAs arrives on the mbarrier (implies tcgen05.fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar1) : "memory");
    }
    }
  };
  if (tile0 < n_tiles) start_tile(tile0, 0);
  uint32_t phase1 = 0;
  for (long long tile = tile0; tile < n_tiles; tile += tile_step, par ^= 1) {
    const long long e = tile * RT + m;
    const bool live = quarter_on && m < RT && e < a.E;

    w_pending = false;
    // ---- while layer 1 may still run: the head's noise (9 Gumbel values + 5 Box-Muller pairs per row, a pure function of seed / env /
    // learner / step), TPG / RT threads per row; Philox call c of a row yields: c = 0, 1: Gumbel 4c..4c+3; c = 2: Gumbel 8
    // and normals 0, 1; c = 3: normals 2..5; c = 4: normals 6..8
    if (a.gumbel | a.explore) {
      // RT is ROWS or ROWS / 2: no runtime division.  One tile per CTA (latency regime): the column group that builds the
      // operand waits for obs_code and the weights first, so the other three share the rows (3 or 6 threads per row)
      int tpr, rr, c_first;
      if (GROUPS == 1) {
        if (RT == ROWS) { tpr = 3; rr = lt / 3; c_first = lt - rr * 3; }
        else { tpr = 6; rr = lt / 6; c_first = lt - rr * 6; }
        if (cg == OPG) c_first = 8;                        // no call
      } else {
        const int tsh = RT == ROWS ? TSH : TSH + 1;
        tpr = 1 << tsh; rr = lt >> tsh; c_first = lt & (tpr - 1);
      }
      const long long er = tile * RT + rr;
      const int n_calls = a.explore ? 5 : 3;
      if (er < a.E)
        for (int c = c_first; c < n_calls; c += tpr) {
          const unsigned long long ge = (unsigned long long)(a.env_id_base + er);
          uint32_t w[4] = {(uint32_t)ge, (uint32_t)(ge >> 32) ^ ((uint32_t)k << 24) ^ ((uint32_t)c << 28),
                           a.step_lo, a.step_hi ^ 0xAC70u};
          gw::philox4x32(w, a.seed_lo, a.seed_hi);
          float* nz = s.noise[gi][par][rr];
          if (c < 2) {
#pragma unroll
            for (int u = 0; u < 4; ++u) nz[4 * c + u] = gumbel_from(w[u]);
          } else if (c == 2) {
            nz[8] = gumbel_from(w[0]);
            if (a.explore) { const float2 g = gauss_pair_from(w[2], w[3]); nz[9] = g.x; nz[10] = g.y; }
          } else {
            const float2 g0 = gauss_pair_from(w[0], w[1]), g1 = gauss_pair_from(w[2], w[3]);
            nz[4 * c - 1] = g0.x; nz[4 * c] = g0.y; nz[4 * c + 1] = g1.x;
            if (c == 3) nz[4 * c + 2] = g1.y;
          }
        }
    }
    mbar_wait(bar1, phase1);                            // layer 1 of this tile (issued one tile ago)
    phase1 ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- LayerNorm 1 + ReLU -> operand of layer 2 (this thread: CPG columns of its row)
    ln_relu_epilogue<CPG, NGRP>(lane_addr + (uint32_t)col0, &s.ln1_g[col0], &s.ln1_b[col0], s.part[gi], cg, m,
                                lane_addr + (uint32_t)HID, col0, quarter_on, cg == 0, rows_sync);
    asm volatile("tcgen05.fence::before_thread_sync;");
    group_sync();
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- layer 2: D[128x128] (TMEM columns 0..127 again: LayerNorm 1 has consumed the layer-1 accumulator) =
    // A2[128 x 144] (TMEM columns 128..199) * W2^T.  (Rows RT..127 of a half-full tile hold whatever TMEM held there: they
    // only reach accumulator rows nobody reads.)
    if (lt == 0) {
      mma_k16_ts(tmem, tmem + (uint32_t)HID, w2_addr, HID / 16 + 1);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- LayerNorm 2 + ReLU -> operand of layer 3 (again this thread's columns of its row, again TMEM columns 128..191)
    ln_relu_epilogue<CPG, NGRP>(lane_addr + (uint32_t)col0, &s.ln2_g[col0], &s.ln2_b[col0], s.part[gi], cg, m,
                                lane_addr + (uint32_t)HID, col0, quarter_on, false, rows_sync);
    asm volatile("tcgen05.fence::before_thread_sync;");
    group_sync();
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- the next tile starts here: its layer 1 may overwrite columns 0..127 (LayerNorm 2 has consumed them)
    if (tile + tile_step < n_tiles) start_tile(tile + tile_step, par ^ 1);

    // ---- layer 3: D[128 x 16] (TMEM columns 200..215, clear of both the accumulator and the operands) = A3[128x128] * W3p^T
    if (lt == 0) {
      mma_k16_ts(tmem + (uint32_t)D3_COL, tmem + (uint32_t)HID, w3_addr, HID / 16, IDESC_N16, 16u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;");
    float logit[NACT];
    if (quarter_on && cg == 0) {
      uint32_t r[16];
      tmem_ld16(lane_addr + (uint32_t)D3_COL, r);
#pragma unroll
      for (int o = 0; o < NACT; ++o) logit[o] = __uint_as_float(r[o]) + s.b3[o];
    }
    if (live && cg == 0) {
      // GumbelSoftmax head (the reference's output activation draws fresh Gumbel noise on EVERY forward, evaluation
      // included) + Gaussian exploration noise (training), both drawn above
      if (a.gumbel) {
#pragma unroll
        for (int o = 0; o < NACT; ++o) logit[o] += s.noise[gi][par][m][o];
      }
      float mx = logit[0];
#pragma unroll
      for (int o = 1; o < NACT; ++o) mx = fmaxf(mx, logit[o]);
      float den = 0.f, pr[NACT];
#pragma unroll
      for (int o = 0; o < NACT; ++o) { pr[o] = __expf(logit[o] - mx); den += pr[o]; }
      const float inv = 1.0f / den;
      int best = 0;
      float best_v = -1e30f;
      const uint32_t mbits = a.action_mask ? (uint32_t)s.maskbits[gi][par][m] : 0x1FFu;
      float* out = a.cont + (e * a.nl + k) * NACT;
#pragma unroll
      for (int o = 0; o < NACT; ++o) {
        float v = pr[o] * inv;
        if (a.explore) v = fminf(fmaxf(v + a.mean_noise + a.expl_noise * s.noise[gi][par][m][9 + o], 0.f), 1.f);
        out[o] = v;
        const bool ok = (mbits >> o) & 1u;
        if (ok && v > best_v) { best_v = v; best = o; }
      }
      a.ids[e * a.nl + k] = (int8_t)best;
    }
    // No barrier at the end of a tile: column group 0 joins the next one at its noise phase, the others are already
    // there.  maskbits / noise of this tile are rewritten two tiles on and the logit columns by the next tile's layer 3,
    // the operand tile and part only behind barriers that every warp reaches after it is done with them.
  }
  if (w_pending && tid == 0) mbar_wait(bar_w, 0u);         // a CTA whose first group has no tile must not exit with the copies in flight

  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_all), "n"(GROUPS * 2 * HID));
}

// bias of output n as the two bf16 columns k0, k0 + 1 of a packed weight matrix: head + remainder
__host__ __device__ inline void put_bias(__nv_bfloat16* w_umma, int n, int k0, float b) {
  const __nv_bfloat16 hi = __float2bfloat16(b);
  w_umma[umma_off(n, k0) / 2] = hi;
  w_umma[umma_off(n, k0 + 1) / 2] = __float2bfloat16(b - __bfloat162float(hi));
}

// Weights straight from the trainer's device tensors (torch layout, fp32) into the kernel's packed form: the same
// arithmetic as pack_weights below (c1 accumulated in double, in cell order), without the trip through the host.
struct PackArgs {
  gw_actor_weights w[GW_MAX_LEARNERS];   // DEVICE pointers
  ActorParams* out;                      // [n_learners]
  uint16_t map_rows[GW_MAX_H];
  int cpo;
};

__global__ void __launch_bounds__(256) actor_pack_kernel(PackArgs a) {
  const gw_actor_weights& W = a.w[blockIdx.y];
  ActorParams& P = a.out[blockIdx.y];
  const int cpo = a.cpo, stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cpo * HID; i += stride) {
    const int j = i / cpo, cell = i % cpo;
    P.w1_umma[umma_off(j, cell) / 2] = __float2bfloat16(W.w1[i]);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HID * HID; i += stride) {
    const int n = i / HID, kk = i % HID;
    P.w2_umma[umma_off(n, kk) / 2] = __float2bfloat16(W.w2[i]);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NACT * HID; i += stride)      // rows 9..15 stay zero (gw_actor_create)
    P.w3_umma[umma_off16(i / HID, i % HID) / 2] = __float2bfloat16(W.w3[i]);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < HID; j += stride) {
    double acc = W.b1[j];
    for (int cell = 0; cell < cpo; ++cell)
      if (!((a.map_rows[cell >> 4] >> (cell & 15)) & 1)) acc -= (double)W.w1[j * cpo + cell];   // template value -1 on inactive cells
    P.c1[j] = (float)acc;
    put_bias(P.w1_umma, j, cpo, (float)acc);
    put_bias(P.w2_umma, j, HID, W.b2[j]);
    P.ln1_g[j] = W.ln1_g[j]; P.ln1_b[j] = W.ln1_b[j];
    P.b2[j] = W.b2[j]; P.ln2_g[j] = W.ln2_g[j]; P.ln2_b[j] = W.ln2_b[j];
    if (j < NACT) P.b3[j] = W.b3[j];
  }
}

}  // namespace gwa

// ====================================================================== host side / C-ABI
struct gw_actor {
  gw_handle* h = nullptr;
  gwa::ActorParams* d_params = nullptr;
  int n_learners = 0;
};

static int pack_weights(gw_handle* h, const gw_actor_weights* w, int nl, std::vector<gwa::ActorParams>& host) {
  const gw_config& c = h->cfg;
  const int cpo = c.height * GW_W;
  host.assign(nl, gwa::ActorParams());
  for (int k = 0; k < nl; ++k) {
    const gw_actor_weights& W = w[k];
    if (!W.w1 || !W.b1 || !W.ln1_g || !W.ln1_b || !W.w2 || !W.b2 || !W.ln2_g || !W.ln2_b || !W.w3 || !W.b3)
      return gw_fail(h, GW_EINVAL, "gw_actor: null weight pointer");
    gwa::ActorParams& P = host[k];
    std::memset(&P, 0, sizeof(P));
    for (int j = 0; j < gwa::HID; ++j) {
      double acc = W.b1[j];
      for (int cell = 0; cell < cpo; ++cell) {
        P.w1_umma[gwa::umma_off(j, cell) / 2] = __float2bfloat16(W.w1[j * cpo + cell]);
        const bool active = (c.map_rows[cell >> 4] >> (cell & 15)) & 1;
        if (!active) acc -= (double)W.w1[j * cpo + cell];              // template value -1 on inactive cells, 0 elsewhere
      }
      P.c1[j] = (float)acc;
      gwa::put_bias(P.w1_umma, j, cpo, (float)acc);
      gwa::put_bias(P.w2_umma, j, gwa::HID, W.b2[j]);
      P.ln1_g[j] = W.ln1_g[j]; P.ln1_b[j] = W.ln1_b[j];
      P.b2[j] = W.b2[j]; P.ln2_g[j] = W.ln2_g[j]; P.ln2_b[j] = W.ln2_b[j];
    }
    for (int n = 0; n < gwa::HID; ++n)
      for (int kk = 0; kk < gwa::HID; ++kk)
        P.w2_umma[gwa::umma_off(n, kk) / 2] = __float2bfloat16(W.w2[n * gwa::HID + kk]);
    for (int o = 0; o < gwa::NACT; ++o) {
      P.b3[o] = W.b3[o];
      for (int j = 0; j < gwa::HID; ++j) P.w3_umma[gwa::umma_off16(o, j) / 2] = __float2bfloat16(W.w3[o * gwa::HID + j]);
    }
  }
  return GW_OK;
}

extern "C" {

int gw_actor_create(gw_handle* h, const gw_actor_weights* weights, int n_learners, gw_actor** out) {
  if (!h || !weights || !out) return gw_fail(h, GW_EINVAL, "gw_actor_create: null argument");
  if (n_learners != h->cfg.n_learners) return gw_fail(h, GW_EINVAL, "gw_actor_create: one weight set per learner is required");
  std::vector<gwa::ActorParams> host;
  if (int rc = pack_weights(h, weights, n_learners, host)) return rc;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  gw_actor* a = new (std::nothrow) gw_actor();
  if (!a) return gw_fail(h, GW_ENOMEM, "gw_actor_create: host allocation failed");
  a->h = h;
  a->n_learners = n_learners;
  cudaError_t e = cudaMalloc(&a->d_params, sizeof(gwa::ActorParams) * n_learners);
  if (e == cudaSuccess) e = cudaMemcpy(a->d_params, host.data(), sizeof(gwa::ActorParams) * n_learners, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gwa::actor_forward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)gwa::smem_bytes<1>(h->cfg.height * GW_W));
  if (e == cudaSuccess && gwa::smem_bytes<2>(h->cfg.height * GW_W) <= (size_t)227 * 1024)
    e = cudaFuncSetAttribute(gwa::actor_forward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)gwa::smem_bytes<2>(h->cfg.height * GW_W));
  if (e != cudaSuccess) {
    if (a->d_params) cudaFree(a->d_params);
    delete a;
    return gw_cuda_fail(h, e, "gw_actor_create");
  }
  *out = a;
  return GW_OK;
}

int gw_actor_update(gw_actor* a, const gw_actor_weights* weights, int n_learners, void* stream) {
  if (!a || !weights) return GW_EINVAL;
  if (n_learners != a->n_learners) return gw_fail(a->h, GW_EINVAL, "gw_actor_update: learner count changed");
  std::vector<gwa::ActorParams> host;
  if (int rc = pack_weights(a->h, weights, n_learners, host)) return rc;
  GW_CUDA(a->h, cudaSetDevice(a->h->cfg.device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GW_CUDA(a->h, cudaMemcpyAsync(a->d_params, host.data(), sizeof(gwa::ActorParams) * n_learners, cudaMemcpyHostToDevice, s));
  GW_CUDA(a->h, cudaStreamSynchronize(s));            // `host` is a temporary
  return GW_OK;
}

int gw_actor_update_device(gw_actor* a, const gw_actor_weights* dev_weights, int n_learners, void* stream) {
  if (!a || !dev_weights) return GW_EINVAL;
  gw_handle* h = a->h;
  if (n_learners != a->n_learners || n_learners > GW_MAX_LEARNERS) return gw_fail(h, GW_EINVAL, "gw_actor_update_device: learner count changed");
  if (int rc = gw_server_stop(h)) return rc;
  gwa::PackArgs p;
  std::memset(&p, 0, sizeof(p));
  for (int k = 0; k < n_learners; ++k) {
    const gw_actor_weights& W = dev_weights[k];
    if (!W.w1 || !W.b1 || !W.ln1_g || !W.ln1_b || !W.w2 || !W.b2 || !W.ln2_g || !W.ln2_b || !W.w3 || !W.b3)
      return gw_fail(h, GW_EINVAL, "gw_actor_update_device: null weight pointer");
    p.w[k] = W;
  }
  p.out = a->d_params;
  for (int r = 0; r < GW_MAX_H; ++r) p.map_rows[r] = h->cfg.map_rows[r];
  p.cpo = h->cfg.height * GW_W;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  gwa::actor_pack_kernel<<<dim3(40, (unsigned)n_learners), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_actor_destroy(gw_actor* a) {
  if (!a) return GW_OK;
  cudaSetDevice(a->h->cfg.device);
  if (a->d_params) cudaFree(a->d_params);
  delete a;
  return GW_OK;
}

int gw_actor_forward(gw_actor* a, int64_t num_envs, const uint64_t* obs_code, const int8_t* action_mask,
                     float* cont_actions, int8_t* action_ids, int training, float expl_noise, float mean_noise,
                     uint64_t seed, uint64_t step, void* stream) {
  if (!a) return GW_EINVAL;
  gw_handle* h = a->h;
  if (int rc = gw_server_stop(h)) return rc;           // a resident step kernel would hold this stream
  if (num_envs < 0 || !obs_code || !cont_actions || !action_ids) return gw_fail(h, GW_EINVAL, "gw_actor_forward: null argument");
  if (num_envs == 0) return GW_OK;
  GW_CUDA(h, cudaSetDevice(h->cfg.device));
  gwa::FwdArgs f;
  std::memset(&f, 0, sizeof(f));
  f.params = a->d_params;
  f.obs_code = reinterpret_cast<const unsigned long long*>(obs_code);
  f.action_mask = action_mask;
  f.cont = cont_actions;
  f.ids = action_ids;
  f.E = num_envs;
  f.env_id_base = h->cfg.env_id_base;
  f.n = h->cfg.n_agents; f.nl = h->cfg.n_learners; f.kind = h->cfg.env_kind; f.cpo = h->cfg.height * GW_W;
  for (int k = 0; k < h->cfg.n_learners; ++k)
    if (h->cfg.apple_row[k] >= 0) f.apple_cells |= (uint32_t)((h->cfg.apple_row[k] << 4) | h->cfg.apple_col[k]) << (8 * k);
  if (training < 0 || training > 2) return gw_fail(h, GW_EINVAL, "gw_actor_forward: training must be 0, 1 or 2");
  f.gumbel = training != 0; f.explore = training == 1; f.expl_noise = expl_noise; f.mean_noise = mean_noise;
  f.seed_lo = (uint32_t)seed; f.seed_hi = (uint32_t)(seed >> 32);
  f.step_lo = (uint32_t)step; f.step_hi = (uint32_t)(step >> 32);
  const int nl = h->cfg.n_learners > 0 ? h->cfg.n_learners : 1;
  // half-full tiles while full ones would leave SMs without a CTA (latency regime)
  // (half tiles only while ALL of them fit in one wave: two half tiles in sequence on one SM are slower than one full tile)
  f.rows_per_tile = ((num_envs + gwa::ROWS / 2 - 1) / (gwa::ROWS / 2)) * nl <= (long long)h->sm_count ? gwa::ROWS / 2 : gwa::ROWS;
  const long long tiles = (num_envs + f.rows_per_tile - 1) / f.rows_per_tile;
  const long long cap = (long long)h->sm_count / nl > 0 ? (long long)h->sm_count / nl : 1;   // one 512-thread CTA per SM over all learners
  // two tiles in flight per CTA as soon as some CTA would otherwise run two tiles one after the other (and the second
  // operand tile fits): 16 384 envs x 2 learners 14.5 -> 12.5 us
  static const int groups_env = [] { const char* v = std::getenv("GW_ACTOR_GROUPS"); return v ? std::atoi(v) : 0; }();
  const bool two = groups_env ? groups_env == 2
                              : (tiles > cap && gwa::smem_bytes<2>(f.cpo) <= (size_t)227 * 1024);
  if (two && gwa::smem_bytes<2>(f.cpo) <= (size_t)227 * 1024) {
    const long long pairs = (tiles + 1) / 2;
    dim3 grid((unsigned)(pairs < cap ? pairs : cap), (unsigned)h->cfg.n_learners);
    gwa::actor_forward_kernel<2><<<grid, gwa::threads_of(2), gwa::smem_bytes<2>(f.cpo), static_cast<cudaStream_t>(stream)>>>(f);
  } else {
    dim3 grid((unsigned)(tiles < cap ? tiles : cap), (unsigned)h->cfg.n_learners);
    gwa::actor_forward_kernel<1><<<grid, gwa::threads_of(1), gwa::smem_bytes<1>(f.cpo), static_cast<cudaStream_t>(stream)>>>(f);
  }
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

}  // extern "C"
