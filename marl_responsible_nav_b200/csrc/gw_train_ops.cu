// Trainer-side element-wise kernels of the MADDPG update: LayerNorm + ReLU of the reference's networks
// (Linear -> LayerNorm -> ReLU twice; AgileRL builds them from configs/mlp.yaml, hidden 128) as ONE kernel
// forward and ONE kernel backward.  The update at batch 128 is a chain of ~100 dependent kernels of launch-latency size
// (Trainer), so what counts is the number of kernels on the chain, not their arithmetic: PyTorch runs
// LayerNorm and ReLU as two kernels forward and three to four backward (dX, dgamma/dbeta in two stages, ReLU mask).
//
// Width is fixed at 128 columns = one float4 per lane, one warp per row; statistics by warp shuffles in fp32
// (two-pass: mean, then centred second moment; biased variance, rstd = rsqrt(var + eps) as torch.nn.LayerNorm).
#include "gw_internal.h"

namespace {

constexpr int LN_W = 128;
constexpr int BWD_THREADS = 1024;              // 32 warps x 4 rows = 128 rows per CTA: batch 128 is ONE CTA (no atomics)
constexpr int BWD_ROWS = 128;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) ln_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps, float* __restrict__ y,
                                                          float* __restrict__ mean, float* __restrict__ rstd, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4 v = reinterpret_cast<const float4*>(x + row * LN_W)[lane];
  const float4 g = reinterpret_cast<const float4*>(gamma)[lane];
  const float4 b = reinterpret_cast<const float4*>(beta)[lane];
  const float mu = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.0f / LN_W);
  const float d0 = v.x - mu, d1 = v.y - mu, d2 = v.z - mu, d3 = v.w - mu;
  const float var = warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) * (1.0f / LN_W);
  const float rs = rsqrtf(var + eps);
  float4 o;
  o.x = fmaxf(fmaf(d0 * rs, g.x, b.x), 0.f);
  o.y = fmaxf(fmaf(d1 * rs, g.y, b.y), 0.f);
  o.z = fmaxf(fmaf(d2 * rs, g.z, b.z), 0.f);
  o.w = fmaxf(fmaf(d3 * rs, g.w, b.w), 0.f);
  reinterpret_cast<float4*>(y + row * LN_W)[lane] = o;
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dy -> (dx, dgamma, dbeta).  pre = xhat * gamma + beta; ReLU mask = pre > 0 (recomputed, nothing but x / mean / rstd saved).
// dgamma / dbeta: per-lane partial sums over the warp's rows, summed over the CTA's warps in shared memory; one CTA writes
// them, several CTAs add them atomically to zero-filled outputs.
__global__ void __launch_bounds__(BWD_THREADS) ln_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  float* __restrict__ dx, float* __restrict__ dgamma,
                                                                  float* __restrict__ dbeta, long long rows) {
  __shared__ float4 s_dg[BWD_THREADS / 32][32];
  __shared__ float4 s_db[BWD_THREADS / 32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 g = reinterpret_cast<const float4*>(gamma)[lane];
  const float4 b = reinterpret_cast<const float4*>(beta)[lane];
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
  const long long row0 = (long long)blockIdx.x * BWD_ROWS;
  for (int i = warp; i < BWD_ROWS; i += BWD_THREADS / 32) {
    const long long row = row0 + i;
    if (row >= rows) break;
    const float4 v = reinterpret_cast<const float4*>(x + row * LN_W)[lane];
    const float4 u = reinterpret_cast<const float4*>(dy + row * LN_W)[lane];
    const float mu = mean[row], rs = rstd[row];
    const float h0 = (v.x - mu) * rs, h1 = (v.y - mu) * rs, h2 = (v.z - mu) * rs, h3 = (v.w - mu) * rs;
    const float p0 = fmaf(h0, g.x, b.x) > 0.f ? u.x : 0.f, p1 = fmaf(h1, g.y, b.y) > 0.f ? u.y : 0.f;
    const float p2 = fmaf(h2, g.z, b.z) > 0.f ? u.z : 0.f, p3 = fmaf(h3, g.w, b.w) > 0.f ? u.w : 0.f;
    ag.x = fmaf(p0, h0, ag.x); ag.y = fmaf(p1, h1, ag.y); ag.z = fmaf(p2, h2, ag.z); ag.w = fmaf(p3, h3, ag.w);
    ab.x += p0; ab.y += p1; ab.z += p2; ab.w += p3;
    const float e0 = p0 * g.x, e1 = p1 * g.y, e2 = p2 * g.z, e3 = p3 * g.w;          // d loss / d xhat
    const float m1 = warp_sum((e0 + e1) + (e2 + e3)) * (1.0f / LN_W);
    const float m2 = warp_sum((e0 * h0 + e1 * h1) + (e2 * h2 + e3 * h3)) * (1.0f / LN_W);
    float4 o;
    o.x = rs * (e0 - m1 - h0 * m2);
    o.y = rs * (e1 - m1 - h1 * m2);
    o.z = rs * (e2 - m1 - h2 * m2);
    o.w = rs * (e3 - m1 - h3 * m2);
    reinterpret_cast<float4*>(dx + row * LN_W)[lane] = o;
  }
  s_dg[warp][lane] = ag;
  s_db[warp][lane] = ab;
  __syncthreads();
  if (threadIdx.x < 2 * LN_W) {                              // 128 threads sum dgamma's columns, 128 dbeta's
    const int which = threadIdx.x >= LN_W, c = threadIdx.x & (LN_W - 1);
    const float* src = reinterpret_cast<const float*>(which ? &s_db[0][0] : &s_dg[0][0]);
    float acc = 0.f;
#pragma unroll 8
    for (int w = 0; w < BWD_THREADS / 32; ++w) acc += src[w * LN_W + c];
    float* dst = which ? dbeta : dgamma;
    if (gridDim.x == 1) dst[c] = acc;
    else atomicAdd(dst + c, acc);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int gw_ln_relu_forward(gw_handle* h, int64_t rows, int32_t width, const float* x, const float* gamma,
                                  const float* beta, float eps, float* y, float* mean, float* rstd, void* stream) {
  if (h == nullptr) return GW_EINVAL;
  if (width != LN_W) return gw_fail(h, GW_EINVAL, "gw_ln_relu_forward: width must be 128");
  if (rows < 0 || !x || !gamma || !beta || !y || !mean || !rstd) return gw_fail(h, GW_EINVAL, "gw_ln_relu_forward: bad argument");
  if (!aligned16(x) || !aligned16(gamma) || !aligned16(beta) || !aligned16(y))
    return gw_fail(h, GW_EINVAL, "gw_ln_relu_forward: pointers must be 16-byte aligned");
  if (rows == 0) return GW_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  ln_relu_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, eps, y, mean, rstd, rows);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

extern "C" int gw_ln_relu_backward(gw_handle* h, int64_t rows, int32_t width, const float* dy, const float* x,
                                   const float* mean, const float* rstd, const float* gamma, const float* beta, float* dx,
                                   float* dgamma, float* dbeta, void* stream) {
  if (h == nullptr) return GW_EINVAL;
  if (width != LN_W) return gw_fail(h, GW_EINVAL, "gw_ln_relu_backward: width must be 128");
  if (rows < 1 || !dy || !x || !mean || !rstd || !gamma || !beta || !dx || !dgamma || !dbeta)
    return gw_fail(h, GW_EINVAL, "gw_ln_relu_backward: bad argument");
  if (!aligned16(dy) || !aligned16(x) || !aligned16(gamma) || !aligned16(beta) || !aligned16(dx))
    return gw_fail(h, GW_EINVAL, "gw_ln_relu_backward: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + BWD_ROWS - 1) / BWD_ROWS);
  if (grid > 1) {                                            // several CTAs add their column sums atomically
    GW_CUDA(h, cudaMemsetAsync(dgamma, 0, LN_W * sizeof(float), st));
    GW_CUDA(h, cudaMemsetAsync(dbeta, 0, LN_W * sizeof(float), st));
  }
  ln_relu_bwd_kernel<<<grid, BWD_THREADS, 0, st>>>(dy, x, mean, rstd, gamma, beta, dx, dgamma, dbeta, rows);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}
