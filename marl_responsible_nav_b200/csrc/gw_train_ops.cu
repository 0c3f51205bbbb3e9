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

// ---- Linear backward: dW = dy^T x, db = column sums of dy, dx = dy W in ONE launch (PyTorch: two cuBLAS GEMMs and a
// reduction kernel, three dependent launches).  At batch 128 and widths <= 338 each product is ~10 MFLOP -- a fraction
// of a microsecond of fp32 FMA throughput -- so a plain shared-memory tiling is enough: 32x32 output tiles, K in chunks
// of 32, 256 threads with 2x2 outputs each.  The first n_g1 CTAs own tiles of dW (those of tile column 0 also sum dy's
// columns for db), the others tiles of dx.
struct LinBwdArgs {
  const float* dy; const float* x; const float* w;
  float* dw; float* db; float* dx;
  int batch, in, out, ldx;
  int tiles_n, n_g1;
};

__global__ void __launch_bounds__(256) linear_bwd_kernel(const LinBwdArgs a) {
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  int t = blockIdx.x;
  const bool g1 = t < a.n_g1;
  if (!g1) t -= a.n_g1;
  const int tm = t / a.tiles_n, tn = t - tm * a.tiles_n;
  const int M = g1 ? a.out : a.batch, N = a.in, K = g1 ? a.batch : a.out;
  const float* __restrict__ bsrc = g1 ? a.x : a.w;
  const int ldb = g1 ? a.ldx : a.in;
  const int m0 = tm * 32, n0 = tn * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc00 = 0.f, acc01 = 0.f, acc10 = 0.f, acc11 = 0.f, colsum = 0.f;
  const bool sums = g1 && tn == 0 && threadIdx.x < 32;
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int i = threadIdx.x; i < 1024; i += 256) {
      const int hi = i >> 5, lo = i & 31;
      float v;
      if (g1) {                                             // A(k, m) = dy[k][m]: rows of dy are contiguous in m
        v = (k0 + hi < K && m0 + lo < M) ? a.dy[(long long)(k0 + hi) * a.out + m0 + lo] : 0.f;
        As[hi][lo] = v;
      } else {                                              // A(k, m) = dy[m][k]: contiguous in k, stored transposed
        v = (k0 + lo < K && m0 + hi < M) ? a.dy[(long long)(m0 + hi) * a.out + k0 + lo] : 0.f;
        As[lo][hi] = v;
      }
      Bs[hi][lo] = (k0 + hi < K && n0 + lo < N) ? bsrc[(long long)(k0 + hi) * ldb + n0 + lo] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[k][ty], a1 = As[k][ty + 16], b0 = Bs[k][tx], b1 = Bs[k][tx + 16];
      acc00 = fmaf(a0, b0, acc00); acc01 = fmaf(a0, b1, acc01);
      acc10 = fmaf(a1, b0, acc10); acc11 = fmaf(a1, b1, acc11);
    }
    if (sums) {
#pragma unroll
      for (int k = 0; k < 32; ++k) colsum += As[k][threadIdx.x];
    }
    __syncthreads();
  }
  float* __restrict__ c = g1 ? a.dw : a.dx;
  const int m_a = m0 + ty, m_b = m0 + ty + 16, n_a = n0 + tx, n_b = n0 + tx + 16;
  if (m_a < M) {
    if (n_a < N) c[(long long)m_a * N + n_a] = acc00;
    if (n_b < N) c[(long long)m_a * N + n_b] = acc01;
  }
  if (m_b < M) {
    if (n_a < N) c[(long long)m_b * N + n_a] = acc10;
    if (n_b < N) c[(long long)m_b * N + n_b] = acc11;
  }
  if (sums && m0 + (int)threadIdx.x < M) a.db[m0 + threadIdx.x] = colsum;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int gw_ln_relu_forward(gw_handle* h, int64_t rows, int32_t width, const float* x, const float* gamma,
                                  const float* beta, float eps, float* y, float* mean, float* rstd, void* stream) {
  if (h == nullptr) return GW_EINVAL;
  if (int rc = gw_server_stop(h)) return rc;           // a resident step kernel would hold this stream
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
  if (int rc = gw_server_stop(h)) return rc;           // a resident step kernel would hold this stream
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

extern "C" int gw_linear_backward(gw_handle* h, int32_t batch, int32_t in_features, int32_t out_features, const float* dy,
                                  const float* x, int32_t x_row_stride, const float* w, float* dw, float* db, float* dx,
                                  void* stream) {
  if (h == nullptr) return GW_EINVAL;
  if (int rc = gw_server_stop(h)) return rc;           // a resident step kernel would hold this stream
  if (batch < 1 || in_features < 1 || out_features < 1 || x_row_stride < in_features || !dy || !x || !w || !dw || !db)
    return gw_fail(h, GW_EINVAL, "gw_linear_backward: bad argument");
  LinBwdArgs a;
  a.dy = dy; a.x = x; a.w = w; a.dw = dw; a.db = db; a.dx = dx;
  a.batch = batch; a.in = in_features; a.out = out_features; a.ldx = x_row_stride;
  a.tiles_n = (in_features + 31) / 32;
  a.n_g1 = ((out_features + 31) / 32) * a.tiles_n;
  const int n_g2 = dx != nullptr ? ((batch + 31) / 32) * a.tiles_n : 0;
  linear_bwd_kernel<<<(unsigned)(a.n_g1 + n_g2), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}
