// Device-side pieces of the replay sampler shared by gw_replay.cu (gw_replay_sample) and gw_maddpg.cu (the fused update
// draws and gathers its own batches): the (time, env) draw and the widening row copy.
#pragma once
#include <cuda_bf16.h>

#include "gw_internal.h"

namespace gwr {

// Philox(seed; sample index, draw number) -> (absolute time step, env) uniform over the newest n_valid steps x num_envs
__device__ __forceinline__ void draw_index(long long b, unsigned long long draw, uint32_t k0, uint32_t k1, long long t_now,
                                           long long n_valid, long long num_envs, long long& t_abs, long long& e) {
  uint32_t c[4] = {(uint32_t)b, (uint32_t)(b >> 32), (uint32_t)draw, (uint32_t)(draw >> 32)};
  gw::philox4x32(c, k0, k1);
  // 64-bit words scaled to the range (multiply-high): bias < range / 2^64
  const unsigned long long u0 = ((unsigned long long)c[0] << 32) | c[1], u1 = ((unsigned long long)c[2] << 32) | c[3];
  const long long k = (long long)__umul64hi(u0, (unsigned long long)n_valid);
  e = (long long)__umul64hi(u1, (unsigned long long)num_envs);
  t_abs = t_now - 1 - k;                                 // the newest n_valid time steps are stored
}

__host__ __device__ __forceinline__ void sample_key(uint64_t seed, uint32_t& k0, uint32_t& k1) {
  k0 = (uint32_t)seed;
  k1 = (uint32_t)(seed >> 32) ^ 0x52455053u;             // "REPS": own key space next to the env's streams
}

__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// n elements of one observation row -> f32
template <typename T>
__device__ __forceinline__ void copy_row(const T* __restrict__ src, float* __restrict__ dst, int n, int lane, bool vec);

template <>
__device__ __forceinline__ void copy_row<float>(const float* __restrict__ src, float* __restrict__ dst, int n, int lane,
                                                bool vec) {
  if (vec) {
    for (int i = lane * 4; i < n; i += 128) *reinterpret_cast<float4*>(dst + i) = load4(src + i);
  } else {
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
  }
}

template <>
__device__ __forceinline__ void copy_row<__nv_bfloat16>(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst,
                                                        int n, int lane, bool vec) {
  if (vec) {
    for (int i = lane * 8; i < n; i += 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {                      // bf16 -> f32 is a 16-bit shift
        o[2 * j] = __uint_as_float(w[j] << 16);
        o[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
      }
      *reinterpret_cast<float4*>(dst + i) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + i + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  } else {
    for (int i = lane; i < n; i += 32) dst[i] = __bfloat162float(src[i]);
  }
}

}  // namespace gwr
