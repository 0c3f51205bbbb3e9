// Throughput of the legacy (mma.sync) tensor path on this GPU: cycles per MMA instruction per SM with 8 warps per SM and four
// independent accumulators per warp.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(float* out, long long* cyc, int iters) {
  float c[4][4] = {};
  unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 11, b1 = 13;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[s][0]), "+f"(c[s][1]), "+f"(c[s][2]), "+f"(c[s][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[s][0]), "+f"(c[s][1]), "+f"(c[s][2]), "+f"(c[s][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 2)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[s][0]), "+f"(c[s][1]), "+f"(c[s][2]), "+f"(c[s][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else {                                                // 32 independent FMAs per "instruction slot" for comparison
#pragma unroll
        for (int j = 0; j < 4; ++j) c[s][j] = fmaf(c[s][j], 1.0001f, 0.5f);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][2] + c[3][3];
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  const char* names[] = {"tf32 m16n8k8", "bf16 m16n8k16", "f16 m16n8k16", "fp32 fma x4 (per thread)"};
  for (int kind = 0; kind < 4; ++kind) for (int warps = 1; warps <= 8; warps *= 2) {
    const int iters = 4096;
    long long h[148];
    for (int rep = 0; rep < 2; ++rep) {
      if (kind == 0) k<0><<<148, 32 * warps>>>(out, cyc, iters); else if (kind == 1) k<1><<<148, 32 * warps>>>(out, cyc, iters);
      else if (kind == 2) k<2><<<148, 32 * warps>>>(out, cyc, iters); else k<3><<<148, 32 * warps>>>(out, cyc, iters);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / (iters * 4.0 * warps);
    printf("%-26s warps/SM %d: %.2f cycles per instruction per SM (%.0f flop/clk/SM)\n", names[kind], warps, per,
           (kind == 0 ? 2048.0 : kind == 3 ? 4 * 32 * 2.0 : 4096.0) / per);
  }
  return 0;
}
