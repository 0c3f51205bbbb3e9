// Dev microbenchmark: what a write-only stream reaches on this GPU (the step kernel is store-dominated).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_bw scripts/micro/store_bw.cu && /tmp/store_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_stg(uint4* dst, size_t n, int cs) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (cs) __stcs(dst + i, v); else dst[i] = v;
  }
}

// every warp owns a 1280-byte row in shared memory and pushes it out with one bulk (TMA) store per "env"
__global__ void k_bulk(uint8_t* dst, size_t n_rows, int row_bytes) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint8_t* row = sm + (size_t)warp * row_bytes;
  for (int i = lane; i < row_bytes / 16; i += 32) reinterpret_cast<uint4*>(row)[i] = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const uint32_t srow = (uint32_t)__cvta_generic_to_shared(row);
  for (size_t r = (size_t)blockIdx.x * nw + warp; r < n_rows; r += (size_t)gridDim.x * nw) {
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + r * row_bytes), "r"(srow), "r"(row_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the step kernel's pattern: a CTA owns tiles of `tile` consecutive rows, warp w writes rows w, w+8, ... of the tile
__global__ void k_bulk_tiled(uint8_t* dst, size_t n_rows, int row_bytes, int tile) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint8_t* row = sm + (size_t)warp * row_bytes;
  for (int i = lane; i < row_bytes / 16; i += 32) reinterpret_cast<uint4*>(row)[i] = make_uint4(1, 2, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const uint32_t srow = (uint32_t)__cvta_generic_to_shared(row);
  const size_t n_tiles = n_rows / tile;
  for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    for (int r = warp; r < tile; r += nw) {
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (t * tile + r) * row_bytes), "r"(srow), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
      __syncwarp();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void k_stg_tiled(uint8_t* dst, size_t n_rows, int row_bytes, int tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const size_t n_tiles = n_rows / tile;
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x)
    for (int r = warp; r < tile; r += nw) {
      uint4* d = reinterpret_cast<uint4*>(dst + (t * tile + r) * row_bytes) + lane;
      for (int i = 0; i < row_bytes / 16; i += 32)
        if (i + lane < row_bytes / 16) __stcs(d + i, v);
    }
}

int main() {
  const size_t bytes = (size_t)1 << 31;   // 2 GiB > L2
  uint8_t* d;
  cudaMalloc(&d, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto timeit = [&](const char* name, auto launch) {
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-40s %8.1f GB/s  (%s)\n", name, bytes * 10.0 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  for (int mult : {4, 8, 16}) {
    char nm[64];
    snprintf(nm, 64, "st.global v4, %d CTAs/SM x 256", mult);
    timeit(nm, [&] { k_stg<<<sms * mult, 256>>>((uint4*)d, bytes / 16, 0); });
    snprintf(nm, 64, "st.global.cs v4, %d CTAs/SM x 256", mult);
    timeit(nm, [&] { k_stg<<<sms * mult, 256>>>((uint4*)d, bytes / 16, 1); });
  }
  for (int rb : {1280, 2560, 5120}) {
    for (int mult : {2, 4}) {
      char nm[64];
      snprintf(nm, 64, "bulk store %d B rows, %d CTAs/SM x 256", rb, mult);
      cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * rb);
      timeit(nm, [&] { k_bulk<<<sms * mult, 256, 8 * rb>>>(d, bytes / rb, rb); });
    }
  }
  for (int tile : {32, 256, 1024}) {
    char nm[64];
    snprintf(nm, 64, "tiled bulk 1280 B, tile %d, 4 CTAs/SM", tile);
    cudaFuncSetAttribute(k_bulk_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1280);
    timeit(nm, [&] { k_bulk_tiled<<<sms * 4, 256, 8 * 1280>>>(d, bytes / 1280 / 1024 * 1024, 1280, tile); });
    snprintf(nm, 64, "tiled st.cs 1280 B, tile %d, 4 CTAs/SM", tile);
    timeit(nm, [&] { k_stg_tiled<<<sms * 4, 256>>>(d, bytes / 1280 / 1024 * 1024, 1280, tile); });
  }
  timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(d, 1, bytes); });
  return 0;
}
