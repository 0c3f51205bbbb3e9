// Internal declarations shared by the translation units of libgridworld_b200.so (not part of the C-ABI).
#pragma once
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

#include "gw_device.cuh"

struct gw_handle {
  gw_config cfg;
  gw::Tables* d_tables = nullptr;
  uint4* d_rng_cache = nullptr;              // small-batch kernel: next step's random words per env (tagged)
  unsigned int* d_tile_ctr = nullptr;        // dynamic tile scheduling of the thread-per-env step kernel
  uint8_t* d_stage_init = nullptr;           // replicated observation rows for the kernels' prologue copy
  uint4* d_state = nullptr;
  unsigned long long* d_stats = nullptr;
  unsigned long long* d_trace = nullptr;     // GW_TRACE (dev)
  bool reset_done = false;
  int sm_count = 148;
  int n_active = 0;
  uint64_t launches = 0;
  uint64_t env_steps = 0;
  std::string err;
};

int gw_fail(gw_handle* h, int code, const std::string& msg);
int gw_cuda_fail(gw_handle* h, cudaError_t e, const char* what);
#define GW_CUDA(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return gw_cuda_fail(h, e_, #call); } while (0)
