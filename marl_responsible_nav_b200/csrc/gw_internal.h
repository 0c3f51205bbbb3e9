// Internal declarations shared by the translation units of libgridworld_b200.so (not part of the C-ABI).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "gw_device.cuh"

// Resident step server (gw_step_host mode 2, gw_kernels.cu): mailboxes and the registered buffer sets.
constexpr int GW_SRV_SETS = 1024;
struct gw_server {
  bool allocated = false, running = false, disabled = false;   // disabled: launches block here (profiler), see server_step
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  gw_io* h_stage = nullptr;                  // pinned staging of one table entry (same allocation as the doorbell)
  unsigned long long* h_bell = nullptr;      // pinned host (one allocation): doorbell word ...
  unsigned int* h_resp = nullptr;            // ... and, one cache line further, [0] completed seq [1] generation that has left
  unsigned long long* d_bell = nullptr;      // device (one allocation): relay word, arrival counter
  unsigned int* d_arrive = nullptr;
  gw_io* d_io_table = nullptr;               // [GW_SRV_SETS]
  std::vector<gw_io> sets;                   // host mirror of the table
  int last_set = -1, blocked_starts = 0;
  unsigned int seq = 0, generation = 0;
  unsigned long long idle_ns = 1000000ull;
  uint64_t launches = 0, relaunches = 0;
};

struct gw_host_call {                        // arguments of one gw_step_host call, prepared once (gw_host_call_prepare)
  gw_io io;
  const int8_t* actions; float* reward; float* shaped; uint8_t* ended;
  int mode;
};

struct gw_handle {
  gw_config cfg;
  gw_server srv;
  std::vector<gw_host_call> host_calls;
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

// GeneratePolicy -> 31-bit cdf thresholds of the device sampler (gw_kernels.cu), shared with the general layout (gw_wide.cu)
extern "C" void gw_policy_thresholds(const float sw[3], const float dw_in[4], bool perturbed, uint32_t thr[8]);
