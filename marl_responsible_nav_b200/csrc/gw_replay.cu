// Replay sampling (K4, read side): one kernel draws a batch of (time, env) indices and gathers the transitions from the
// device-resident ring straight into the trainer's batch tensors.
//
// Replaces the sample call of the reference's trainer (maddpg/agent.py:209-211: `memory.sample(BATCH_SIZE)` on AgileRL's
// MultiAgentReplayBuffer, fields state / action / reward / next_state / done, maddpg/agent.py:70).  The ring stores every
// observation once (gw_step writes observation t+1 into slot (t+1) % T), so `next_state` is the neighbouring slot, or
// the terminal observation in `final_obs` where the episode ended and the env was re-spawned inside the step.
//
// One warp per sample: lane 0 draws the indices (Philox4x32-10 keyed by the seed, counter = sample index and draw
// number), the warp copies the two 1 280-byte observation rows with 128-bit loads / stores and the short fields with
// one lane each.  A batch of 128 moves 0.35 MB: the kernel is launch-bound (a few microseconds) and replaces the ~25
// indexing kernels of the PyTorch formulation.
#include "gw_replay_dev.cuh"

namespace {

struct SampleArgs {
  gw_replay_view ring;
  long long t_now, n_valid;
  long long batch;
  uint32_t k0, k1;
  unsigned long long draw;
  const long long* t_in;
  const long long* env_in;
  float* state; float* action; float* reward; float* next_state; float* done;
  long long* t_out; long long* env_out;
};

template <typename T>
__global__ void __launch_bounds__(128) gw_replay_sample_kernel(const SampleArgs a, const bool vec) {
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= a.batch) return;
  const gw_replay_view& r = a.ring;
  long long t_abs = 0, e = 0;
  if (lane == 0) {
    if (a.t_in != nullptr) {
      t_abs = a.t_in[b];
      e = a.env_in[b];
    } else {
      gwr::draw_index(b, a.draw, a.k0, a.k1, a.t_now, a.n_valid, r.num_envs, t_abs, e);
    }
  }
  t_abs = __shfl_sync(0xffffffffu, t_abs, 0);
  e = __shfl_sync(0xffffffffu, e, 0);
  // caller-supplied indices are not trusted with the address arithmetic: wrap the time, clamp the env
  long long s = t_abs % r.slots;
  if (s < 0) s += r.slots;
  const long long s1 = s + 1 == r.slots ? 0 : s + 1;
  e = e < 0 ? 0 : (e >= r.num_envs ? r.num_envs - 1 : e);
  const long long row = s * r.num_envs + e, row1 = s1 * r.num_envs + e;
  const int L = r.n_learners, n_obs = L * r.obs_len, n_act = L * r.action_dim;
  const bool ended = r.ended[row] != 0;
  const T* obs = static_cast<const T*>(r.obs);
  const T* fin = static_cast<const T*>(r.final_obs);
  gwr::copy_row<T>(obs + row * n_obs, a.state + b * n_obs, n_obs, lane, vec);
  gwr::copy_row<T>(ended ? fin + row * n_obs : obs + row1 * n_obs, a.next_state + b * n_obs, n_obs, lane, vec);
  for (int i = lane; i < n_act; i += 32) a.action[b * n_act + i] = r.action[row * n_act + i];
  if (lane < L) {
    a.reward[b * L + lane] = r.reward[row * L + lane];
    a.done[b * L + lane] = r.terminated[row * L + lane] ? 1.0f : 0.0f;
  }
  if (lane == 0) {
    if (a.t_out) a.t_out[b] = t_abs;
    if (a.env_out) a.env_out[b] = e;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int gw_replay_sample(gw_handle* h, const gw_replay_view* ring, int64_t t_now, int64_t batch, uint64_t seed,
                                uint64_t draw, const int64_t* t_in, const int64_t* env_in, float* state, float* action,
                                float* reward, float* next_state, float* done, int64_t* t_out, int64_t* env_out,
                                void* stream) {
  if (h == nullptr) return GW_EINVAL;
  if (int rc = gw_server_stop(h)) return rc;           // a resident step kernel would hold this stream
  if (ring == nullptr || ring->struct_size != sizeof(gw_replay_view))
    return gw_fail(h, GW_EINVAL, "gw_replay_sample: ring view missing or of another size");
  const gw_replay_view& r = *ring;
  if (r.slots < 3 || r.num_envs < 1 || r.n_learners < 1 || r.n_learners > 32 || r.obs_len < 1 || r.action_dim < 1)
    return gw_fail(h, GW_EINVAL, "gw_replay_sample: bad ring shape");
  if (r.obs_dtype != GW_OBS_F32 && r.obs_dtype != GW_OBS_BF16) return gw_fail(h, GW_EINVAL, "gw_replay_sample: bad obs_dtype");
  if (!r.obs || !r.final_obs || !r.action || !r.reward || !r.terminated || !r.ended)
    return gw_fail(h, GW_EINVAL, "gw_replay_sample: ring pointer missing");
  if (!state || !action || !reward || !next_state || !done)
    return gw_fail(h, GW_EINVAL, "gw_replay_sample: output pointer missing");
  if ((t_in == nullptr) != (env_in == nullptr))
    return gw_fail(h, GW_EINVAL, "gw_replay_sample: t_in and env_in go together");
  if (batch < 0) return gw_fail(h, GW_EINVAL, "gw_replay_sample: negative batch");
  const int64_t n_valid = t_now < r.slots - 1 ? t_now : r.slots - 1;   // slot t+1 holds the next observation
  if (n_valid < 1) return gw_fail(h, GW_ESTATE, "gw_replay_sample: the ring is empty");
  if (batch == 0) return GW_OK;
  SampleArgs a;
  a.ring = r;
  a.t_now = t_now; a.n_valid = n_valid; a.batch = batch;
  gwr::sample_key(seed, a.k0, a.k1);
  a.draw = draw;
  a.t_in = reinterpret_cast<const long long*>(t_in); a.env_in = reinterpret_cast<const long long*>(env_in);
  a.state = state; a.action = action; a.reward = reward; a.next_state = next_state; a.done = done;
  a.t_out = reinterpret_cast<long long*>(t_out); a.env_out = reinterpret_cast<long long*>(env_out);
  const int n_obs = r.n_learners * r.obs_len;
  const int per_vec = r.obs_dtype == GW_OBS_F32 ? 4 : 8;
  const bool vec = n_obs % per_vec == 0 && aligned16(r.obs) && aligned16(r.final_obs) && aligned16(state) && aligned16(next_state);
  const unsigned grid = (unsigned)((batch + 3) / 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (r.obs_dtype == GW_OBS_F32) gw_replay_sample_kernel<float><<<grid, 128, 0, st>>>(a, vec);
  else gw_replay_sample_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(a, vec);
  GW_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}
