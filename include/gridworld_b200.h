/*
 * gridworld_b200.h -- C-ABI of the B200-native batched grid-world step / FeAR /
 * observation-render path (libgridworld_b200.so).
 *
 * This is the drop-in boundary for the hot path of Henweiz/MARL-Responsible-Nav
 * (reference files are cited as <file>:<line>, relative to the reference root):
 *
 *   gw_reset          <- CustomMAEnv.reset / setup_env     custom/ma_customenv.py:169-215, :338-429
 *                        CustomEnv.reset                   custom/customenv.py:186-356
 *   gw_step           <- CustomMAEnv.step  / setup_step    custom/ma_customenv.py:217-334, :432-452
 *                        CustomEnv.step                    custom/customenv.py:78-183
 *                        (GWorld.UpdateGWorld              custom/grid_world.py:424-563,
 *                         Responsibility.FeAR_4_one_actor  custom/Responsibility.py:135-210,
 *                         get_action_mask                  custom/ma_customenv.py:467-506,
 *                         obs flatten + reward shaping     maddpg/agent.py:89,128-131,160)
 *   gw_update_world   <- GWorld.UpdateGWorld (operator level, arbitrary positions/actions)
 *   gw_fear_one_actor <- Responsibility.FeAR_4_one_actor (operator level, arbitrary close list)
 *   gww_*             <- the same calls for maps / agent counts beyond the shipped scenarios (GWorld is generic:
 *                        custom/grid_world.py:17-86, :104-150; Scenario['N_Agents'], custom/ma_customenv.py:28)
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every entry returns gw_status
 *     (0 = OK) and gw_last_error() gives the message;
 *   - all array arguments are DEVICE pointers owned by the caller (PyTorch
 *     tensors: tensor.data_ptr()); the handle owns only the packed per-env
 *     state, its lookup tables and the statistics accumulators;
 *   - calls are asynchronous and ordered on the caller's CUDA stream
 *     (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream);
 *     nothing in gw_step / gw_reset synchronises or allocates, so both are
 *     CUDA-graph capturable;
 *   - there is NO CPU implementation behind this interface: without a CUDA
 *     device gw_create fails with GW_ENODEV.
 *   - cells are (row, col) with row-major flat index row*W+col.  Two state layouts: gw_* (packed: W must be 16, H <= 16,
 *     <= 4 agents -- all shipped scenarios are 10x16 with 3..4 agents -- and every kernel of the library) and gww_* (general:
 *     up to 64 x 64, up to 16 agents; the section at the end of this file).
 */
#ifndef GRIDWORLD_B200_H
#define GRIDWORLD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GW_ABI_VERSION 1
#define GW_MAX_AGENTS 4
#define GW_MAX_LEARNERS 2
#define GW_N_ACTIONS 9        /* custom/custom_agent.py:140-150 */
#define GW_MAX_BLOCKED 256
#define GW_MAX_POLICIES 16
#define GW_MAX_H 16
#define GW_W 16

typedef enum gw_status {
  GW_OK = 0,
  GW_EINVAL = -1,   /* bad config / argument / alignment */
  GW_ENOMEM = -2,
  GW_ECUDA = -3,    /* CUDA runtime error (message has the cudaError string) */
  GW_ENODEV = -4,   /* no CUDA device: there is no CPU fallback */
  GW_ESTATE = -5    /* call order (step before reset) */
} gw_status;

typedef enum gw_env_kind {
  GW_ENV_MULTI = 0,   /* custom/ma_customenv.py: learners 0..n_learners-1, one apple each, int rewards */
  GW_ENV_SINGLE = 1   /* custom/customenv.py: learner 0, one apple, +0.1 shaping, raw-id observation */
} gw_env_kind;

typedef enum gw_obs_dtype { GW_OBS_F32 = 0, GW_OBS_BF16 = 1 } gw_obs_dtype;

/* Plain-old-data environment description (what the reference keeps in
 * Scenarios.json + module constants + configs/custom*.yaml). */
typedef struct gw_config {
  int32_t struct_size;                 /* = sizeof(gw_config) */
  int32_t abi_version;                 /* = GW_ABI_VERSION */
  int32_t height, width;               /* Scenario['Map']['Region'] shape */
  uint16_t map_rows[GW_MAX_H];         /* bit c of map_rows[r] = Region[r][c] == 1 */
  int32_t n_agents;                    /* Scenario['N_Agents'] (2..4) */
  int32_t n_learners;                  /* N_INTELLIGENT_AGENTS (1..2), ma_customenv.py:19 */
  int32_t env_kind;                    /* gw_env_kind */
  int8_t apple_row[GW_MAX_LEARNERS];   /* ma_customenv.py:422 / customenv.py:334; -1 = none */
  int8_t apple_col[GW_MAX_LEARNERS];
  uint8_t policy_map[GW_MAX_H * GW_W]; /* policy-region index per cell, ma_customenv.py:346-354 */
  uint8_t mdr_map[GW_MAX_H * GW_W];    /* MdR ACTION id per cell (mdrs[key]['mdr']), :357-365,:445-447 */
  int32_t n_policies;
  float step_weights[GW_MAX_POLICIES][3];   /* Policies[key]['stepWeights'] */
  float dir_weights[GW_MAX_POLICIES][4];    /* Policies[key]['directionWeights'] (Up,Down,Left,Right) */
  double perturb_prob;                 /* 0.25, ma_customenv.py:441 */
  int32_t fear;                        /* compute FeAR (WITH_FEAR / fear=True) */
  int32_t fear_radius;                 /* 5, ma_customenv.py:249 */
  double fear_weight;                  /* FeAR_weight, maddpg/agent.py:128-130 (shaped_reward output) */
  int32_t max_steps;                   /* TRAIN_STEPS episode cap used by auto-reset, 0 = none */
  int32_t auto_reset;                  /* re-spawn finished envs inside gw_step */
  int32_t obs_dtype;                   /* gw_obs_dtype */
  int32_t device;                      /* CUDA device ordinal */
  int64_t num_envs;                    /* E on this handle */
  int64_t env_id_base;                 /* global id of env 0 (sharding: results do not depend on the split) */
  uint64_t seed;                       /* device RNG key (Philox4x32-10) */
  /* Restricted paths (GWorld.RestrictedPaths, custom/grid_world.py:32-86): single-cell moves `from -> to` that are refused
   * (the agent stays, the move counts as restricted, :493-509).  A wall contributes both directions, a one-way the
   * direction against it.  Cells as (row << 4) | col.  The reference builds these from Scenario['Map']['Walls'/'OneWays'],
   * where JSON lists never compare equal to its tuple paths, so they are inert there; here they are enforced as the
   * commented-out tuple conversion of LoadJsonScenario (:654-665) intends.  0 = none (every shipped scenario). */
  int32_t n_blocked;
  uint8_t blocked_from[GW_MAX_BLOCKED];
  uint8_t blocked_to[GW_MAX_BLOCKED];
} gw_config;

/* Device pointers for one gw_reset / gw_step call.  NULL = not wanted / not given. */
typedef struct gw_io {
  /* inputs */
  const int8_t* learner_actions; /* [E, n_learners]   required by gw_step */
  const int8_t* npc_actions;     /* [E, n_agents]     replay mode: SelectActionsForAll's draw for every agent
                                                      (learner columns ignored); NULL = device RNG */
  const int8_t* spawn;           /* [E, n_agents, 2]  replay mode: (row, col) spawn cells, used whenever an env
                                                      (re)spawns in this call; NULL = device RNG */
  /* outputs */
  void* obs;                     /* [E, n_learners, H*W] f32 or bf16: next policy input (reset obs after an auto-reset) */
  void* final_obs;               /* [E, n_learners, H*W] written ONLY for envs whose episode ended in this call */
  float* reward;                 /* [E, n_learners] env reward */
  float* shaped_reward;          /* [E, n_learners] fear_weight*fear + reward, fp64 math rounded once */
  double* fear;                  /* [E, n_learners] info["fear"] */
  uint8_t* terminated;           /* [E, n_learners] */
  uint8_t* truncated;            /* [E, n_learners] */
  uint8_t* ended;                /* [E] episode boundary: all-truncated (multi) / terminated|truncated (single) / max_steps */
  int8_t* action_mask;           /* [E, n_learners, 9] */
  int8_t* positions;             /* [E, n_agents, 2] agent cells after the step (before any auto-reset) */
  uint32_t* info;                /* [E] bits 0-3 crash, 4-7 restricted, 8-9 learner crashes, 10-11 apples rewarded,
                                        12 ended, 13 unresolved collisions (reference prints a warning),
                                        14-15 distance-shaping reward fired for learner 0/1 */
  uint64_t* obs_code;            /* [E] compact form of `obs` for gw_actor_forward: bits 0-31 the four agent cells
                                        ((row<<4)|col, 8 bits each), 32-33 apples still shown, 34 fresh spawn (0.5 marker) */
} gw_io;

typedef struct gw_stats {          /* sums since gw_create / gw_reset_stats, this handle only */
  uint64_t env_steps;
  uint64_t agent_steps;            /* env_steps * n_learners */
  uint64_t episodes;
  uint64_t episode_len_sum;
  uint64_t crashes;                /* learner crashes */
  uint64_t apples;                 /* apples rewarded */
  uint64_t unresolved;
  uint64_t fear_nonzero;           /* learner-steps with fear != 0 */
  double return_sum;               /* sum of env rewards over finished episodes (all learners) */
  double fear_sum;
  uint64_t fear_tasks;             /* (actor, affected) pairs that needed counterfactual simulation (18 sims each) */
} gw_stats;

typedef struct gw_handle gw_handle;

int gw_abi_version(void);
const char* gw_build_info(void);              /* arch, nvcc version */
int gw_default_config(gw_config* cfg);        /* zero + struct_size/abi_version + reference constants */
int gw_create(const gw_config* cfg, gw_handle** out);
int gw_destroy(gw_handle* h);
const char* gw_last_error(const gw_handle* h); /* h may be NULL (errors from gw_create) */

/* reset envs with reset_mask[e] != 0 (device uint8 [E]); NULL = all */
int gw_reset(gw_handle* h, const uint8_t* reset_mask, const gw_io* io, void* stream);
int gw_step(gw_handle* h, const gw_io* io, void* stream);
int gw_sync(gw_handle* h, void* stream);       /* cudaStreamSynchronize + surface async faults */

/* Device-side rollout: `steps` consecutive gw_step's in ONE launch -- the inner loop of MADDPGAgent.train
 * (maddpg/agent.py:85-197: env.step, reward shaping, replay write) for a whole batch, when the actions of those steps
 * are already on the device (recorded / scripted action streams, open-loop evaluation, or the rollout of a policy
 * whose actions were computed ahead).  Environments are independent, so a CTA steps its 32 envs `steps` times without
 * any grid-wide synchronisation; packed state and random words stay in registers between the steps.
 * `rings` holds the pointers of slot 0 of TIME-MAJOR arrays (the replay ring's layout):
 *   outputs [ring_slots, <per-step shape of gw_io>]: the transition of step k (reward, shaped_reward, fear, terminated,
 *   truncated, ended, info, positions, final_obs) goes to slot (first_slot + k) % ring_slots, what the policy reads
 *   next (obs, obs_code, action_mask) to the slot after it, (first_slot + k + 1) % ring_slots;
 *   inputs [action_slots, ...] (learner_actions; npc_actions / spawn in replay mode): step k reads slot
 *   (first_action + k) % action_slots.
 * rings->obs and rings->learner_actions are required.  Results are bit-identical to `steps` gw_step calls.
 * Stream-ordered, no synchronisation, CUDA-graph capturable. */
typedef struct gw_rollout_plan {
  uint32_t struct_size;          /* sizeof(gw_rollout_plan) */
  int32_t steps;
  int64_t ring_slots;
  int64_t first_slot;
  int64_t action_slots;
  int64_t first_action;
} gw_rollout_plan;
int gw_rollout(gw_handle* h, const gw_io* rings, const gw_rollout_plan* plan, void* stream);

/* Host-driven step (the reference's calling pattern, maddpg/agent.py:121-131: actions arrive as host integers, rewards
 * and done flags are read on the host): copies `host_actions` [E, n_learners] int8 (pinned) to io->learner_actions,
 * runs gw_step, copies io->reward -> host_reward f32 [E, n_learners], io->shaped_reward -> host_shaped (nullable),
 * io->ended -> host_ended u8 [E] (nullable), and synchronises the stream.  One call, no Python between the stages.
 * zero_copy != 0: no memcpy nodes at all -- the kernel itself loads the actions from the pinned host buffer and stores
 * rewards / flags into the pinned host buffers over PCIe (unified addressing); same bytes, fewer stream operations.
 * zero_copy == GW_HOST_RESIDENT (2): as 1, but through a RESIDENT kernel (handles of up to 24 576 envs; larger ones fall
 * back to 1).  The first call launches it on `stream`; it keeps the tables in shared memory, polls a doorbell
 * in pinned host memory, steps, and reports completion through pinned host memory, so a step costs neither a launch nor
 * a stream synchronisation.  On return the host buffers hold the step's results and the device outputs (observations,
 * masks ...) are complete in HBM; work queued on `stream` runs once the resident kernel has left: at the next gw_* call
 * on this handle other than gw_step_host(mode 2) (they all stop it first), at gw_server_stop, or by itself after
 * GW_SERVER_IDLE_US (default 1000) microseconds without a step -- a later call relaunches it transparently. */
#define GW_HOST_MEMCPY 0
#define GW_HOST_ZERO_COPY 1
#define GW_HOST_RESIDENT 2
int gw_step_host(gw_handle* h, const gw_io* io, const int8_t* host_actions, float* host_reward, float* host_shaped,
                 uint8_t* host_ended, int zero_copy, void* stream);
/* The same call with its arguments stored in the handle: bindings whose per-argument marshalling costs as much as a PCIe
 * round trip (ctypes: ~1 us for the eight arguments) prepare each distinct set of buffers once and run it by token. */
int gw_host_call_prepare(gw_handle* h, const gw_io* io, const int8_t* host_actions, float* host_reward, float* host_shaped,
                         uint8_t* host_ended, int zero_copy, int* token);
int gw_host_call_run(gw_handle* h, int token, void* stream);
int gw_host_call_reset(gw_handle* h);             /* forget every prepared call */
int gw_server_stop(gw_handle* h);                 /* no-op when the resident kernel is not running */
/* running: 1 resident now, 0 not, -1 given up for this handle (kernel launches block in this process -- profiler,
 * CUDA_LAUNCH_BLOCKING -- so mode 2 runs as mode 1); any pointer may be NULL */
int gw_server_info(gw_handle* h, int* running, uint64_t* launches, uint64_t* relaunches, int* n_sets);

/* packed env state (16 bytes/env) for checkpoint / resume; device or host destination */
size_t gw_state_bytes(const gw_handle* h);
int gw_get_state(gw_handle* h, void* dst, int dst_is_device, void* stream);
int gw_set_state(gw_handle* h, const void* src, int src_is_device, void* stream);

int gw_get_stats(gw_handle* h, gw_stats* host_out, void* stream);   /* synchronises the stream */
int gw_reset_stats(gw_handle* h, void* stream);
int gw_launch_count(const gw_handle* h, uint64_t* kernels_launched); /* kernels launched by this handle so far */
/* development aid: with GW_TRACE=1 in the environment gw_create allocates a per-CTA stamp buffer (16 x uint64 per CTA) that the
 * resident step kernel fills; this copies it to the host (scripts/trace_server.py, scripts/trace_phases.py).  GW_EINVAL without it. */
int gw_debug_trace(gw_handle* h, unsigned long long* host_out, int max_ctas);

/* Operator level: C independent world updates (GWorld.UpdateGWorld with explicit actions).
 * n_agents_per_case NULL = cfg.n_agents everywhere.  apples: [C,2,2] (row,col), row<0 = absent, NULL = no apples;
 * eaters are agents 0..min(2,n)-1.  Outputs: new_positions [C,4,2], crash/restricted [C,4] (0/1),
 * caught [C,2,2] = times eater i stood on apple k over the 4 sub-steps. */
int gw_update_world(gw_handle* h, int64_t n_cases, const int8_t* n_agents_per_case, const int8_t* positions,
                    const int8_t* actions, const int8_t* apples, int8_t* new_positions, uint8_t* crash,
                    uint8_t* restricted, int8_t* caught, void* stream);

/* Operator level: C independent FeAR_4_one_actor evaluations.  in_list [C,4] marks the agents present in
 * ActionID4Agents (the actor's entry is forced on).  Outputs resp [C,4] fp64 (row `actor` of the matrix),
 * n_mdr / n_act [C,4] valid-move counts. */
int gw_fear_one_actor(gw_handle* h, int64_t n_cases, const int8_t* n_agents_per_case, const int8_t* positions,
                      const int8_t* actions, const int8_t* mdr, const int8_t* actor, const uint8_t* in_list,
                      double* resp, int8_t* n_mdr, int8_t* n_act, void* stream);

/* Operator level: Responsibility.FeAR for ALL actors (custom/Responsibility.py:57-132): resp / n_mdr / n_act [C,4,4]
 * (row = actor, column = affected).  Unlike gw_fear_one_actor the actor is NOT forced into the list: an actor outside
 * ActionID4Agents cannot be swapped, so its row is zero, exactly as in the reference. */
int gw_fear_matrix(gw_handle* h, int64_t n_cases, const int8_t* n_agents_per_case, const int8_t* positions,
                   const int8_t* actions, const int8_t* mdr, const uint8_t* in_list, double* resp, int8_t* n_mdr,
                   int8_t* n_act, void* stream);

/* Operator level: Responsibility.FeAL (custom/Responsibility.py:213-303): per agent, the valid moves it has when every
 * other listed agent plays its action (n_act) vs its Move de Rigueur (n_mdr); feal = clip(n_act / (n_mdr + 1e-6), -1, 1).
 * Outputs [C,4]. */
int gw_feal(gw_handle* h, int64_t n_cases, const int8_t* n_agents_per_case, const int8_t* positions,
            const int8_t* actions, const int8_t* mdr, const uint8_t* in_list, double* feal, int8_t* n_mdr,
            int8_t* n_act, void* stream);

/* ---- replay sampling (K4, read side): `memory.sample(BATCH_SIZE)` as called at maddpg/agent.py:209-211 -----------
 * The replay ring is caller-owned device memory (time-major: T slots x E envs) that gw_step filled through gw_io:
 * slot t % T holds observation t, slot (t+1) % T the next one; `final_obs` holds the terminal observation of the rows
 * whose episode ended in that step (`ended` != 0).  One kernel draws `batch` (time, env) pairs uniformly over the
 * newest min(t_now, T-1) time steps (with replacement; Philox(seed, sample index, draw)) and gathers
 * state / action / reward / next_state / done into f32 batch tensors -- the fields of maddpg/agent.py:70. */
typedef struct gw_replay_view {
  uint32_t struct_size;          /* sizeof(gw_replay_view) */
  int32_t obs_dtype;             /* GW_OBS_F32 / GW_OBS_BF16: element type of obs / final_obs */
  int64_t slots;                 /* T >= 3 */
  int64_t num_envs;              /* E */
  int32_t n_learners;            /* L */
  int32_t obs_len;               /* H*W */
  int32_t action_dim;            /* 9 */
  int32_t pad_;
  const void* obs;               /* [T, E, L, obs_len] */
  const void* final_obs;         /* [T, E, L, obs_len] */
  const float* action;           /* [T, E, L, action_dim] the actor's continuous action vectors */
  const float* reward;           /* [T, E, L] what the trainer stores (the shaped reward) */
  const uint8_t* terminated;     /* [T, E, L] */
  const uint8_t* ended;          /* [T, E] */
} gw_replay_view;

/* t_now = transitions recorded per env so far.  t_in / env_in (device int64 [batch], both or neither): take these
 * indices instead of drawing (absolute time steps within the stored window).  Outputs (device): state / next_state
 * f32 [batch, L, obs_len], action f32 [batch, L, action_dim], reward / done f32 [batch, L]; t_out / env_out int64
 * [batch], nullable.  `h` provides the device, the error string and the launch count.  Stream-ordered, no sync. */
int gw_replay_sample(gw_handle* h, const gw_replay_view* ring, int64_t t_now, int64_t batch, uint64_t seed,
                     uint64_t draw, const int64_t* t_in, const int64_t* env_in, float* state, float* action,
                     float* reward, float* next_state, float* done, int64_t* t_out, int64_t* env_out, void* stream);

/* ---- MADDPG update, element-wise part: LayerNorm + ReLU of the reference's networks -----------------
 * (Linear -> LayerNorm -> ReLU, hidden 128: configs/mlp.yaml:3) as one kernel forward and one backward; the
 * GEMMs of the update stay with cuBLAS.  width must be 128; all pointers device, fp32, 16-byte aligned rows.
 * forward: y = relu((x - mean) * rstd * gamma + beta), mean / rstd [rows] kept for the backward.
 * backward: dx [rows, 128], dgamma / dbeta [128] (overwritten, not accumulated). */
int gw_ln_relu_forward(gw_handle* h, int64_t rows, int32_t width, const float* x, const float* gamma, const float* beta,
                       float eps, float* y, float* mean, float* rstd, void* stream);
int gw_ln_relu_backward(gw_handle* h, int64_t rows, int32_t width, const float* dy, const float* x, const float* mean,
                        const float* rstd, const float* gamma, const float* beta, float* dx, float* dgamma, float* dbeta,
                        void* stream);

/* Backward of y = x W^T + b (torch.nn.Linear; x [batch, in] with row stride x_row_stride, W [out, in], dy [batch, out]
 * contiguous) in one launch: dW [out, in] = dy^T x, db [out] = column sums of dy, dx [batch, in] = dy W (dx nullable:
 * the input needs no gradient).  fp32 FMA, outputs overwritten. */
int gw_linear_backward(gw_handle* h, int32_t batch, int32_t in_features, int32_t out_features, const float* dy,
                       const float* x, int32_t x_row_stride, const float* w, float* dw, float* db, float* dx,
                       void* stream);

/* ---- fused MADDPG update: `agent.learn(experiences)` as called at maddpg/agent.py:209-213, :218-224 --------
 * One persistent cooperative kernel runs whole updates: (optionally) draw + gather the batch from the replay ring, target
 * actors, TD target, critic forward / backward, Adam, the actor loss through the updated critic, actor backward, Adam,
 * soft target update (TAU) -- for all agents, `updates` times per launch (the reference learns num_envs // LEARN_STEP
 * times after every env step, maddpg/agent.py:214-224).  Networks: the reference's shapes: actor
 * Linear(obs_len,128)-LayerNorm-ReLU-Linear(128,128)-LayerNorm-ReLU-Linear(128,action_dim)-GumbelSoftmax, critic
 * Linear(n*(obs_len+action_dim),128)-LN-ReLU-Linear(128,128)-LN-ReLU-Linear(128,1).  fp32 FMA arithmetic throughout
 * (the reference trains in fp32; the update is latency-bound, not FLOP-bound: 0.23 GFLOP per update).
 * All buffers are caller-owned DEVICE memory (PyTorch tensors); parameters live in flat fp32 vectors in which every
 * network is one block in torch parameter order (w1,b1,ln1_g,ln1_b,w2,b2,ln2_g,ln2_b,w3,b3), actors first, then critics,
 * each block starting at a multiple of 4 floats -- the nn.Parameters of the Python twin are views into them. */
typedef struct gw_learner_config {
  uint32_t struct_size;          /* sizeof(gw_learner_config) */
  int32_t n_agents;              /* 1..GW_MAX_LEARNERS */
  int32_t obs_len;               /* H*W, a multiple of 16 */
  int32_t action_dim;            /* 9 */
  int32_t batch;                 /* BATCH_SIZE: a multiple of 32, <= 512 */
  float lr_actor, lr_critic;     /* LR_ACTOR, LR_CRITIC (Adam, betas / eps below) */
  float gamma, tau;              /* GAMMA, TAU */
  float beta1, beta2, adam_eps;  /* 0.9, 0.999, 1e-8 */
  float ln_eps;                  /* 1e-5 */
  uint64_t seed;                 /* Gumbel noise of the actors' output activation: Philox(seed; row, agent, update) */
} gw_learner_config;

typedef struct gw_learner_layout {
  uint32_t struct_size;          /* sizeof(gw_learner_layout) */
  int32_t n_nets;                /* 2 * n_agents: actors 0..n-1, critics n..2n-1 */
  int64_t net_offset[2 * GW_MAX_LEARNERS];   /* floats, into every parameter-shaped vector */
  int64_t net_params[2 * GW_MAX_LEARNERS];   /* 38 793 / 60 545 for the reference's shapes */
  int64_t param_floats;          /* length of params / targets / adam_m / adam_v / grads */
  int64_t scratch_bytes;         /* activations, partial sums, batch staging, grid barrier */
} gw_learner_layout;

typedef struct gw_learner_buffers {
  float* params;                 /* [param_floats] online networks */
  float* targets;                /* [param_floats] target networks */
  float* adam_m;                 /* [param_floats] exp_avg */
  float* adam_v;                 /* [param_floats] exp_avg_sq */
  float* grads;                  /* [param_floats] gradients of the last update (what a data-parallel run all-reduces) */
  float* adam_steps;             /* [n_nets] f32 step counters (torch.optim.Adam(capturable=True) keeps the same) */
  void* scratch;                 /* [scratch_bytes], 256-byte aligned, zero-filled once by the caller */
} gw_learner_buffers;

/* an explicit batch (device, f32, contiguous): state / next_state [B, n, obs_len], action [B, n, action_dim], reward / done
 * [B, n].  gumbel_next / gumbel_cur [B, n, action_dim], nullable: Gumbel noise g (already -log(-log u)) added to the logits
 * of the target actors (on next_state) / the actors (on state) instead of the kernel's own Philox draws (tests). */
typedef struct gw_learn_batch {
  const float* state; const float* action; const float* reward; const float* next_state; const float* done;
  const float* gumbel_next; const float* gumbel_cur;
} gw_learn_batch;

typedef struct gw_learner gw_learner;
#define GW_LEARN_ALL 0           /* the whole update */
#define GW_LEARN_CRITIC_GRADS 1  /* data-parallel runs: up to the critics' gradients (in `grads`) ... all-reduce them ... */
#define GW_LEARN_ACTOR_GRADS 2   /* ... critic Adam steps, actor loss, up to the actors' gradients ... all-reduce ... */
#define GW_LEARN_FINISH 3        /* ... actor Adam steps, soft updates */
int gw_learner_layout_of(const gw_learner_config* cfg, gw_learner_layout* out);   /* host arithmetic only */
int gw_learner_create(gw_handle* h, const gw_learner_config* cfg, const gw_learner_buffers* buf, gw_learner** out);
int gw_learner_destroy(gw_learner* l);
/* `updates` consecutive updates in one launch.  Exactly one of `batch` (updates must be 1) and `ring` is given; with
 * `ring` update u draws its batch like gw_replay_sample(seed = sample_seed, draw = draw_base + u) and gathers it inside
 * the kernel.  segment: GW_LEARN_ALL, or one of the three pieces of a data-parallel update (updates must be 1; grads are
 * multiplied by grad_scale -- 1 / world size after a SUM all-reduce -- when they are applied).  losses: device f32
 * [updates, 2, n_agents] (actor loss, critic loss per agent), nullable.  Stream-ordered, no synchronisation, capturable. */
int gw_learner_update(gw_learner* l, const gw_learn_batch* batch, const gw_replay_view* ring, int64_t t_now,
                      uint64_t sample_seed, uint64_t draw_base, int32_t updates, int32_t segment, float grad_scale,
                      float* losses, void* stream);
/* Two implementations of the same update: GW_LEARN_KERNEL_PHASE (csrc/gw_maddpg.cu: one CTA per SM, a grid barrier
 * between the ~19 steps, fp32 FMA, every supported shape) and GW_LEARN_KERNEL_CLUSTER (csrc/gw_maddpg_cluster.cu: clusters
 * of 4 CTAs carry 16 batch rows through a whole forward / backward chain over distributed shared memory, 3xTF32 tensor-core
 * tiles, four grid barriers per update; the reference's shape only: 2 learners, obs_len 160, batch % 16 == 0, batch <= 256 on
 * a B200).  AUTO = cluster where possible.  gw_learner_kernel returns the kind the next update will run. */
#define GW_LEARN_KERNEL_AUTO 0
#define GW_LEARN_KERNEL_PHASE 1
#define GW_LEARN_KERNEL_CLUSTER 2
int gw_learner_set_kernel(gw_learner* l, int32_t kind);
int gw_learner_kernel(const gw_learner* l);
/* Data-parallel training, gradient exchange INSIDE the update kernel (cluster kernel, GW_LEARN_ALL): one process per GPU,
 * every rank exports a small exchange block (one flat gradient slot per rank + arrival words, a cudaMalloc of the library's own) as
 * a CUDA IPC handle, the host exchanges the handles (torch.distributed all_gather), every rank connects.  From then on each
 * Adam phase of gw_learner_update pushes this rank's gradient into its slot of EVERY rank's block (posted stores over NVLink),
 * passes a barrier across ALL ranks' grids (arrival words stored into the peers' blocks) and adds the world's gradients from
 * its own block in rank order: no NCCL call and no host round trip per update, `updates` per launch as on one GPU, parameters bit-identical on all ranks.
 * Every rank must issue the same sequence of gw_learner_update calls.  A peer that does not arrive within 5 s
 * (GW_PEER_TIMEOUT_MS) is reported by gw_learner_peer_status instead of hanging the GPU.  The segmented calls
 * (GW_LEARN_CRITIC_GRADS ...) remain for exchanges done by the caller (NCCL: maddpg/agent.py has no counterpart, the reference
 * trains on one device). */
#define GW_MAX_PEERS 8
typedef struct gw_peer_handle { uint8_t bytes[64]; } gw_peer_handle;
int gw_learner_peer_export(gw_learner* l, gw_peer_handle* out);
int gw_learner_peer_connect(gw_learner* l, int32_t rank, int32_t world, const gw_peer_handle* handles /* [world] */);
int gw_learner_peer_status(gw_learner* l, int32_t* world, int32_t* timed_out);
int gw_learner_peer_disable(gw_learner* l);   /* exchanges are the caller's again (use when some rank failed to connect) */
/* intermediate tensors of the last update, by name (tests / debugging): "a2", "q", "y", "dq", "anew", "ax", "dz2", "dh1",
 * "dz1", "adz2", "adh1", "adz1", and "<pass>.<z1|h1|st1|z2|h2|st2>" with pass in ta / ct / c / ac / c2; index = agent */
int gw_learner_debug_ptr(gw_learner* l, const char* name, int index, float** ptr, int64_t* floats);

/* ---- actor forward (K5): AgileRL `MADDPG.get_action` as called at maddpg/agent.py:109-113 -----------------------
 * One actor per learner: Linear(H*W,128)-LayerNorm-ReLU-Linear(128,128)-LayerNorm-ReLU-Linear(128,9)-GumbelSoftmax
 * (shapes from the reference's checkpoints), Gaussian exploration noise, clip to [0,1], action mask,
 * arg-max.  The observation is never read: the first layer is evaluated from `obs_code` (an observation is the
 * constant map template plus <= 5 special cells, so W1*obs is a constant vector plus <= 5 columns of W1); the
 * 128x128 layer runs on the tensor cores (tcgen05.mma, bf16 inputs, fp32 accumulation in TMEM).
 * All weight pointers are HOST pointers to fp32 arrays in PyTorch layout ([out, in]); they are packed and uploaded. */
typedef struct gw_actor_weights {
  const float* w1; const float* b1; const float* ln1_g; const float* ln1_b;   /* [128, H*W], [128], [128], [128] */
  const float* w2; const float* b2; const float* ln2_g; const float* ln2_b;   /* [128, 128], [128], [128], [128] */
  const float* w3; const float* b3;                                           /* [9, 128], [9] */
} gw_actor_weights;

typedef struct gw_actor gw_actor;

/* `h` provides the map (template row), agent / learner counts and the device.  weights[k] belongs to learner k. */
int gw_actor_create(gw_handle* h, const gw_actor_weights* weights, int n_learners, gw_actor** out);
int gw_actor_update(gw_actor* a, const gw_actor_weights* weights, int n_learners, void* stream);   /* after a learn step */
/* the same from DEVICE tensors (fp32, the layouts above): packed by a kernel on `stream`, no host copy, no synchronisation */
int gw_actor_update_device(gw_actor* a, const gw_actor_weights* dev_weights, int n_learners, void* stream);
int gw_actor_destroy(gw_actor* a);
/* obs_code [E] (device, from gw_step / gw_reset); action_mask int8 [E, L, 9] (device, nullable);
 * cont_actions f32 [E, L, 9] and action_ids int8 [E, L] (device outputs).  Noise comes from Philox(seed, GLOBAL env id =
 * cfg.env_id_base + row, learner, step): a sharded rollout draws what the unsharded one draws.  training == 1: Gumbel noise of the output activation + Gaussian exploration noise (maddpg/agent.py:109-113,
 * training=True); training == 2: Gumbel noise only -- what the reference's evaluation does (customeval.py:90-94,
 * training=False: the GumbelSoftmax activation still samples on every forward); training == 0: plain softmax, no
 * noise (deterministic; tests). */
int gw_actor_forward(gw_actor* a, int64_t num_envs, const uint64_t* obs_code, const int8_t* action_mask,
                     float* cont_actions, int8_t* action_ids, int training, float expl_noise, float mean_noise,
                     uint64_t seed, uint64_t step, void* stream);

/* ---- GENERAL layout: grids up to 64 x 64, up to 16 world agents -------------------------------------
 * The packed layout above (16 bytes per env, 4-bit coordinates, four 8-bit cells, a pair table over 41 offsets) is sized
 * for the shipped scenarios: W = 16, H <= 16, <= 4 agents.  The reference itself is data-driven -- GWorld takes any map and
 * any number of agents (custom/grid_world.py:17-31, :104-150), both env classes read the map, N_Agents, Policies, MdRs and
 * restricted paths from the scenario dict (custom/ma_customenv.py:27-28, :338-365; custom/customenv.py:25-27, :186-240) --
 * so a scenario file with a larger map or more agents runs there.  gww_* is the second state layout for those: the same
 * step semantics (every rule cites the same reference lines), 64 bytes of state per env, 16-bit cells, per-cell move
 * restrictions instead of the folded next-cell table, and the collision rules evaluated on the paths themselves (no pair
 * table: offsets and agent counts are unbounded here).  A host binding picks the layout when it creates the world
 * (BatchedGridWorld does: packed when the scenario fits, general otherwise).  Learners stay agents 0..n_learners-1 with
 * n_learners <= 2 (N_INTELLIGENT_AGENTS and the apple dict are constants of the reference's env modules).
 * gw_io / gw_stats are shared; shapes use this config's n_agents and H*W:
 *   obs / final_obs [E, L, H*W]; npc_actions [E, n_agents]; spawn / positions [E, n_agents, 2] (int8 row, col);
 *   info bits 8-15 as above, bits 0-3 / 4-7 crash / restricted of agents 0..3, bits 16-31 crash of agents 0..15;
 *   obs_code is not written (the actor kernel is tied to the packed layout).
 * Device RNG: the same Philox spec; spawn draw k uses call 0x100 + k/4, word k%4 (= the packed layout's for k < 4). */
#define GWW_MAX_AGENTS 16
#define GWW_MAX_DIM 64
#define GWW_MAX_CELLS (GWW_MAX_DIM * GWW_MAX_DIM)
#define GWW_MAX_BLOCKED 2048

typedef struct gww_config {
  int32_t struct_size;                 /* = sizeof(gww_config) */
  int32_t abi_version;                 /* = GW_ABI_VERSION */
  int32_t height, width;               /* 1..64 each */
  uint64_t map_rows[GWW_MAX_DIM];      /* bit c of map_rows[r] = Region[r][c] == 1 */
  int32_t n_agents;                    /* 1..16 */
  int32_t n_learners;                  /* 1..2, <= n_agents */
  int32_t env_kind;                    /* gw_env_kind */
  int8_t apple_row[GW_MAX_LEARNERS];
  int8_t apple_col[GW_MAX_LEARNERS];
  uint8_t policy_map[GWW_MAX_CELLS];   /* index row * width + col (dense, unlike the packed layout's pitch of 16) */
  uint8_t mdr_map[GWW_MAX_CELLS];
  int32_t n_policies;
  float step_weights[GW_MAX_POLICIES][3];
  float dir_weights[GW_MAX_POLICIES][4];
  double perturb_prob;
  int32_t fear;
  int32_t fear_radius;
  double fear_weight;
  int32_t max_steps;
  int32_t auto_reset;
  int32_t obs_dtype;
  int32_t device;
  int64_t num_envs;
  int64_t env_id_base;
  uint64_t seed;
  int32_t n_blocked;                   /* restricted paths as above; cells as (row << 8) | col */
  uint16_t blocked_from[GWW_MAX_BLOCKED];
  uint16_t blocked_to[GWW_MAX_BLOCKED];
} gww_config;

typedef struct gww_env_state {         /* gww_get_state / gww_set_state: 64 bytes per env */
  uint16_t cell[GWW_MAX_AGENTS];       /* (row << 8) | col; unused entries 0xFFFF */
  uint32_t flags;                      /* bits 0-1 apples present, 2-3 sticky terminations, 4 sticky truncation, 5-6 previous distance valid */
  uint32_t tick;                       /* RNG tick (resets + steps so far) */
  int32_t episode_return[GW_MAX_LEARNERS];   /* reward units (single env: tenths) */
  uint16_t prev_distance[GW_MAX_LEARNERS];
  uint32_t steps;                      /* steps in the running episode */
  uint32_t reserved[2];
} gww_env_state;

typedef struct gww_handle gww_handle;

int gww_default_config(gww_config* cfg);
int gww_create(const gww_config* cfg, gww_handle** out);
int gww_destroy(gww_handle* h);
const char* gww_last_error(const gww_handle* h);   /* h may be NULL (errors from gww_create) */
int gww_reset(gww_handle* h, const uint8_t* reset_mask, const gw_io* io, void* stream);   /* as gw_reset */
int gww_step(gww_handle* h, const gw_io* io, void* stream);                               /* as gw_step */
int gww_sync(gww_handle* h, void* stream);
size_t gww_state_bytes(const gww_handle* h);
int gww_get_state(gww_handle* h, void* dst, int dst_is_device, void* stream);
int gww_set_state(gww_handle* h, const void* src, int src_is_device, void* stream);
int gww_get_stats(gww_handle* h, gw_stats* host_out, void* stream);
int gww_reset_stats(gww_handle* h, void* stream);
int gww_launch_count(const gww_handle* h, uint64_t* kernels_launched);
/* Operator level, as gw_update_world / gw_fear_one_actor / gw_fear_matrix / gw_feal with every per-agent dimension of
 * length GWW_MAX_AGENTS: positions / new_positions [C,16,2], actions / mdr / crash / restricted / in_list / resp (one
 * actor) / n_mdr / n_act [C,16]; matrix outputs [C,16,16]; apples / caught [C,2,2]. */
int gww_update_world(gww_handle* h, int64_t n_cases, const int8_t* n_agents_per_case, const int8_t* positions,
                     const int8_t* actions, const int8_t* apples, int8_t* new_positions, uint8_t* crash,
                     uint8_t* restricted, int8_t* caught, void* stream);
int gww_fear_one_actor(gww_handle* h, int64_t n_cases, const int8_t* n_agents_per_case, const int8_t* positions,
                       const int8_t* actions, const int8_t* mdr, const int8_t* actor, const uint8_t* in_list,
                       double* resp, int8_t* n_mdr, int8_t* n_act, double* fear_sum /* [C] np.sum of the matrix, nullable */,
                       void* stream);
int gww_fear_matrix(gww_handle* h, int64_t n_cases, const int8_t* n_agents_per_case, const int8_t* positions,
                    const int8_t* actions, const int8_t* mdr, const uint8_t* in_list, double* resp, int8_t* n_mdr,
                    int8_t* n_act, void* stream);
int gww_feal(gww_handle* h, int64_t n_cases, const int8_t* n_agents_per_case, const int8_t* positions,
             const int8_t* actions, const int8_t* mdr, const uint8_t* in_list, double* feal, int8_t* n_mdr,
             int8_t* n_act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRIDWORLD_B200_H */
