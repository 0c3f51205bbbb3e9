"""ctypes binding of libgridworld_b200.so (include/gridworld_b200.h).  Fails loudly: there is no fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GW_LIB") or os.path.join(HERE, "csrc", "libgridworld_b200.so")

GW_MAX_AGENTS, GW_MAX_LEARNERS, GW_N_ACTIONS, GW_MAX_POLICIES, GW_MAX_H, GW_W = 4, 2, 9, 16, 16, 16
GW_MAX_BLOCKED = 256
GW_OK, GW_EINVAL, GW_ENOMEM, GW_ECUDA, GW_ENODEV, GW_ESTATE = 0, -1, -2, -3, -4, -5
GW_ENV_MULTI, GW_ENV_SINGLE = 0, 1
GW_OBS_F32, GW_OBS_BF16 = 0, 1

EXPORTS = ["gw_abi_version", "gw_build_info", "gw_default_config", "gw_create", "gw_destroy", "gw_last_error",
           "gw_reset", "gw_step", "gw_rollout", "gw_step_host", "gw_host_call_prepare", "gw_host_call_run", "gw_host_call_reset", "gw_server_stop", "gw_server_info", "gw_sync", "gw_state_bytes", "gw_get_state", "gw_set_state", "gw_get_stats",
           "gw_reset_stats", "gw_launch_count", "gw_debug_trace", "gw_update_world", "gw_fear_one_actor", "gw_fear_matrix", "gw_feal", "gw_actor_create", "gw_actor_update", "gw_actor_update_device",
           "gw_actor_destroy", "gw_actor_forward", "gw_replay_sample", "gw_ln_relu_forward", "gw_ln_relu_backward", "gw_linear_backward",
           "gw_learner_layout_of", "gw_learner_create", "gw_learner_destroy", "gw_learner_update", "gw_learner_debug_ptr", "gw_learner_set_kernel", "gw_learner_kernel",
           "gw_learner_peer_export", "gw_learner_peer_connect", "gw_learner_peer_status", "gw_learner_peer_disable",
           "gww_default_config", "gww_create", "gww_destroy", "gww_last_error", "gww_reset", "gww_step", "gww_sync",
           "gww_state_bytes", "gww_get_state", "gww_set_state", "gww_get_stats", "gww_reset_stats", "gww_launch_count", "gww_update_world",
           "gww_fear_one_actor", "gww_fear_matrix", "gww_feal"]


class GwActorWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w1", "b1", "ln1_g", "ln1_b", "w2", "b2", "ln2_g", "ln2_b", "w3", "b3")]


class GwReplayView(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("obs_dtype", C.c_int32), ("slots", C.c_int64), ("num_envs", C.c_int64),
                ("n_learners", C.c_int32), ("obs_len", C.c_int32), ("action_dim", C.c_int32), ("pad_", C.c_int32),
                ("obs", C.c_void_p), ("final_obs", C.c_void_p), ("action", C.c_void_p), ("reward", C.c_void_p),
                ("terminated", C.c_void_p), ("ended", C.c_void_p)]


class GwLearnerConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_agents", C.c_int32), ("obs_len", C.c_int32), ("action_dim", C.c_int32),
                ("batch", C.c_int32), ("lr_actor", C.c_float), ("lr_critic", C.c_float), ("gamma", C.c_float),
                ("tau", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float),
                ("ln_eps", C.c_float), ("seed", C.c_uint64)]


class GwLearnerLayout(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_nets", C.c_int32), ("net_offset", C.c_int64 * (2 * GW_MAX_LEARNERS)),
                ("net_params", C.c_int64 * (2 * GW_MAX_LEARNERS)), ("param_floats", C.c_int64), ("scratch_bytes", C.c_int64)]


class GwLearnerBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("params", "targets", "adam_m", "adam_v", "grads", "adam_steps", "scratch")]


class GwLearnBatch(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("state", "action", "reward", "next_state", "done", "gumbel_next", "gumbel_cur")]


GW_MAX_PEERS = 8


class GwPeerHandle(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 64)]


GW_LEARN_ALL, GW_LEARN_CRITIC_GRADS, GW_LEARN_ACTOR_GRADS, GW_LEARN_FINISH = 0, 1, 2, 3
GW_LEARN_KERNEL_AUTO, GW_LEARN_KERNEL_PHASE, GW_LEARN_KERNEL_CLUSTER = 0, 1, 2


class GwConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("abi_version", C.c_int32),
        ("height", C.c_int32), ("width", C.c_int32),
        ("map_rows", C.c_uint16 * GW_MAX_H),
        ("n_agents", C.c_int32), ("n_learners", C.c_int32), ("env_kind", C.c_int32),
        ("apple_row", C.c_int8 * GW_MAX_LEARNERS), ("apple_col", C.c_int8 * GW_MAX_LEARNERS),
        ("policy_map", C.c_uint8 * (GW_MAX_H * GW_W)), ("mdr_map", C.c_uint8 * (GW_MAX_H * GW_W)),
        ("n_policies", C.c_int32),
        ("step_weights", (C.c_float * 3) * GW_MAX_POLICIES), ("dir_weights", (C.c_float * 4) * GW_MAX_POLICIES),
        ("perturb_prob", C.c_double),
        ("fear", C.c_int32), ("fear_radius", C.c_int32), ("fear_weight", C.c_double),
        ("max_steps", C.c_int32), ("auto_reset", C.c_int32), ("obs_dtype", C.c_int32), ("device", C.c_int32),
        ("num_envs", C.c_int64), ("env_id_base", C.c_int64), ("seed", C.c_uint64),
        ("n_blocked", C.c_int32), ("blocked_from", C.c_uint8 * GW_MAX_BLOCKED), ("blocked_to", C.c_uint8 * GW_MAX_BLOCKED),
    ]


GWW_MAX_AGENTS, GWW_MAX_DIM, GWW_MAX_BLOCKED = 16, 64, 2048
GWW_MAX_CELLS = GWW_MAX_DIM * GWW_MAX_DIM


class GwwConfig(C.Structure):
    """gww_config: the general layout (grids up to 64 x 64, up to 16 agents)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("abi_version", C.c_int32),
        ("height", C.c_int32), ("width", C.c_int32),
        ("map_rows", C.c_uint64 * GWW_MAX_DIM),
        ("n_agents", C.c_int32), ("n_learners", C.c_int32), ("env_kind", C.c_int32),
        ("apple_row", C.c_int8 * GW_MAX_LEARNERS), ("apple_col", C.c_int8 * GW_MAX_LEARNERS),
        ("policy_map", C.c_uint8 * GWW_MAX_CELLS), ("mdr_map", C.c_uint8 * GWW_MAX_CELLS),
        ("n_policies", C.c_int32),
        ("step_weights", (C.c_float * 3) * GW_MAX_POLICIES), ("dir_weights", (C.c_float * 4) * GW_MAX_POLICIES),
        ("perturb_prob", C.c_double),
        ("fear", C.c_int32), ("fear_radius", C.c_int32), ("fear_weight", C.c_double),
        ("max_steps", C.c_int32), ("auto_reset", C.c_int32), ("obs_dtype", C.c_int32), ("device", C.c_int32),
        ("num_envs", C.c_int64), ("env_id_base", C.c_int64), ("seed", C.c_uint64),
        ("n_blocked", C.c_int32), ("blocked_from", C.c_uint16 * GWW_MAX_BLOCKED), ("blocked_to", C.c_uint16 * GWW_MAX_BLOCKED),
    ]


class GwwEnvState(C.Structure):
    _fields_ = [("cell", C.c_uint16 * GWW_MAX_AGENTS), ("flags", C.c_uint32), ("tick", C.c_uint32),
                ("episode_return", C.c_int32 * GW_MAX_LEARNERS), ("prev_distance", C.c_uint16 * GW_MAX_LEARNERS),
                ("steps", C.c_uint32), ("reserved", C.c_uint32 * 2)]


class GwIO(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in (
        "learner_actions", "npc_actions", "spawn", "obs", "final_obs", "reward", "shaped_reward", "fear",
        "terminated", "truncated", "ended", "action_mask", "positions", "info", "obs_code")]


class GwRolloutPlan(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("steps", C.c_int32), ("ring_slots", C.c_int64), ("first_slot", C.c_int64),
                ("action_slots", C.c_int64), ("first_action", C.c_int64)]


class GwStats(C.Structure):
    _fields_ = [("env_steps", C.c_uint64), ("agent_steps", C.c_uint64), ("episodes", C.c_uint64),
                ("episode_len_sum", C.c_uint64), ("crashes", C.c_uint64), ("apples", C.c_uint64),
                ("unresolved", C.c_uint64), ("fear_nonzero", C.c_uint64), ("return_sum", C.c_double),
                ("fear_sum", C.c_double), ("fear_tasks", C.c_uint64)]


_lib = None


def load():
    """Load the shared library once.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m marl_responsible_nav_b200.build` "
            "(or __graft_entry__.build()).  This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i64 = C.c_void_p, C.c_int64
    lib.gw_abi_version.restype = C.c_int
    lib.gw_build_info.restype = C.c_char_p
    lib.gw_default_config.argtypes = [C.POINTER(GwConfig)]
    lib.gw_create.argtypes = [C.POINTER(GwConfig), C.POINTER(vp)]
    lib.gw_destroy.argtypes = [vp]
    lib.gw_last_error.argtypes = [vp]
    lib.gw_last_error.restype = C.c_char_p
    lib.gw_reset.argtypes = [vp, vp, C.POINTER(GwIO), vp]
    lib.gw_step.argtypes = [vp, C.POINTER(GwIO), vp]
    lib.gw_rollout.argtypes = [vp, C.POINTER(GwIO), C.POINTER(GwRolloutPlan), vp]
    lib.gw_sync.argtypes = [vp, vp]
    lib.gw_step_host.argtypes = [vp, C.POINTER(GwIO), vp, vp, vp, vp, C.c_int, vp]
    lib.gw_host_call_prepare.argtypes = [vp, C.POINTER(GwIO), vp, vp, vp, vp, C.c_int, C.POINTER(C.c_int)]
    lib.gw_host_call_run.argtypes = [vp, C.c_int, vp]
    lib.gw_host_call_reset.argtypes = [vp]
    lib.gw_server_stop.argtypes = [vp]
    lib.gw_server_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    lib.gw_state_bytes.argtypes = [vp]
    lib.gw_state_bytes.restype = C.c_size_t
    lib.gw_get_state.argtypes = [vp, vp, C.c_int, vp]
    lib.gw_set_state.argtypes = [vp, vp, C.c_int, vp]
    lib.gw_get_stats.argtypes = [vp, C.POINTER(GwStats), vp]
    lib.gw_reset_stats.argtypes = [vp, vp]
    lib.gw_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.gw_update_world.argtypes = [vp, i64] + [vp] * 8 + [vp]
    lib.gw_fear_one_actor.argtypes = [vp, i64] + [vp] * 9 + [vp]
    lib.gw_fear_matrix.argtypes = [vp, i64] + [vp] * 8 + [vp]
    lib.gw_feal.argtypes = [vp, i64] + [vp] * 8 + [vp]
    lib.gw_actor_create.argtypes = [vp, C.POINTER(GwActorWeights), C.c_int, C.POINTER(vp)]
    lib.gw_actor_update.argtypes = [vp, C.POINTER(GwActorWeights), C.c_int, vp]
    lib.gw_actor_update_device.argtypes = [vp, C.POINTER(GwActorWeights), C.c_int, vp]
    lib.gw_actor_destroy.argtypes = [vp]
    lib.gw_actor_forward.argtypes = [vp, i64, vp, vp, vp, vp, C.c_int, C.c_float, C.c_float, C.c_uint64, C.c_uint64, vp]
    lib.gw_replay_sample.argtypes = [vp, C.POINTER(GwReplayView), i64, i64, C.c_uint64, C.c_uint64] + [vp] * 9 + [vp]
    lib.gw_ln_relu_forward.argtypes = [vp, i64, C.c_int32, vp, vp, vp, C.c_float, vp, vp, vp, vp]
    lib.gw_ln_relu_backward.argtypes = [vp, i64, C.c_int32] + [vp] * 9 + [vp]
    lib.gw_linear_backward.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_int32, vp, vp, vp, vp, vp]
    lib.gw_learner_layout_of.argtypes = [C.POINTER(GwLearnerConfig), C.POINTER(GwLearnerLayout)]
    lib.gw_learner_create.argtypes = [vp, C.POINTER(GwLearnerConfig), C.POINTER(GwLearnerBuffers), C.POINTER(vp)]
    lib.gw_learner_destroy.argtypes = [vp]
    lib.gw_learner_update.argtypes = [vp, C.POINTER(GwLearnBatch), C.POINTER(GwReplayView), i64, C.c_uint64, C.c_uint64,
                                      C.c_int32, C.c_int32, C.c_float, vp, vp]
    lib.gw_learner_peer_export.argtypes = [vp, C.POINTER(GwPeerHandle)]
    lib.gw_learner_peer_connect.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(GwPeerHandle)]
    lib.gw_learner_peer_disable.argtypes = [vp]
    lib.gw_learner_peer_status.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.gw_learner_set_kernel.argtypes = [vp, C.c_int32]
    lib.gw_learner_kernel.argtypes = [vp]
    lib.gw_learner_debug_ptr.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    lib.gww_default_config.argtypes = [C.POINTER(GwwConfig)]
    lib.gww_create.argtypes = [C.POINTER(GwwConfig), C.POINTER(vp)]
    lib.gww_destroy.argtypes = [vp]
    lib.gww_last_error.argtypes = [vp]
    lib.gww_last_error.restype = C.c_char_p
    lib.gww_reset.argtypes = [vp, vp, C.POINTER(GwIO), vp]
    lib.gww_step.argtypes = [vp, C.POINTER(GwIO), vp]
    lib.gww_sync.argtypes = [vp, vp]
    lib.gww_state_bytes.argtypes = [vp]
    lib.gww_state_bytes.restype = C.c_size_t
    lib.gww_get_state.argtypes = [vp, vp, C.c_int, vp]
    lib.gww_set_state.argtypes = [vp, vp, C.c_int, vp]
    lib.gww_get_stats.argtypes = [vp, C.POINTER(GwStats), vp]
    lib.gww_reset_stats.argtypes = [vp, vp]
    lib.gww_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.gww_update_world.argtypes = [vp, i64] + [vp] * 8 + [vp]
    lib.gww_fear_one_actor.argtypes = [vp, i64] + [vp] * 10 + [vp]
    lib.gww_fear_matrix.argtypes = [vp, i64] + [vp] * 8 + [vp]
    lib.gww_feal.argtypes = [vp, i64] + [vp] * 8 + [vp]
    if lib.gw_abi_version() != 1:
        raise RuntimeError("libgridworld_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, handle=None, what="", wide=False):
    if rc == GW_OK:
        return
    msg = load().gww_last_error(handle) if wide else load().gw_last_error(handle)
    raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def build_config(scenario, num_envs=1, env_kind="multi", fear=True, fear_weight=0.0, fear_radius=5, n_agents=None,
                 n_learners=None, apples=None, max_steps=150, auto_reset=True, obs_bf16=False, seed=0, env_id_base=0,
                 perturb_prob=0.25, device=0) -> GwConfig:
    """Fill a gw_config from a Scenario (host only; no library call, no GPU needed)."""
    kind = {"multi": GW_ENV_MULTI, "single": GW_ENV_SINGLE}[env_kind]
    H, W = scenario.shape
    n_agents = int(n_agents if n_agents is not None else scenario.n_agents)
    n_learners = int(n_learners if n_learners is not None else (2 if env_kind == "multi" else 1))
    if apples is None:
        apples = ((9, 0), (5, 10))[:n_learners] if env_kind == "multi" else ((9, 15),)   # ma_customenv.py:422 / customenv.py:334
    cfg = GwConfig()
    cfg.struct_size, cfg.abi_version = C.sizeof(GwConfig), 1
    cfg.height, cfg.width = H, W
    for r, bits in enumerate(scenario.map_rows()):
        cfg.map_rows[r] = bits
    cfg.n_agents, cfg.n_learners, cfg.env_kind = n_agents, n_learners, kind
    for k in range(GW_MAX_LEARNERS):
        a = apples[k] if k < len(apples) else None
        cfg.apple_row[k], cfg.apple_col[k] = (int(a[0]), int(a[1])) if a is not None else (-1, -1)
    for r in range(H):
        for c in range(W):
            cfg.policy_map[r * GW_W + c] = int(scenario.policy_index[r, c])
            cfg.mdr_map[r * GW_W + c] = int(scenario.mdr_action[r, c])
    cfg.n_policies = len(scenario.policies)
    for i, (sw, dw) in enumerate(scenario.policies):
        for s_ in range(3):
            cfg.step_weights[i][s_] = float(sw[s_])
        for d in range(4):
            cfg.dir_weights[i][d] = float(dw[d])
    cfg.perturb_prob = float(perturb_prob)
    cfg.fear, cfg.fear_radius, cfg.fear_weight = int(bool(fear)), int(fear_radius), float(fear_weight)
    cfg.max_steps, cfg.auto_reset = int(max_steps), int(bool(auto_reset))
    cfg.obs_dtype = GW_OBS_BF16 if obs_bf16 else GW_OBS_F32
    cfg.device = int(device)
    cfg.num_envs, cfg.env_id_base, cfg.seed = int(num_envs), int(env_id_base), int(seed) & (2 ** 64 - 1)
    blocked = list(getattr(scenario, "blocked", None) or [])
    if len(blocked) > GW_MAX_BLOCKED:
        raise ValueError(f"at most {GW_MAX_BLOCKED} restricted paths (walls count twice)")
    cfg.n_blocked = len(blocked)
    for k, (a, b) in enumerate(blocked):
        cfg.blocked_from[k], cfg.blocked_to[k] = (int(a[0]) << 4) | int(a[1]), (int(b[0]) << 4) | int(b[1])
    return cfg


def fits_packed_layout(scenario, n_agents=None) -> bool:
    """True when the scenario fits the packed 16-byte state (W = 16, H <= 16, <= 4 agents, <= 256 restricted paths)."""
    H, W = scenario.shape
    n = int(n_agents if n_agents is not None else scenario.n_agents)
    return W == GW_W and H <= GW_MAX_H and n <= GW_MAX_AGENTS and len(getattr(scenario, "blocked", None) or []) <= GW_MAX_BLOCKED


def build_wide_config(scenario, num_envs=1, env_kind="multi", fear=True, fear_weight=0.0, fear_radius=5, n_agents=None,
                      n_learners=None, apples=None, max_steps=150, auto_reset=True, obs_bf16=False, seed=0, env_id_base=0,
                      perturb_prob=0.25, device=0) -> GwwConfig:
    """Fill a gww_config (general layout) from a Scenario: the same arguments and defaults as `build_config`."""
    kind = {"multi": GW_ENV_MULTI, "single": GW_ENV_SINGLE}[env_kind]
    H, W = scenario.shape
    if not (1 <= H <= GWW_MAX_DIM and 1 <= W <= GWW_MAX_DIM):
        raise ValueError(f"grid {H}x{W}: the general layout holds up to {GWW_MAX_DIM}x{GWW_MAX_DIM}")
    n_agents = int(n_agents if n_agents is not None else scenario.n_agents)
    n_learners = int(n_learners if n_learners is not None else (2 if env_kind == "multi" else 1))
    if apples is None:
        apples = ((9, 0), (5, 10))[:n_learners] if env_kind == "multi" else ((9, 15),)   # ma_customenv.py:422 / customenv.py:334
    cfg = GwwConfig()
    cfg.struct_size, cfg.abi_version = C.sizeof(GwwConfig), 1
    cfg.height, cfg.width = H, W
    for r, bits in enumerate(scenario.map_rows()):
        cfg.map_rows[r] = bits
    cfg.n_agents, cfg.n_learners, cfg.env_kind = n_agents, n_learners, kind
    for k in range(GW_MAX_LEARNERS):
        a = apples[k] if k < len(apples) else None
        cfg.apple_row[k], cfg.apple_col[k] = (int(a[0]), int(a[1])) if a is not None else (-1, -1)
    pm = scenario.policy_index.astype("uint8").reshape(-1).tolist()
    mm = scenario.mdr_action.astype("uint8").reshape(-1).tolist()
    for i in range(H * W):
        cfg.policy_map[i], cfg.mdr_map[i] = pm[i], mm[i]
    if len(scenario.policies) > GW_MAX_POLICIES:
        raise ValueError(f"at most {GW_MAX_POLICIES} policy regions")
    cfg.n_policies = len(scenario.policies)
    for i, (sw, dw) in enumerate(scenario.policies):
        for s_ in range(3):
            cfg.step_weights[i][s_] = float(sw[s_])
        for d in range(4):
            cfg.dir_weights[i][d] = float(dw[d])
    cfg.perturb_prob = float(perturb_prob)
    cfg.fear, cfg.fear_radius, cfg.fear_weight = int(bool(fear)), int(fear_radius), float(fear_weight)
    cfg.max_steps, cfg.auto_reset = int(max_steps), int(bool(auto_reset))
    cfg.obs_dtype = GW_OBS_BF16 if obs_bf16 else GW_OBS_F32
    cfg.device = int(device)
    cfg.num_envs, cfg.env_id_base, cfg.seed = int(num_envs), int(env_id_base), int(seed) & (2 ** 64 - 1)
    blocked = list(getattr(scenario, "blocked", None) or [])
    if len(blocked) > GWW_MAX_BLOCKED:
        raise ValueError(f"at most {GWW_MAX_BLOCKED} restricted paths (walls count twice)")
    cfg.n_blocked = len(blocked)
    for k, (a, b) in enumerate(blocked):
        cfg.blocked_from[k], cfg.blocked_to[k] = (int(a[0]) << 8) | int(a[1]), (int(b[0]) << 8) | int(b[1])
    return cfg
