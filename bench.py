#!/usr/bin/env python
"""bench.py -- agent-steps/sec of the batched grid-world step path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E] [--fear 0|1] [--obs f32|bf16]

A "step" is one gw_step launch over E environments per GPU (weak scaling: E per GPU is fixed).  Default workload
is BASELINE.json configs[1]: custom_fear_5 (Level 3, 4 agents / 2 learners, FeAR on, weight -5), E = 4096, fp32
MLP observations.  Learner actions are synthetic (uniform over the 9 actions, pre-generated on the device); NPC
actions and respawns come from the device RNG; finished episodes auto-reset.

Prints ONE JSON line (rank 0).  `value` = device-timed whole-job agent-steps/s with inputs resident in HBM;
`e2e` = the same through the public API with host action buffers (pinned H2D per step, D2H of rewards/flags per
step); `roofline` = algorithmic bytes of the step kernel / measured launch duration vs the measured HBM peak;
`cpu_baseline` = the C restatement of the reference path (oracle/) on the host cores, bounded sample.
`--impl reference` times that CPU path alone (the reference is pure Python and does not travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

ALGO_BYTES_PER_AGENT_STEP = {"f32": 668.0, "bf16": 348.0}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/r1d_*_ncu_raw.csv; Level 3, f32 obs): (envs, fear) -> bytes
NCU_DRAM_BYTES = {(4096, 1): 370.0e3 + 0.0, (1 << 20, 1): 23.2e6 + 1413.0e6, (1 << 20, 0): 19.33e6 + 1378.3e6}
# gw_rollout_kernel<true,f32>, (envs, fear, steps per launch): one launch under ncu --set full (profiles/r2c_rollout_ncu_details.txt)
NCU_ROLLOUT_DRAM_BYTES = {(4096, 1, 20): 551.2e3 + 53.04e6}      # gw_rollout_split_kernel, profiles/r2r_rollout_split_ncu_details.txt


def kernel_name(envs, fear, obs, mode="step"):
    """The kernel gw_step picks for this batch (csrc/gw_kernels.cu: pick_small / pick_tile); gw_rollout has one."""
    f = "true" if fear else "false"
    split = fear and envs <= 32 * 148 and os.environ.get("GW_SPLIT", "1") != "0"     # one tile per SM: FeAR on helper warps
    if mode == "rollout":
        return f"gw_rollout_split_kernel<{obs}>" if split else f"gw_rollout_kernel<{f},{obs}>"
    if envs <= 6144:
        return f"gw_step_small_split_kernel<{obs}>" if split else f"gw_step_small_kernel<{f},{obs}>"
    tile = 32 if envs <= 24576 else (128 if envs <= 196608 else 256)
    return f"gw_step_kernel<256,{tile},{f},{obs}>"
METRIC = "agent-steps/sec (batched envs, device-timed) at 1/2/4/8 B200 vs CPU ref"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--fear", type=int, default=1)
    ap.add_argument("--obs", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--scenario", default="Level 3")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-python", action="store_true", help="skip timing the reference's own Python env")
    ap.add_argument("--python-seconds", type=float, default=8.0, help="seconds per setting (FeAR off / on) of the Python reference leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scale-points", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the config-5 extra (MADDPG training loop, custom_fear_10)")
    ap.add_argument("--train-steps", type=int, default=6, help="timed vector steps of the config-5 extra at the reference learn cadence")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every step from Python instead of replaying a CUDA graph")
    ap.add_argument("--mode", default="rollout", choices=["rollout", "step"],
                    help="rollout: gw_rollout, up to --rollout-steps env steps per launch (default); step: one gw_step launch per step")
    ap.add_argument("--rollout-steps", type=int, default=64, help="env steps per gw_rollout launch")
    return ap.parse_args()


def workload_name(a):
    return (f"custom_fear_5 ({a.scenario}, 4 agents / 2 learners, FeAR {'on, weight -5' if a.fear else 'off'}, "
            f"{a.envs} envs per GPU, {a.obs} MLP obs, uniform synthetic learner actions, device-RNG NPCs, auto-reset)")


def bench_config(a, world):
    """`config` of the JSON line: the same keys and values in both arms (ours / --impl reference)."""
    return {"workload": workload_name(a), "envs_per_gpu": a.envs, "global_envs": world * a.envs}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU path (oracle) timing
def time_cpu_oracle(a, seconds, steps=None, warmup=2):
    """The C restatement of the reference path on the host cores (all threads), bounded sample."""
    import numpy as np
    import c_oracle
    cores = os.cpu_count() or 1
    E = a.envs
    o = c_oracle.COracle(a.scenario, num_envs=E, threads=cores, fear=bool(a.fear), fear_weight=-5.0, auto_reset=True,
                         max_steps=150, seed=42, obs_bf16=(a.obs == "bf16"))
    o.reset()
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 9, size=(16, E, 2)).astype(np.int8)
    for i in range(warmup):
        o.step(acts[i % 16])
    n, t0 = 0, time.perf_counter()
    while True:
        o.step(acts[n % 16])
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and el >= seconds):
            break
    return {"value": E * o.L * n / el, "steps": n, "seconds": el, "cores": cores, "ms_per_step": 1e3 * el / n,
            "envs": E, "stats": o.stats()}


def time_reference_python(seconds=8.0):
    """The reference's own Python env (unmodified files staged into baseline/_ref by build()) on every host core, FeAR off
    and on, in a child process (forked workers must not inherit this process's CUDA context)."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "reference_python.py"), "--seconds", str(seconds)]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=seconds * 2 + 120)
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as exc:
        return {"kind": "reference-python", "unavailable": repr(exc)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bound the whole run: K + W steps of E envs each must end within ~2 minutes
    probe = time_cpu_oracle(a, seconds=1.0, warmup=1)
    budget = 100.0
    full = probe["ms_per_step"] * 1e-3 * (a.steps + a.warmup)
    sample_envs = a.envs
    if full > budget:
        sample_envs = max(64, int(a.envs * budget / full) // 64 * 64)
    a2 = argparse.Namespace(**vars(a))
    a2.envs = sample_envs
    r = time_cpu_oracle(a2, seconds=None, steps=a.steps, warmup=a.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "agent-steps/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32 cells + f64 FeAR", "data": "synthetic",
        "config": bench_config(a, a.gpus),
        "parallelism": "host threads (one CPU arm per run, rank 0)",
        "sample": f"{sample_envs} of {a.envs} envs per step, {a.steps} steps",
        "note": "C restatement of the reference path (oracle/gw_oracle.c, pinned to reference-recorded golden vectors), all "
                "host threads; the reference's own Python env is timed next to it (reference_python)",
        "cpu_baseline": {"value": r["value"], "unit": "agent-steps/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{sample_envs} envs x {a.steps} steps",
                         "reference_python": None if a.no_reference_python else time_reference_python(a.python_seconds)},
        "e2e": {"value": r["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU path
ROLL_FIELDS = ("obs", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "info", "positions", "obs_code",
               "action_mask")       # what a gw_step call of the bench writes as well (no final_obs)


def device_timed_rollout(a, E, fear, K, W, world, rank, dev, sample_clocks=False):
    """K timed env steps over E envs on this rank through gw_rollout: ceil(K / T) launches of up to T = --rollout-steps
    steps each, every step with its own action tensor and its own slot of the output rings (CUDA events on the launching
    stream, max over ranks)."""
    import torch
    import torch.distributed as dist
    from marl_responsible_nav_b200 import BatchedGridWorld
    obs_dtype = torch.float32 if a.obs == "f32" else torch.bfloat16
    env = BatchedGridWorld(a.scenario, num_envs=E, device=dev, fear=bool(fear), fear_weight=-5.0, auto_reset=True,
                           max_steps=150, obs_dtype=obs_dtype, seed=42, env_id_base=rank * E)
    L = env.n_learners
    obs_bytes = E * L * env.obs_len * (4 if a.obs == "f32" else 2)
    T = max(1, int(a.rollout_steps))
    slots = max(2, -(-(320 << 20) // obs_bytes), T + 1)          # ring larger than L2 (126 MB), as in device_timed
    slots = min(slots, max(2, (8 << 30) // obs_bytes))
    rings = env.new_rings(slots, fields=ROLL_FIELDS)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_act = 64
    actions = torch.randint(0, 9, (n_act, E, L), generator=gen, device=dev, dtype=torch.int8)
    env.reset(obs_out=rings.obs[0])

    def run(n, t0):
        done, launches = 0, 0
        while done < n:
            k = min(T, n - done)
            env.rollout(actions, k, rings, first_slot=(t0 + done) % slots, first_action=(t0 + done) % n_act)
            done += k
            launches += 1
        return launches

    run(W, 0)
    launches = run(K, W)                                          # the same launches as the timed region, untimed
    env.sync()
    # The timed launches are replayed from a CUDA graph (gw_rollout is capturable), as in device_timed: the timed region then
    # holds the K env steps and one graph launch, not the Python / ctypes marshalling of the rollout call (~50 us, a third of
    # a 20-step region).  --no-graph: the call itself is timed.
    graph = None
    if a.graph and K > 0:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run(K, W + K)
        graph.replay()                                            # warm-up of the graph itself
        env.sync()
    env.reset_stats()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(dev.index)
    if sample_clocks and rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # L2 flush (256 MiB > 126 MB) enqueued on the same stream right before the timed region: the events and the timed
    # launches queue behind it, so the region starts cold and holds device time only (with an idle GPU the host's
    # submission latency of the first launch, 5-10 us, would sit between the two events)
    flush.zero_()
    ev0.record()
    if graph is not None:
        graph.replay()
    else:
        launches = run(K, W + K)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if (sample_clocks and rank == 0) else None
    st = env.stats()
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    stat_vec = torch.tensor([st["episodes"], st["episode_len_sum"], st["crashes"], st["apples"], st["fear_nonzero"],
                             st["return_sum"], st["fear_sum"], st["unresolved"], st["fear_tasks"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(stat_vec, op=dist.ReduceOp.SUM)
    plan = {"mode": "rollout", "rollout_launches": launches, "steps_per_launch": min(T, K), "graph_steps": K if graph is not None else 0,
            "graph_replays": 1 if graph is not None else 0, "tail_graph_steps": 0, "eager_steps": 0}
    return {"ms": float(t_ms.item()), "stats": stat_vec, "env": env, "ring": rings.obs, "slots": slots, "L": L,
            "obs_bytes": obs_bytes, "clocks": clocks, "n_act": n_act, "plan": plan, "launches": launches}


def device_timed(a, E, fear, K, W, world, rank, dev, sample_clocks=False):
    """K timed gw_step launches over E envs on this rank (CUDA events on the launching stream, max over ranks)."""
    import torch
    import torch.distributed as dist
    from marl_responsible_nav_b200 import BatchedGridWorld
    obs_dtype = torch.float32 if a.obs == "f32" else torch.bfloat16
    env = BatchedGridWorld(a.scenario, num_envs=E, device=dev, fear=bool(fear), fear_weight=-5.0, auto_reset=True,
                           max_steps=150, obs_dtype=obs_dtype, seed=42, env_id_base=rank * E)
    L = env.n_learners
    obs_bytes = E * L * env.obs_len * (4 if a.obs == "f32" else 2)
    # observation ring: consecutive steps write different slots and the ring is larger than L2 (126 MB), so the store
    # stream cannot be absorbed by the cache -- this is also the layout of the device replay buffer
    slots = max(2, -(-(320 << 20) // obs_bytes))
    slots = min(slots, max(2, (8 << 30) // obs_bytes))
    ring = torch.empty((slots, E, L, env.obs_len), dtype=obs_dtype, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_act = 64
    actions = torch.randint(0, 9, (n_act, E, L), generator=gen, device=dev, dtype=torch.int8)
    env.reset(obs_out=ring[0])

    def run(n, t0=0):
        for t in range(t0, t0 + n):
            env.step(actions[t % n_act], obs_out=ring[(t + 1) % slots])

    run(W)
    env.sync()
    # The timed loop is replayed from CUDA graphs, whatever K is: one graph of g = min(K, n_act) consecutive steps (each
    # with its own action tensor and ring slot) replayed K // g times, plus a second graph for the K % g steps left, so
    # that the device never waits on a Python/ctypes call between launches (--no-graph: eager launches).
    graph = tail = None
    g_steps = min(K, n_act)
    plan = {"mode": "step", "graph_steps": 0, "graph_replays": 0, "tail_graph_steps": 0, "eager_steps": K}
    if a.graph and K > 0:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run(g_steps, W)
        if K % g_steps:
            tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(tail):
                run(K % g_steps, W + g_steps)
        env.sync()
        plan = {"mode": "step", "graph_steps": g_steps, "graph_replays": K // g_steps, "tail_graph_steps": K % g_steps, "eager_steps": 0}

    def timed(k):
        if graph is None:
            run(k, W)
            return
        for _ in range(k // g_steps):
            graph.replay()
        if k % g_steps:
            if tail is not None and k % g_steps == K % g_steps:
                tail.replay()
            else:
                run(k % g_steps, W)

    timed(K)                                                      # graph warm-up (the same launches as the timed region)
    env.sync()
    env.reset_stats()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(dev.index)
    if sample_clocks and rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()                      # L2 flush (256 MiB > 126 MB) on the same stream, the timed launches queue behind it (see device_timed_rollout)
    ev0.record()
    timed(K)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if (sample_clocks and rank == 0) else None
    st = env.stats()
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    stat_vec = torch.tensor([st["episodes"], st["episode_len_sum"], st["crashes"], st["apples"], st["fear_nonzero"],
                             st["return_sum"], st["fear_sum"], st["unresolved"], st["fear_tasks"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)          # time = max over ranks
        dist.all_reduce(stat_vec, op=dist.ReduceOp.SUM)      # episode statistics over NVLink (the only collective)
    return {"ms": float(t_ms.item()), "stats": stat_vec, "env": env, "ring": ring, "slots": slots, "L": L,
            "obs_bytes": obs_bytes, "clocks": clocks, "n_act": n_act, "plan": plan, "launches": K}


def masked_uniform_runs(a, E, dev, seeds=(0, 42, 66), steps=384):
    """the same workload with learner actions drawn uniformly among the ALLOWED actions (get_action_mask) --
    fewer wall bumps, another crash rate -- for three env seeds.  The action of step t+1 depends on the mask step t wrote,
    so the sampler (three small PyTorch kernels) sits between the steps and inside the timed region; the loop is replayed
    from a CUDA graph like the headline loop.  Reported next to the headline, never instead of it."""
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld
    out = []
    for seed in seeds:
        env = BatchedGridWorld(a.scenario, num_envs=E, device=dev, fear=bool(a.fear), fear_weight=-5.0, auto_reset=True,
                               max_steps=150, seed=seed)
        gen = torch.Generator(device=dev).manual_seed(seed)
        state = {"o": env.reset()}
        u = torch.empty((E, env.n_learners, 9), device=dev)

        def run(n):
            for _ in range(n):
                u.uniform_(0.01, 1.0, generator=gen)
                ids = (u * state["o"].action_mask).argmax(-1).to(torch.int8)
                state["o"] = env.step(ids)

        run(32)
        env.sync()
        g = torch.cuda.CUDAGraph()
        try:
            g.register_generator_state(gen)
            with torch.cuda.graph(g):
                run(64)
        except Exception:                                   # no graph-safe generator support: eager launches
            g = None
        env.sync()
        env.reset_stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        if g is None:
            run(steps)
        else:
            for _ in range(steps // 64):
                g.replay()
        ev1.record()
        torch.cuda.synchronize()
        st = env.stats()
        n = float(E * steps)                                # (the library's own step counter does not see graph replays)
        out.append({"seed": seed, "steps": steps, "graph": g is not None,
                    "agent_steps_per_s": E * env.n_learners * steps / (ev0.elapsed_time(ev1) * 1e-3),
                    "learner_crashes_per_env_step": st["crashes"] / n, "apples_per_env_step": st["apples"] / n,
                    "mean_episode_len": st["episode_len_sum"] / max(1, st["episodes"]),
                    "fear_nonzero_per_agent_step": st["fear_nonzero"] / (n * env.n_learners)})
        env.close()
    return out


def general_layout_extra(a, dev, cpu_seconds=4.0):
    """the GENERAL state layout (gww_*, csrc/gw_wide.cu) on a scenario the packed layout cannot hold -- the 20 x 28
    map with 7 agents, walls and one-ways that tests/golden/make_wide_golden.py ran the reference on.  Device-timed step
    launches (CUDA events around a 32-step CUDA graph), FeAR on / off, at the headline batch, at 65 536 and at 524 288 envs; next to
    it the C oracle built on gww_config, all host threads, on a bounded sample.  An extra line, never the headline."""
    import numpy as np
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld, load_scenario_json
    sc = load_scenario_json(os.path.join(ROOT, "tests", "golden", "wide_scenarios.json"), "Wide 20x28", walls="enforce")
    H, W = sc.shape
    algo = 64 + 64 + 2 + 8 + 8 + 16 + 4 + 1 + 4 + 18 + 2 * sc.n_agents + 2 * H * W * 4     # per env-step: state r/w, scalars, masks, positions, 2 obs rows
    peak, _ = peaks()
    out = {"scenario": f"{H}x{W} map, {sc.n_agents} agents, {len(sc.blocked)} restricted paths, 2 learners", "kernel": "gww_step_kernel",
           "algorithmic_bytes_per_env_step": algo, "points": []}
    for E, fear in ((a.envs, 1), (a.envs, 0), (65536, 1), (65536, 0), (524288, 1), (524288, 0)):
        env = BatchedGridWorld(sc, num_envs=E, device=dev, fear=bool(fear), fear_weight=-5.0, seed=42, max_steps=150)
        env.reset()
        acts = torch.randint(0, 9, (8, E, 2), dtype=torch.int8, device=dev)
        for t in range(8):
            env.step(acts[t % 8])
        env.sync()
        K = 32
        g = torch.cuda.CUDAGraph()                          # the launches replayed from a CUDA graph: kernel time, not ctypes time
        with torch.cuda.graph(g):
            for t in range(K):
                env.step(acts[t % 8])
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay()
        torch.cuda.synchronize()
        ev0.record()
        g.replay()
        ev1.record()
        torch.cuda.synchronize()
        per = ev0.elapsed_time(ev1) * 1e-3 / K
        st = env.stats()
        out["points"].append({"envs": E, "fear": bool(fear), "ms_per_step": per * 1e3, "agent_steps_per_s": E * 2 / per,
                              "achieved_gbs": algo * E / per / 1e9, "frac_of_hbm_peak": algo * E / per / 1e9 / peak,
                              "fear_tasks_per_env_step": st["fear_tasks"] / float(E * (8 + 3 * K))})
        env.close()
        del env, acts
        torch.cuda.empty_cache()
    try:
        import c_oracle
        threads = os.cpu_count() or 1
        Ec = 2048
        ora = c_oracle.COracle(sc, num_envs=Ec, threads=threads, fear=True, fear_weight=-5.0, seed=42)
        ora.reset()
        rng = np.random.default_rng(0)
        la = rng.integers(0, 9, size=(Ec, 2)).astype(np.int8)
        ora.step(la)
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < cpu_seconds:
            ora.step(la)
            n += 1
        el = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": Ec * 2 * n / el, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                               "sample": f"{Ec} envs x {n} steps ({el:.1f} s), oracle/gw_oracle.c built with -DGWO_WIDE, FeAR on"}
    except Exception as exc:
        out["cpu_baseline"] = {"failed": repr(exc)}
    return out


def train_extra(a, world, rank, dev):
    """BASELINE.json configs[4]: the full MADDPG loop (rollout through the actor kernel, device replay ring, update kernel)
    with custom_fear_10.yaml, a.envs environments per GPU.  Two cadences: the reference's (maddpg/agent.py:199-224 on the
    global env count: global_envs // LEARN_STEP updates after every vector step -- the update kernel is the cost) and
    round 1's batched one (one update every LEARN_STEP vector steps -- the rollout is the cost).  Device-timed region with a
    barrier + synchronize on both sides, max over ranks; gradients all-reduced over NCCL when world > 1."""
    import torch
    import torch.distributed as dist
    from marl_responsible_nav_b200 import maddpg
    hp = maddpg.preset("custom_fear_10")
    E = a.envs
    out = {"config": f"custom_fear_10 (Level 3, FeAR weight -10, BATCH_SIZE 128, LEARN_STEP 10, MEMORY_SIZE 200000), {E} envs per GPU",
           "envs_per_gpu": E, "global_envs": E * world}

    def timed(tr, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        u0, t0 = tr.updates_done, time.perf_counter()
        st = tr.train(k)
        torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        el = float(el.item())
        return {"vector_steps": k, "agent_steps_per_s": world * E * tr.env.n_learners * k / el, "updates": tr.updates_done - u0,
                "updates_per_s": (tr.updates_done - u0) / el, "seconds": el,
                "mean_return": st["return_sum"] / max(1, st["episodes"])}

    for cadence, warm, k in (("reference", 2, a.train_steps), ("batched", 120, 600)):
        env = maddpg.make_env(hp, E, device=dev, env_id_base=rank * E)
        tr = maddpg.BatchedTrainer(env, hp=hp, seed=hp["SEED"], learn_cadence=cadence, global_envs=E * world)
        exchange = tr.connect(0)
        tr.train(warm)
        r = timed(tr, k)
        r["updates_per_vector_step"] = tr.learn_schedule(0)
        r["update_kernel"] = tr.learner.kernel if tr.learner is not None else "torch graph"
        if world > 1:
            r["gradient_exchange"] = ("inside the update kernel over NVLink peer memory (all-gather of the ranks' flat gradients + local sum, "
                                      "one launch per block of updates)" if exchange == "peer" else
                                      "2 NCCL all-reduces per update (critics' / actors' flat gradients) between the update kernel's three segments")
        out[cadence + "_cadence"] = r
        del tr, env
        torch.cuda.empty_cache()
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, K, W = a.envs, a.steps, a.warmup
    rollout_mode = a.mode == "rollout" and E <= 65536           # beyond that the thread-per-env step kernel is the faster one
    per_step = None
    if rollout_mode:
        # the same K steps with one gw_step launch per step (CUDA-graph replay), reported next to the headline
        rs = device_timed(a, E, a.fear, K, W, world, rank, dev)
        per_step = {"value": world * E * rs["L"] * K / (rs["ms"] * 1e-3), "unit": "agent-steps/s", "ms_per_step": rs["ms"] / K,
                    "kernel": kernel_name(E, a.fear, a.obs), "launch_plan": rs["plan"]}
        rs["env"].close()
        del rs
        torch.cuda.empty_cache()
    r = (device_timed_rollout if rollout_mode else device_timed)(a, E, a.fear, K, W, world, rank, dev, sample_clocks=True)
    n_launches = r["launches"]
    ms, stat_vec, env, ring, slots, L, obs_bytes, n_act = (r[k] for k in ("ms", "stats", "env", "ring", "slots", "L", "obs_bytes", "n_act"))
    env_agents = env.n_agents
    clocks, r_plan = r["clocks"], r["plan"]
    value = world * E * L * K / (ms * 1e-3)

    # ---- e2e: public API, host action buffers, H2D + D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        Ke = min(K, 2000)
        host_actions = torch.randint(0, 9, (n_act, E, L), dtype=torch.int8).pin_memory()
        host_rew = torch.empty((E, L), dtype=torch.float32).pin_memory()
        host_end = torch.empty((E,), dtype=torch.uint8).pin_memory()
        # env.step_host = the public host-driven call: pinned actions in, rewards + ended flags out in pinned host memory,
        # complete on return.  Default mode for this batch: the resident step kernel (no launch / stream sync per step).
        import math

        # observation slots of the e2e loop: as many as action tensors, so that the (actions, slot) pairs repeat with period
        # n_act and every pair's marshalled argument set (the library keeps 1024) is prepared in the warm-up; 64 slots x 5.2 MB
        # still exceed L2
        e_slots = min(slots, n_act)
        act_views = [host_actions[i] for i in range(n_act)]
        slot_views = [ring[i] for i in range(e_slots)]

        def host_loop(n, **kw):
            for t in range(n):
                env.step_host(act_views[t % n_act], host_rew, host_end, obs_out=slot_views[t % e_slots], **kw)   # returns with the results on the host

        def timed_host_loop(**kw):
            host_loop(min(256, max(10, math.lcm(n_act, e_slots))), **kw)    # every (actions, ring slot) pair seen once: marshalling cached
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            # a stream synchronisation makes a resident step kernel leave the GPU: one more untimed call brings it back (part
            # of the warm-up, like the launch itself); the call returns with its results on the host, so nothing is in flight
            # when the clock starts
            host_loop(1, **kw)
            t0 = time.perf_counter()
            host_loop(Ke, **kw)
            el = time.perf_counter() - t0
            env.sync()
            t_e = torch.tensor([el], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
            return world * E * L * Ke / float(t_e.item())

        srv0 = env.server_info()
        e2e = {"value": timed_host_loop(), "unit": "agent-steps/s",
               "h2d_bytes_per_step": E * L, "d2h_bytes_per_step": E * L * 4 + E, "steps": Ke}
        srv1 = env.server_info()
        resident = srv1["launches"] > srv0["launches"]
        e2e["api"] = ("BatchedGridWorld.step_host (default" + (", resident step kernel): the kernel stays on the GPU between steps, "
                      "takes each step's doorbell from pinned host memory, loads the actions from the pinned host buffer, stores rewards + "
                      "ended flags into pinned host buffers over PCIe and signals completion through pinned host memory"
                      if resident else "): one gw_step launch whose kernel loads the step's actions from the pinned host buffer and "
                      "stores rewards + ended flags into pinned host buffers over PCIe, then a stream sync") +
                      "; observations stay in HBM (replay ring)")
        e2e["resident_kernel_launches"] = srv1["launches"] - srv0["launches"]
        # the same call, one launch + stream sync per step (resident=False), and with cudaMemcpyAsync H2D / D2H around
        # the kernel (zero_copy=False): reported aside
        if resident:
            e2e["launch_per_step_value"] = timed_host_loop(resident=False)
        e2e["memcpy_value"] = timed_host_loop(zero_copy=False)
    del env, ring, r

    # ---- masked-uniform learner actions, seeds {0, 42, 66}, N = 1 only
    variants = None
    if world == 1 and not a.no_scale_points:
        try:
            variants = masked_uniform_runs(a, E, dev)
        except Exception as exc:                            # an extra: must not cost the headline line
            variants = [{"failed": repr(exc)}]

    # ---- the same kernel at batches that fill the GPU: 1 M envs on one GPU (the step kernel's roofline point); with
    # several ranks BASELINE.json configs[3]'s split -- 10^6 envs over the ranks, statistics all-reduced
    scale_points = None
    if not a.no_scale_points:
        scale_points = []
        peak, _ = peaks()
        points = ((1 << 20, 1), (1 << 20, 0)) if world == 1 else ((1000000 // world, 1),)
        for Es, fs in points:
            try:
                torch.cuda.empty_cache()
                rs = device_timed(a, Es, fs, 128, 64, world, rank, dev)
                per = rs["ms"] * 1e-3 / 128
                gbs = ALGO_BYTES_PER_AGENT_STEP[a.obs] * Es * rs["L"] / per / 1e9
                scale_points.append({"envs_per_gpu": Es, "global_envs": Es * world, "fear": bool(fs), "kernel": kernel_name(Es, fs, a.obs),
                                     "ms_per_step": rs["ms"] / 128, "agent_steps_per_s": world * Es * rs["L"] / per,
                                     "achieved_gbs_per_gpu": gbs, "frac_of_hbm_peak": gbs / peak,
                                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_AGENT_STEP[a.obs] * Es * rs["L"],
                                     "launch_plan": rs["plan"],
                                     "traffic": NCU_DRAM_BYTES.get((Es, fs)) if (a.obs == "f32" and a.scenario == "Level 3") else None})
                del rs
            except Exception as exc:                        # an extra: must not cost the headline line
                scale_points.append({"envs_per_gpu": Es, "failed": repr(exc)})
                if world > 1:
                    break                                   # the ranks may be out of step: no further collectives in extras

    # ---- BASELINE.json configs[4]: the MADDPG training loop (every N)
    train = None
    if not a.no_train:
        try:
            train = train_extra(a, world, rank, dev)
        except Exception as exc:                            # an extra: must not cost the headline line
            train = {"failed": repr(exc)}

    general = None
    if world == 1 and not a.no_scale_points:
        try:
            general = general_layout_extra(a, dev)
        except Exception as exc:                            # an extra: must not cost the headline line
            general = {"failed": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    algo = ALGO_BYTES_PER_AGENT_STEP[a.obs] * E * L              # per launch (one GPU)
    achieved = algo / (ms * 1e-3 / K) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 cells / int32 rewards / f64 FeAR; obs " + a.obs, "data": "synthetic",
        "config": bench_config(a, world),
        "parallelism": f"env-shard x{world}, no per-step collective",
        "l2": f"L2 flushed (256 MiB write) on the launching stream immediately before the first event of the timed region (the timed launches are queued behind it: device time only); obs stores stream through a {slots}-slot ring of "
              f"{slots * obs_bytes / 2**20:.0f} MiB (> 126 MB L2), never re-read; the packed env state (16 B/env) is L2-resident by design",
        "launch_plan": r_plan,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(n_launches), "launch_per_step": per_step,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (None if not (a.obs == "f32" and a.scenario == "Level 3") else
                                 NCU_DRAM_BYTES.get((E, int(bool(a.fear)))) if r_plan["mode"] == "step" else
                                 NCU_ROLLOUT_DRAM_BYTES.get((E, int(bool(a.fear)), r_plan["steps_per_launch"]))),
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/r1d_*_ncu_raw.csv); "
                                       "at 4096 envs the stores are still in L2 when the kernel ends, so DRAM writes show as 0; gw_rollout (20 steps per launch, "
                                       "profiles/r2r_rollout_split_ncu_details.txt): 54 MB of the launch's 109 MB reach DRAM before it ends, the rest is written back from L2 later",
                     "kernel": kernel_name(E, a.fear, a.obs, r_plan["mode"]),
                     "algorithmic_bytes_per_launch": algo * K / n_launches, "algorithmic_bytes_per_step": algo, "peak_source": peak_src,
                     "launch_ms": ms / n_launches,
                     "note": "achieved = algorithmic bytes of the K timed steps / CUDA-event time over the timed region (includes launch gaps); "
                             + (f"gw_rollout: {n_launches} launch(es) of up to {r_plan['steps_per_launch']} env steps each" if r_plan["mode"] == "rollout"
                                else (f"CUDA graphs: {r_plan['graph_replays']} x {r_plan['graph_steps']} steps + {r_plan['tail_graph_steps']} steps"
                                      if r_plan["eager_steps"] == 0 else f"{r_plan['eager_steps']} eager launches"))
                             + f"; at {E} envs one step moves {algo / 2**20:.1f} MiB ({algo / peak / 1e3:.2f} us at peak): small batches are "
                               "latency-bound (a warp's dependent chain per step), see scale_points for the step kernel at 1M envs"},
        "scale_points": scale_points, "train": train, "masked_uniform_runs": variants, "general_layout": general,
        "workload_stats": {"episodes": stat_vec[0].item(), "mean_episode_len": stat_vec[1].item() / max(1.0, stat_vec[0].item()),
                           "learner_crashes_per_env_step": stat_vec[2].item() / (world * E * K),
                           "apples_per_env_step": stat_vec[3].item() / (world * E * K),
                           "fear_nonzero_per_agent_step": stat_vec[4].item() / (world * E * L * K),
                           "unresolved": stat_vec[7].item(),
                           "fear_tasks_per_env_step": stat_vec[8].item() / (world * E * K),
                           "counterfactual_sims_per_s": 18.0 * stat_vec[8].item() / (ms * 1e-3),
                           "reference_sims_per_env_step": 108},
        "env_steps_per_s": value / L, "entity_steps_per_s": value / L * env_agents,
    }
    if not a.no_cpu_baseline and world == 1:
        try:
            c = time_cpu_oracle(a, a.cpu_seconds)
            line["cpu_baseline"] = {"value": c["value"], "unit": "agent-steps/s", "cores": c["cores"], "kind": "port",
                                    "sample": f"{c['envs']} envs x {c['steps']} steps ({c['seconds']:.1f} s), C restatement "
                                              "of the reference path (oracle/gw_oracle.c), all host threads"}
            if not a.no_reference_python:
                line["cpu_baseline"]["reference_python"] = time_reference_python(a.python_seconds)
        except Exception as exc:   # the checker failing must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "agent-steps/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {exc}"}
    if world > 1:
        dist.destroy_process_group()
    return line


class _QuietStdout:
    """Everything libraries write to fd 1 while the bench runs (NCCL prints its version line there) goes to stderr, so
    that stdout carries exactly the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        with _QuietStdout():
            result = run_ours(args)
        if result is not None:
            print(json.dumps(result))
