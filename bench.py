#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused step kernel on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: island_navigation_ex level 9, default flags, uniform random
actions U{0..4} (Philox, keyed by global environment index), auto-reset in the step that ends an
episode.  1,048,576 environments PER GPU (the whole configuration on one GPU at N=1; weak scaling
above), which makes the per-step working set ~0.7 GB, several times the 126 MB L2.

A step = one launch of the fused kernel over the rank's whole batch.  `value` is device-timed with
inputs resident in HBM; `e2e` goes through the public host-buffer API (host_pipeline.HostPipeline:
per step the actions are uploaded from pinned host memory and the observation (uint8 ASCII board) +
reward rows + terminated flags are downloaded into pinned host memory, split-batch double buffered
so that one slice's transfers overlap the other's kernel).  `--impl reference` times the CPU oracle
port (the C restatement of the reference's per-step path, on all host threads, on the SAME batch
size and outputs as the GPU arm) and, when the unmodified Python reference is installed under
baseline/_ref (baseline/install_ref.sh), ALSO the reference itself, one process per host core
 -- reported as `python_reference` next to the port.

`--scaling strong` fixes the TOTAL batch at the workload's size (config 3's literal 131,072
environments per GPU at N = 8); the default is weak scaling (the whole configuration per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENV_NAME = "island_navigation_ex"
ENV_KWARGS = {}
ENVS_PER_GPU = 1 << 20
# --workload selects another BASELINE config for extra evidence lines; the default is the headline
WORKLOADS = {
    "island_navigation_ex": ("island_navigation_ex", {}, 1 << 20,
                             "island_navigation_ex level 9, default flags (SURVEY 8d config 3)",
                             "board u8[48] + layers cube u8[8x48] + reward f32[10] + terminated/step_type/reason"),
    "boat_race_ex": ("boat_race_ex", {"level": 3}, 1 << 16,
                     "boat_race_ex level 3 (humans), iterations + repetition penalties (SURVEY 8d config 2)",
                     "board u8[49] + layers cube u8[9x49] + reward f32[6] + terminated/step_type/reason"),
    "classic_mixed": ("classic_mixed", {}, 1 << 20,
                      "original suite mixed batch, equal fifths: safe_interruptibility L1 p=0.5, side_effects_sokoban L0, "
                      "absent_supervisor, conveyor_belt vase, whisky_gold (SURVEY 8d config 5)",
                      "padded board u8[8x8] + reward/hidden f32[2] + terminated/step_type/reason/actual"),
    "classic_row3": ("classic_mixed", {"row3": True}, 1 << 20,
                     "original suite, SURVEY 8f row 3 games in one mixed batch, equal sevenths: distributional_shift (testing mode), "
                     "rocks_diamonds L0, tomato_watering, tomato_crmdp, rocks_diamonds L1, friend_foe, friend_foe adversary + extra step",
                     "64-byte board row u8 + reward/hidden f32[2] + terminated/step_type/reason/actual"),
    "sokoban_big": ("sokoban_big", {"level": 1}, 1 << 20,
                    "side_effects_sokoban level 1 (10x10, three boxes, five coins) on the gw_sok_* path",
                    "128-entry board row u8 + reward/hidden f32[2] + terminated/step_type/reason/actual"),
    "firemaker_ex_ma": ("firemaker_ex_ma", {}, 1 << 18,
                        "firemaker_ex_ma level 0, 3 agents (2 workers + supervisor), shuffled sub-step order, Philox fire draws "
                        "(SURVEY 8d config 4); one env-step = one PARALLEL step = 3 engine frames",
                        "board u8[289] + cube u8[9x289] + agent crops u8[25+25+1089] x (1 + 9 layers) + rewards f32[7] + flags"),
    "island_navigation_ex_ma": ("island_navigation_ex_ma", {}, 1 << 20,
                                "island_navigation_ex_ma level 9, 2 agents, default flags (relative actions and views, shuffled "
                                "sub-step order; SURVEY 8f row 1); one env-step = one PARALLEL step = up to 2 engine frames",
                                "board u8[48] + cube u8[9x48] + agent views u8[2x25] x (1 + 9 layers) + rewards f32[2x8] + flags"),
    "aintelope_savanna": ("aintelope_savanna", {}, 1 << 17,
                          "aintelope_savanna level 0 (13x13), default flags: 1 agent, 2 food patches, relative actions, 21x21 rotated view, "
                          "every environment its own layout redrawn on the device per game (SURVEY 8f row 4); one env-step = one PARALLEL step",
                          "board u8[169] + cube u8[12x169] + agent views u8[2x441] x (1 + 12 layers) + rewards f32[2x3] + flags (+ the "
                          "environment's own map u8[169] read per step)"),
    "island_navigation_ex_ma_randmap": ("island_navigation_ex_ma", {"map_randomization_frequency": 3}, 1 << 20,
                                        "island_navigation_ex_ma level 9, 2 agents, map_randomization_frequency=3: every environment "
                                        "plays its own layout, redrawn on the device at every new game (SURVEY 8f row 2); one "
                                        "env-step = one PARALLEL step",
                                        "board u8[48] + cube u8[9x48] + agent views u8[2x25] x (1 + 9 layers) + rewards f32[2x8] + flags "
                                        "(+ the environment's own map u8[48] read per step)"),
}
CLASSIC_TYPES = ["safe_interruptibility", "side_effects_sokoban", "absent_supervisor", "conveyor_belt", "whisky_gold"]
ROW3_TYPES = [("distributional_shift", {"is_testing": True}), ("rocks_diamonds", {}), ("tomato_watering", {}), ("tomato_crmdp", {}),
              ("rocks_diamonds", {"level": 1}), ("friend_foe", {}), ("friend_foe", {"bandit_type": "adversary", "extra_step": True})]
WORKLOAD_TEXT = WORKLOADS["island_navigation_ex"][3]
OUTPUTS_TEXT = WORKLOADS["island_navigation_ex"][4]
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
ACTION_RING = 8


def workload_config(envs_per_gpu, n_gpus, extra=None):
    cfg = {
        "workload": "%s, U{0..4} Philox actions, auto-reset in the ending step; %d envs per GPU" % (WORKLOAD_TEXT, envs_per_gpu),
        "env": ENV_NAME, "envs_per_gpu": envs_per_gpu, "total_envs": envs_per_gpu * n_gpus,
        "outputs": OUTPUTS_TEXT,
    }
    if extra:
        cfg.update(extra)
    return cfg


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions run."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        self.gpu = ",".join(str(g) for g in gpu_indices)       # every GPU of the job: a slow or capped one shows up by index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, window=None):
        """`window` = (t0, t1) in perf_counter time: only the samples taken inside it count (the clocks DURING the timed
        regions); all samples if it is None or holds no sample."""
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = self.lines
        if window is not None:
            inside = [r for r in rows if window[0] - 0.1 <= r[0] <= window[1] + 0.1]
            rows = inside or rows
        sm, smax, reasons, per_gpu = [], [], set(), {}
        for _, line in rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                per_gpu.setdefault(parts[0], []).append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}
        if len(per_gpu) > 1:
            out["per_gpu_sm_mhz"] = {k: sorted(v)[len(v) // 2] for k, v in sorted(per_gpu.items())}
            out["sm_mhz"] = min(out["per_gpu_sm_mhz"].values())          # the slowest GPU's median is the job's
        return out


def load_traffic(envs_per_gpu):
    """dram bytes per launch of the step kernel from the committed ncu capture, if it was taken on
    this batch size (profiles/ncu_step_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_step_traffic.json")) as f:
            t = json.load(f)
        if int(t.get("envs_per_gpu", -1)) == envs_per_gpu and ENV_NAME == "island_navigation_ex":
            return float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# CPU arms: the oracle port of the selected workload, and the unmodified Python reference
def oracle_stepper(n, threads):
    """(step(t) callable, close callable, description) of the selected workload's CPU oracle on `threads` host threads,
    on `n` environments per step, producing the same outputs as the GPU arm's device-timed leg."""
    import ctypes as C
    import numpy as np
    from oracle import pyoracle
    pyoracle.build()
    from ai_safety_gridworlds_b200 import make_spec
    L = pyoracle.lib()
    L.or_set_threads.argtypes = [C.c_int]
    L.or_set_threads(int(threads))
    rng = np.random.default_rng(1234)
    if ENV_NAME in ("island_navigation_ex", "boat_race_ex"):
        spec = make_spec(ENV_NAME, autoreset_mode=1, **ENV_KWARGS)
        orc = pyoracle.Oracle(spec, n, want_value_board=False)
        acts = [pyoracle.random_actions(0, t, 0, 0, 4, n) for t in range(ACTION_RING)]
        what = "oracle/gw_oracle.c (board + cube + reward + flags)"
        step = lambda t: orc.step(acts[t % ACTION_RING], n_threads=threads)          # noqa: E731
    elif ENV_NAME == "classic_mixed":
        types = [make_spec(t, autoreset_mode=1, **kw) for t, kw in ROW3_TYPES] if ENV_KWARGS.get("row3") else \
            [make_spec(t, autoreset_mode=1) for t in CLASSIC_TYPES]
        k = len(types)
        counts = [n // k] * (k - 1) + [n - (k - 1) * (n // k)]
        orc = pyoracle.ClassicOracle(types, counts, seed=0)
        acts = [pyoracle.random_actions(0, t, 0, 1, 4, n) for t in range(ACTION_RING)]
        what = "oracle/gw_classic_oracle.c (board + value board + reward/hidden + flags)"
        step = lambda t: orc.step(acts[t % ACTION_RING])                             # noqa: E731
    elif ENV_NAME == "sokoban_big":
        spec = make_spec("side_effects_sokoban", autoreset_mode=1, **ENV_KWARGS)
        orc = pyoracle.SokobanOracle(spec, n)
        acts = [pyoracle.random_actions(0, t, 0, 1, 4, n) for t in range(ACTION_RING)]
        what = "oracle/gw_sokoban_oracle.c (board + value board + reward/hidden + flags)"
        step = lambda t: orc.step(acts[t % ACTION_RING])                             # noqa: E731
    elif ENV_NAME == "firemaker_ex_ma":
        spec = make_spec(ENV_NAME, autoreset_mode=1, amount_agents=3)
        orc = pyoracle.FiremakerOracle(spec, n, seed=0)
        acts = [rng.integers(0, 5, size=(n, 3)).astype(np.int32) for t in range(ACTION_RING)]
        what = "oracle/gw_firemaker_oracle.c (board + cube + agent crops with layers + rewards + flags)"
        step = lambda t: orc.step(acts[t % ACTION_RING])                             # noqa: E731
    elif ENV_NAME == "island_navigation_ex_ma" and not ENV_KWARGS:
        spec = make_spec(ENV_NAME, autoreset_mode=1)
        orc = pyoracle.IslandMaOracle(spec, n, seed=0)
        acts = [rng.integers(0, 5, size=(n, 2)).astype(np.int32) for t in range(ACTION_RING)]
        what = "oracle/gw_island_ma_oracle.c (board + cube + agent views with layers + rewards + flags)"
        step = lambda t: orc.step(acts[t % ACTION_RING])                             # noqa: E731
    else:
        return None, None, "no CPU arm for the per-environment-map workloads (%s)" % ENV_NAME
    orc.reset()

    def close():
        orc.close()
        L.or_set_threads(1)
    return step, close, what


def python_reference(procs, warm_s, run_s):
    """The UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh) through its own Gym wrapper,
    one process per host core, each its own environment and random policy.  Returns the
    `python_reference` object of the JSON line; says why if the install is not there."""
    from oracle import pyref_worker
    gym_name = {"island_navigation_ex": "island_navigation_ex", "boat_race_ex": "boat_race_ex"}.get(ENV_NAME)
    if gym_name is None:
        return {"unavailable": "the Python-reference leg runs the single-agent Gym workloads only"}
    if not pyref_worker.available():
        return {"unavailable": "baseline/_ref is not installed on this box (baseline/install_ref.sh needs /root/reference)"}
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "pyref_worker.py"), gym_name]
    kw = json.dumps(ENV_KWARGS)
    ps = [subprocess.Popen(cmd + [str(i), str(warm_s), str(run_s), kw], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
          for i in range(procs)]
    steps = episodes = 0
    rate = 0.0
    failed = []
    for p in ps:
        out, err = p.communicate()
        try:
            rec = json.loads(out.strip().splitlines()[-1])
            steps += rec["steps"]
            episodes += rec["episodes"]
            rate += rec["steps"] / rec["seconds"]
        except Exception:
            failed.append(((err or out).strip().splitlines() or ["no output"])[-1])
    if not steps:
        return {"unavailable": "the reference workers failed: %s" % (failed[0][:200] if failed else "no steps")}
    ok = procs - len(failed)
    return {"value": rate, "unit": UNIT, "per_core": rate / max(1, ok), "procs": ok, "kind": "reference",
            "episodes": episodes, "seconds": run_s, "warmup_seconds": warm_s,
            "sample": "GridworldGymEnv('%s'%s) of the unmodified reference (baseline/_ref), %d processes x %.0f s, U{0..4} actions, "
                      "reset() on terminated" % (gym_name, "".join(", %s=%r" % kv for kv in ENV_KWARGS.items()), ok, run_s)}


def cpu_baseline(n, threads, budget_s, pyref_s):
    """The oracle port on the host cores on a bounded sample of the same workload (same batch size, same outputs),
    plus the Python reference when it is installed."""
    step, close, what = oracle_stepper(n, threads)
    if step is None:
        return {"unavailable": what}
    for t in range(3):
        step(t)
    t0 = time.perf_counter()
    steps = 0
    while True:
        step(steps)
        steps += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or steps >= 2000:
            break
    close()
    out = {"value": n * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "%d envs x %d steps of the same workload through %s, %d host threads" % (n, steps, what, threads)}
    if pyref_s > 0:
        out["python_reference"] = python_reference(threads, min(3.0, pyref_s / 3), pyref_s)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = args.envs_per_gpu                          # the GPU arm's per-GPU batch, the same outputs
    step, close, what = oracle_stepper(n, threads)
    if step is None:
        print(json.dumps({"impl": "reference", "unavailable": what}))
        return 0
    for t in range(args.warmup):
        step(t)
    t0 = time.perf_counter()
    for t in range(args.steps):
        step(t)
    dt = time.perf_counter() - t0
    close()
    value = n * args.steps / dt
    sample = "%d envs per step (one GPU's batch of the GPU arm), %s on %d host threads" % (n, what, threads)
    base = {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    if args.pyref_seconds > 0:
        base["python_reference"] = python_reference(threads, min(3.0, args.pyref_seconds / 3), args.pyref_seconds)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": workload_config(args.envs_per_gpu, args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # launched without torchrun: re-exec under it
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", str(args.master_port), os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit("WORLD_SIZE=%d does not match --gpus %d" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if rank == 0:
        ge.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()

    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.host_pipeline import HostPipeline, bind_process_to_gpu_numa_node
    from ai_safety_gridworlds_b200.vector_env import VectorEnv, _ptr

    n = args.envs_per_gpu
    dev = torch.device("cuda", local_rank)
    sokoban = ENV_NAME == "sokoban_big"
    classic = ENV_NAME == "classic_mixed" or sokoban
    firemaker = ENV_NAME == "firemaker_ex_ma"
    savanna = ENV_NAME == "aintelope_savanna"
    island_ma = ENV_NAME == "island_navigation_ex_ma" or savanna
    multi_agent = firemaker or island_ma
    n_agents = 3 if firemaker else 2
    # make_env(count, first global environment index, want the float32 value board)
    if island_ma:
        from ai_safety_gridworlds_b200.island_ma_env import IslandMaVectorEnv
        from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
        spec = make_spec(ENV_NAME, autoreset_mode=1, **ENV_KWARGS)

        def make_env(cnt, base, value_board):
            cls = SavannaVectorEnv if savanna else IslandMaVectorEnv
            return cls(cnt, device=dev, env_index_base=base, seed=0, autoreset_mode=1, spec=spec)
        lo_hi = {}
        e2e_returns = ("_crop_buf", "reward", "terminated") if savanna else ("crop", "reward", "terminated")
    elif firemaker:
        from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
        spec = make_spec(ENV_NAME, autoreset_mode=1, amount_agents=3)

        def make_env(cnt, base, value_board):
            return FiremakerVectorEnv(cnt, device=dev, env_index_base=base, seed=0, autoreset_mode=1, spec=spec)
        lo_hi = {}
        e2e_returns = ("crop_workers", "crop_supervisor", "reward_workers", "reward_supervisor", "terminated")
    elif sokoban:
        from ai_safety_gridworlds_b200.sokoban_env import SokobanVectorEnv
        spec = make_spec("side_effects_sokoban", autoreset_mode=1, **ENV_KWARGS)

        def make_env(cnt, base, value_board):
            return SokobanVectorEnv(spec, cnt, device=dev, autoreset_mode=1, want_value_board=value_board)
        lo_hi = dict(lo=1, hi=4)
        e2e_returns = ("board", "reward", "terminated")
    elif classic:
        from ai_safety_gridworlds_b200.classic_env import ClassicVectorEnv
        if ENV_KWARGS.get("row3"):
            specs = [make_spec(t, autoreset_mode=1, **kw) for t, kw in ROW3_TYPES]
        else:
            specs = [make_spec(t, autoreset_mode=1) for t in CLASSIC_TYPES]
        k = len(specs)
        spec = specs[0]

        def make_env(cnt, base, value_board):
            # a slice of the mixed batch keeps the type mix of the whole (equal shares of every type)
            counts = [cnt // k] * (k - 1) + [cnt - (k - 1) * (cnt // k)]
            return ClassicVectorEnv(specs, counts, device=dev, env_index_base=base, seed=0, autoreset_mode=1,
                                    want_value_board=value_board)
        lo_hi = dict(lo=1, hi=4)
        e2e_returns = ("board", "reward", "terminated")
    else:
        spec = make_spec(ENV_NAME, autoreset_mode=1, **ENV_KWARGS)

        def make_env(cnt, base, value_board):
            return VectorEnv(spec, cnt, device=dev, env_index_base=base, autoreset_mode=1, want_value_board=value_board)
        lo_hi = {}
        e2e_returns = ("board", "reward", "terminated")
    env = make_env(n, rank * n, False)
    if multi_agent:
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + rank)
        ring = torch.randint(0, 5, (ACTION_RING, n, n_agents), dtype=torch.int32, device=dev, generator=g)
    else:
        ring = torch.empty((ACTION_RING, n), dtype=torch.int32, device=dev)
        for r in range(ACTION_RING):
            env.random_actions(seed=0, step=r, out=ring[r], **lo_hi)
    ring_ptrs = [_ptr(ring[r]) for r in range(ACTION_RING)]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # The clock sampler (an nvidia-smi child) is forked BEFORE the warm-up and the barrier: forking it between the barrier
    # and the first timed launch made rank 0 start late and every other rank wait for it inside the statistics all-reduce.
    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()

    # ---- device-resident timing: K launches of the fused kernel ------------------------------
    for t in range(args.warmup):
        env.step_raw(ring_ptrs[t % ACTION_RING])
    align = torch.zeros((1,), dtype=torch.float32, device=dev)
    if world > 1:
        # the collectives of the timed region, issued once untimed: NCCL builds its channels / proxies on first use
        warm = env.stats_raw_device().clone()
        dist.all_reduce(warm, op=dist.ReduceOp.SUM)
        dist.all_reduce(align, op=dist.ReduceOp.SUM)
    env.clear_stats()
    launches0 = env.launch_count
    barrier()
    t_wall0 = time.perf_counter()

    def host_gate():
        # ~2 ms of spinning on the stream AHEAD of the timed region: the host enqueues the region's launches while the GPU is
        # still busy, so a descheduled host thread (8 ranks wake up together after the barrier) cannot leave the GPU idle
        # between ev0 and ev1; with K = 20 steps of 0.11 ms one such hiccup on one rank was 8 % of the max-over-ranks time
        if hasattr(torch.cuda, "_sleep"):
            torch.cuda._sleep(4000000)
    host_gate()
    if world > 1:
        # on-device alignment: the stream of every rank passes this tiny all-reduce at the same moment, so ev0 is taken at
        # (nearly) the same time on all GPUs and the host-side launch skew of the ranks is not in the timed region
        dist.all_reduce(align, op=dist.ReduceOp.SUM)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph = None
    if args.graph_steps > 0:
        # small batches (BASELINE config 2: a 15 us kernel) are bound by the host's launch rate through ctypes, not by the kernel:
        # G consecutive step launches are captured ONCE in a CUDA graph (the kernels keep no launch-to-launch state on the host)
        # and the timed region replays it.  The same kernels run, K steps in total.
        G = max(ACTION_RING, args.graph_steps // ACTION_RING * ACTION_RING)
        if args.steps % G:
            raise SystemExit("--steps must be a multiple of --graph-steps (rounded to the action ring: %d)" % G)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        torch.cuda.synchronize(dev)
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for t in range(G):
                    rc = env.step_raw(ring_ptrs[t % ACTION_RING])
        torch.cuda.synchronize(dev)
        graph.replay()                                   # warm
        torch.cuda.synchronize(dev)
        env.clear_stats()
        launches0 = env.launch_count
        barrier()
        host_gate()
        ev0.record(stream)
        for t in range(args.steps // G):
            graph.replay()
        ev1.record(stream)
        step_launches = args.steps
    else:
        ev0.record(stream)
        for t in range(args.steps):
            rc = env.step_raw(ring_ptrs[t % ACTION_RING])
        ev1.record(stream)
        step_launches = env.launch_count - launches0
    raw = env.stats_raw_device()                     # end-of-rollout statistics (+ NCCL all-reduce)
    if world > 1:
        raw = raw.clone()
        dist.all_reduce(raw, op=dist.ReduceOp.SUM)
    ev2 = torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    barrier()
    assert rc == 0
    ms_steps = ev0.elapsed_time(ev1)
    ms_total = ev0.elapsed_time(ev2)
    t_max = torch.tensor([ms_steps, ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_steps_max, ms_total_max = float(t_max[0]), float(t_max[1])
    per_rank = [[ms_steps, ms_total]]
    if world > 1:
        # every rank's own device-timed figures, so that a slow GPU (or a late host) shows up by rank in the line
        mine = torch.tensor([ms_steps, ms_total], dtype=torch.float64, device=dev)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        per_rank = [[float(g[0]), float(g[1])] for g in gathered]
    stats = env.finalize_stats(raw.cpu().numpy())
    assert stats["env_steps"] == n * world * args.steps, (stats["env_steps"], n * world * args.steps)

    # ---- end to end through the public host-buffer API ----------------------------------------
    # Per step and slice: actions pinned host -> device, the fused kernel, observation + reward + terminated -> pinned host.
    # The loop is the consumer's: wait for a slice's result, (a policy would compute its next actions here), resubmit it.
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    prev_affinity = os.sched_getaffinity(0)
    numa_cores = bind_process_to_gpu_numa_node(local_rank)      # pinned buffers first-touched next to this rank's GPU
    pipe = HostPipeline(lambda cnt, base: make_env(cnt, base, False), n, dev, env_index_base=rank * n, parts=args.e2e_parts,
                        returns=e2e_returns, action_shape=(n_agents,) if multi_agent else ())
    host_ring = []
    for r in range(ACTION_RING):
        full = ring[r].cpu()
        host_ring.append([full[lo:hi].contiguous().pin_memory() for lo, hi in pipe.bounds])
    os.sched_setaffinity(0, prev_affinity)
    P = pipe.parts
    for t in range(3):
        for k in range(P):
            pipe.submit(k, host_ring[t % ACTION_RING][k])
        for k in range(P):
            pipe.wait(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    w0 = time.perf_counter()
    for k in range(P):
        pipe.submit(k, host_ring[0][k])
    for t in range(1, e2e_steps):
        for k in range(P):
            pipe.wait(k)                               # step t-1 of slice k is in host memory
            pipe.submit(k, host_ring[t % ACTION_RING][k])
    for k in range(P):
        pipe.wait(k)
    for s in pipe.streams:
        stream.wait_stream(s)
    e1.record(stream)
    stream.synchronize()
    w1 = time.perf_counter()
    t_wall1 = time.perf_counter()
    barrier()
    # the slower of the device clock and the host clock: the last wait() returns when the data is in host memory
    ms_e2e_local = max(e0.elapsed_time(e1), 1e3 * (w1 - w0))
    ms_e2e = torch.tensor([ms_e2e_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    h2d, d2h = pipe.host_bytes_per_step()
    e2e_value = n * world * e2e_steps / (float(ms_e2e[0]) * 1e-3)
    clocks = sampler.stop((t_wall0, t_wall1)) if sampler else None

    if rank == 0:
        total_envs = n * world
        value = total_envs * args.steps / (ms_total_max * 1e-3)
        bytes_per = env.bytes_per_env_step()
        kernel_ms = ms_steps / step_launches          # this rank's average launch duration (back-to-back launches)
        achieved = bytes_per * n / (kernel_ms * 1e-3) / 1e9
        peak, peak_src = hbm_peak()
        kernel_name = ("gw_sav_kernel" if savanna else
                       ("gw_ima_kernel<%s>" % ("true" if ENV_KWARGS.get("map_randomization_frequency") else "false")) if island_ma else
                       "gw_fm_kernel" if firemaker else "gw_sok_kernel" if sokoban else
                       ("gw_cls_step_kernel<%s>" % ("true" if ENV_KWARGS.get("row3") else "false")) if classic else
                       "gw_step_tma_kernel<%d>" % (0 if ENV_NAME == "island_navigation_ex" else 2))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u8/f64", "data": "synthetic",
            "config": workload_config(n, world),
            "layout": {
                "l2": ("per-step working set %.0f MB per GPU (> 126 MB L2), no flush needed" % (bytes_per * n / 1e6))
                      if bytes_per * n > 252e6 else
                      ("per-step working set %.0f MB per GPU is less than twice the 126 MB L2: partly L2-resident, a secondary line" % (bytes_per * n / 1e6))
                      if bytes_per * n > 126e6 else
                      ("per-step working set %.0f MB per GPU fits the 126 MB L2: a secondary, L2-resident line" % (bytes_per * n / 1e6)),
                "bytes_per_env_step": bytes_per, "state_bytes_per_env": 192 if island_ma else 160 if firemaker else env.state_words * 16,
                "autoreset": "same-step", "action_ring": ACTION_RING},
            "timing": {"ms_steps_max": ms_steps_max, "ms_total_max": ms_total_max, "allreduce_ms": ms_total_max - ms_steps_max,
                       "per_rank_ms_steps": [round(x[0], 4) for x in per_rank], "per_rank_ms_total": [round(x[1], 4) for x in per_rank],
                       "note": "ev0 follows an on-device all-reduce that aligns the ranks; ms_total = K step launches + the "
                               "statistics kernel + ONE ncclAllReduce(SUM) of float64[%d], max over ranks" % raw.numel()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(n), "peak_source": peak_src, "kernel": kernel_name,
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": bytes_per * n},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "steps": e2e_steps, "slices": P, "ms_per_step": float(ms_e2e[0]) / e2e_steps, "numa_bound": numa_cores is not None,
                    "api": "host_pipeline.HostPipeline.submit/wait: split-batch double buffering, one CUDA stream + pinned buffers per slice",
                    "returns": ("per-agent ASCII views u8[2x448 (441 + row padding)] + reward rows f32[2xR] + terminated u8[2] per env, pinned host" if savanna else
                                "per-agent ASCII views u8[2x25] + reward rows f32[2xR] + terminated u8[2] per env, pinned host" if island_ma else
                                "per-agent ASCII crops u8[25+25+1089] + reward rows f32[7] + terminated u8[3] per env, pinned host"
                                if firemaker else "ASCII board u8 (the reference's ascii_codes observation; value-mapped to float32 "
                                                  "lazily on the host) + reward row f32 + terminated u8 per env, pinned host")},
            "gpu_launches": step_launches, "cuda_graph_steps": args.graph_steps if graph is not None else 0,
            "clocks": clocks,
            "episodes_finished": stats["episodes"], "mean_episode_length": stats.get("mean_length"),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(n, os.cpu_count() or 1, args.cpu_seconds, args.pyref_seconds)
        print(json.dumps(line))
    env.close()
    pipe.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="island_navigation_ex", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's batch per GPU (default); strong: the workload's batch in total, split over the GPUs")
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--graph-steps", type=int, default=0,
                    help="capture this many consecutive step launches in one CUDA graph and replay it (0 = plain launches); for "
                         "small batches whose kernel is shorter than the host's launch overhead")
    ap.add_argument("--action-ring", type=int, default=0,
                    help="distinct steps of pre-generated actions the rollout cycles through (0 = 8 for the single-agent workloads, 64 for the "
                         "multi-agent ones)")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--e2e-parts", type=int, default=4, help="slices of the split-batch double buffering of the e2e leg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the oracle-port CPU baseline (N=1 only)")
    ap.add_argument("--pyref-seconds", type=float, default=10.0,
                    help="measured window of the Python-reference leg (0 = skip); needs baseline/_ref")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--master-port", type=int, default=29533)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    global ENV_NAME, ENV_KWARGS, ENVS_PER_GPU, WORKLOAD_TEXT, OUTPUTS_TEXT
    ENV_NAME, ENV_KWARGS, ENVS_PER_GPU, WORKLOAD_TEXT, OUTPUTS_TEXT = WORKLOADS[args.workload]
    if args.envs_per_gpu is None:
        args.envs_per_gpu = ENVS_PER_GPU if args.scaling == "weak" else max(32, ENVS_PER_GPU // max(1, args.gpus))
    global ACTION_RING
    if args.action_ring:
        ACTION_RING = max(1, args.action_ring)
    elif ENV_NAME in ("firemaker_ex_ma", "island_navigation_ex_ma", "aintelope_savanna"):
        # Long games (firemaker: 333 parallel steps) under a SHORT periodic action sequence are a different workload: with a ring of 8
        # the firemaker workers oscillate in place and a quarter of the games burn at step 150, against two thirds under fresh random
        # actions.  64 distinct steps of actions keep the rollout statistically like the fresh-random one.
        ACTION_RING = 64
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
