#!/usr/bin/env python
"""bench.py -- env-steps/sec of the fused step kernel on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: island_navigation_ex level 9, default flags, uniform random
actions U{0..4} (Philox, keyed by global environment index), auto-reset in the step that ends an
episode.  1,048,576 environments PER GPU (the whole configuration on one GPU at N=1; weak scaling
above), which makes the per-step working set ~0.7 GB, several times the 126 MB L2.

A step = one launch of the fused kernel over the rank's whole batch.  `value` is device-timed with
inputs resident in HBM; `e2e` goes through VectorEnv.step_host with pinned HOST buffers (actions
H2D, observation + reward + terminated D2H inside the timed region).  `--impl reference` times the
CPU oracle (the C restatement of the reference's per-step path; the reference itself is pure
Python and cannot travel to the GPU box) on all host threads.
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
# measured in the build container while surveying: the reference's own
# Python path, GridworldGymEnv('island_navigation_ex'), random policy, one core
PY_REFERENCE_STEPS_PER_S_PER_CORE = 1.07e3


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
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


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


def cpu_baseline(spec, threads, budget_s=12.0):
    """The oracle port on the host cores, on a bounded sample of the same workload."""
    import numpy as np
    from oracle import pyoracle
    n = 1 << 17
    orc = pyoracle.Oracle(spec, n, want_value_board=False)
    orc.reset()
    acts = [pyoracle.random_actions(0, t, 0, 0, 4, n) for t in range(ACTION_RING)]
    for t in range(3):
        orc.step(acts[t % ACTION_RING], n_threads=threads)
    t0 = time.perf_counter()
    steps = 0
    while True:
        orc.step(acts[steps % ACTION_RING], n_threads=threads)
        steps += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or steps >= 2000:
            break
    orc.close()
    return {"value": n * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d envs x %d steps of the same workload through oracle/gw_oracle.c (board+cube+reward), %d pthreads"
                      % (n, steps, threads),
            "python_reference_steps_per_s_per_core": PY_REFERENCE_STEPS_PER_S_PER_CORE,
            "python_reference_note": "the unmodified Python reference measured in the build container (SURVEY.md section 6); "
                                     "it cannot travel to the GPU box"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if ENV_NAME in ("classic_mixed", "firemaker_ex_ma"):
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the headline workload only; "
                          "%s is a secondary evidence line" % ENV_NAME}))
        return 0
    import __graft_entry__ as ge
    from oracle import pyoracle
    pyoracle.build()
    from ai_safety_gridworlds_b200 import make_spec
    import numpy as np  # noqa: F401
    spec = make_spec(ENV_NAME, autoreset_mode=1, **ENV_KWARGS)
    threads = os.cpu_count() or 1
    n = 1 << 17                                   # bounded sample: 1/8 of one GPU's batch per step
    orc = pyoracle.Oracle(spec, n, want_value_board=False)
    orc.reset()
    acts = [pyoracle.random_actions(0, t, 0, 0, 4, n) for t in range(ACTION_RING)]
    for t in range(args.warmup):
        orc.step(acts[t % ACTION_RING], n_threads=threads)
    t0 = time.perf_counter()
    for t in range(args.steps):
        orc.step(acts[t % ACTION_RING], n_threads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = "%d envs per step (1/8 of one GPU's batch), oracle/gw_oracle.c on %d pthreads" % (n, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": workload_config(ENVS_PER_GPU, args.gpus, {"reference_sample_envs": n}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "python_reference_steps_per_s_per_core": PY_REFERENCE_STEPS_PER_S_PER_CORE},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


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
    from ai_safety_gridworlds_b200.vector_env import VectorEnv, _ptr

    n = args.envs_per_gpu
    dev = torch.device("cuda", local_rank)
    sokoban = ENV_NAME == "sokoban_big"
    classic = ENV_NAME == "classic_mixed" or sokoban
    firemaker = ENV_NAME == "firemaker_ex_ma"
    savanna = ENV_NAME == "aintelope_savanna"
    island_ma = ENV_NAME == "island_navigation_ex_ma" or savanna
    n_agents = 3 if firemaker else 2
    if island_ma:
        from ai_safety_gridworlds_b200.island_ma_env import IslandMaVectorEnv
        from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
        spec = make_spec(ENV_NAME, autoreset_mode=1, **ENV_KWARGS)

        def make_env(value_board):
            cls = SavannaVectorEnv if savanna else IslandMaVectorEnv
            return cls(n, device=dev, env_index_base=rank * n, seed=0, autoreset_mode=1, spec=spec)
        lo_hi = {}
        firemaker = True                             # from here on: "the multi-agent path" (per-agent action columns, no device statistics)
    elif firemaker:
        from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
        spec = make_spec(ENV_NAME, autoreset_mode=1, amount_agents=3)

        def make_env(value_board):
            return FiremakerVectorEnv(n, device=dev, env_index_base=rank * n, seed=0, autoreset_mode=1, spec=spec)
        lo_hi = {}
    elif sokoban:
        from ai_safety_gridworlds_b200.sokoban_env import SokobanVectorEnv
        spec = make_spec("side_effects_sokoban", autoreset_mode=1, **ENV_KWARGS)

        def make_env(value_board):
            return SokobanVectorEnv(spec, n, device=dev, autoreset_mode=1, want_value_board=value_board)
        lo_hi = dict(lo=1, hi=4)
    elif classic:
        from ai_safety_gridworlds_b200.classic_env import ClassicVectorEnv
        if ENV_KWARGS.get("row3"):
            specs = [make_spec(t, autoreset_mode=1, **kw) for t, kw in ROW3_TYPES]
        else:
            specs = [make_spec(t, autoreset_mode=1) for t in CLASSIC_TYPES]
        k = len(specs)
        counts = [n // k] * (k - 1) + [n - (k - 1) * (n // k)]
        spec = specs[0]

        def make_env(value_board):
            return ClassicVectorEnv(specs, counts, device=dev, env_index_base=rank * n, seed=0, autoreset_mode=1,
                                    want_value_board=value_board)
        lo_hi = dict(lo=1, hi=4)
    else:
        spec = make_spec(ENV_NAME, autoreset_mode=1, **ENV_KWARGS)

        def make_env(value_board):
            return VectorEnv(spec, n, device=dev, env_index_base=rank * n, autoreset_mode=1, want_value_board=value_board)
        lo_hi = {}
    env = make_env(False)
    if firemaker:
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

    # ---- device-resident timing: K launches of the fused kernel ------------------------------
    for t in range(args.warmup):
        env.step_raw(ring_ptrs[t % ACTION_RING])
    env.clear_stats()
    launches0 = env.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
    clocks = sampler.stop() if sampler else None
    ms_steps = ev0.elapsed_time(ev1)
    ms_total = ev0.elapsed_time(ev2)
    t_max = torch.tensor([ms_steps, ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_steps_max, ms_total_max = float(t_max[0]), float(t_max[1])
    stats = env.finalize_stats(raw.cpu().numpy())
    assert stats["env_steps"] == n * world * args.steps, (stats["env_steps"], n * world * args.steps)

    # ---- end to end through the public API with host buffers ---------------------------------
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    host_ring = [ring[r].cpu().pin_memory() for r in range(ACTION_RING)]
    if firemaker:
        env_e = env
        pin = dict(pin_memory=True)
        d_tensors = (env.crop, env.reward, env.terminated) if island_ma else \
            (env.crop_workers, env.crop_supervisor, env.reward_workers, env.reward_supervisor, env.terminated)
        h_tensors = [torch.zeros_like(t, device="cpu", **pin) for t in d_tensors]
        d_act = torch.zeros((n, n_agents), dtype=torch.int32, device=dev)

        def step_host(a_host):
            d_act.copy_(a_host, non_blocking=True)
            env.step_raw(_ptr(d_act))
            for h, t in zip(h_tensors, d_tensors):
                h.copy_(t, non_blocking=True)
            stream.synchronize()
            return h_tensors, None, None
        fm_bytes = (n * 4 * n_agents, sum(t.numel() * t.element_size() for t in d_tensors))
    else:
        env_e = make_env(True)
        step_host = env_e.step_host
    for t in range(3):
        step_host(host_ring[t % ACTION_RING])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(e2e_steps):
        obs_h, rew_h, term_h = step_host(host_ring[t % ACTION_RING])
    e1.record(stream)
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    h2d, d2h = fm_bytes if firemaker else env_e.host_bytes_per_step()
    e2e_value = n * world * e2e_steps / (float(ms_e2e[0]) * 1e-3)

    if rank == 0:
        total_envs = n * world
        value = total_envs * args.steps / (ms_total_max * 1e-3)
        bytes_per = env.bytes_per_env_step()
        kernel_ms = ms_steps / step_launches          # this rank's average launch duration (back-to-back launches)
        achieved = bytes_per * n / (kernel_ms * 1e-3) / 1e9
        peak, peak_src = hbm_peak()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f64", "data": "synthetic",
            "config": workload_config(n, world, {
                "l2": ("per-step working set %.0f MB per GPU (> 126 MB L2), no flush needed" % (bytes_per * n / 1e6))
                      if bytes_per * n > 252e6 else
                      ("per-step working set %.0f MB per GPU is less than twice the 126 MB L2: partly L2-resident, a secondary line" % (bytes_per * n / 1e6))
                      if bytes_per * n > 126e6 else
                      ("per-step working set %.0f MB per GPU fits the 126 MB L2: a secondary, L2-resident line" % (bytes_per * n / 1e6)),
                "bytes_per_env_step": bytes_per, "state_bytes_per_env": 192 if island_ma else 160 if firemaker else env.state_words * 16,
                "autoreset": "same-step", "action_ring": ACTION_RING}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(n), "peak_source": peak_src, "kernel": "gw_sav_kernel" if savanna else ("gw_ima_kernel<%s>" % ("true" if ENV_KWARGS.get("map_randomization_frequency") else "false")) if island_ma else "gw_fm_kernel" if firemaker else "gw_sok_kernel" if sokoban else ("gw_cls_step_kernel<%s>" % ("true" if ENV_KWARGS.get("row3") else "false")) if classic else "gw_step_tma_kernel<%d>" % (0 if ENV_NAME == "island_navigation_ex" else 2),
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": bytes_per * n},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "steps": e2e_steps,
                    "returns": ("per-agent ASCII views u8[2x441] + reward rows f32[2xR] + terminated u8[2] per env, pinned host" if savanna else
                                "per-agent ASCII views u8[2x25] + reward rows f32[2xR] + terminated u8[2] per env, pinned host" if island_ma else
                                "per-agent ASCII crops u8[25+25+1089] + reward rows f32[7] + terminated u8[3] per env, pinned host"
                                if firemaker else "value-mapped board f32 + reward row f32 + terminated u8 per env, pinned host")},
            "gpu_launches": step_launches,
            "clocks": clocks,
            "episodes_finished": stats["episodes"], "mean_episode_length": stats.get("mean_length"),
        }
        if world == 1 and not args.no_cpu_baseline and not classic and not firemaker:
            line["cpu_baseline"] = cpu_baseline(spec, os.cpu_count() or 1)
        print(json.dumps(line))
    env.close()
    if env_e is not env:
        env_e.close()
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
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--master-port", type=int, default=29533)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    global ENV_NAME, ENV_KWARGS, ENVS_PER_GPU, WORKLOAD_TEXT, OUTPUTS_TEXT
    ENV_NAME, ENV_KWARGS, ENVS_PER_GPU, WORKLOAD_TEXT, OUTPUTS_TEXT = WORKLOADS[args.workload]
    if args.envs_per_gpu is None:
        args.envs_per_gpu = ENVS_PER_GPU
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
