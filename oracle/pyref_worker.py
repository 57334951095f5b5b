#!/usr/bin/env python
"""One worker of the Python-reference CPU baseline.  TEST / BENCH INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference -- the `ai_safety_gridworlds` + `pycolab` packages pip-installed from
/root/reference into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot) -- through
its own public boundary `GridworldGymEnv(env).reset/step` (helpers/gridworld_gym_env.py:455,588) under a
uniform random policy with reset() on `terminated`, for a fixed wall-clock window, and prints one JSON
line {"steps": ..., "episodes": ..., "seconds": ...}.  bench.py starts one worker per host core and
sums the lines.  The `gymnasium` / `pettingzoo` stubs of oracle/stubs/ stand in for the packages the
image does not have; oracle/shims.py restores `np.Inf` for numpy 2.

    python oracle/pyref_worker.py ENV SEED WARMUP_S MEASURE_S [kwargs-json]
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "ai_safety_gridworlds")) and os.path.isdir(os.path.join(REF, "pycolab"))


def main(argv):
    env_name, seed, warm_s, run_s = argv[1], int(argv[2]), float(argv[3]), float(argv[4])
    kwargs = json.loads(argv[5]) if len(argv) > 5 else {}
    if not available():
        print(json.dumps({"unavailable": "baseline/_ref is not installed"}))
        return 0
    sys.path[:0] = [REF, os.path.join(HERE, "stubs"), HERE]
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np
    import shims  # noqa: F401
    from ai_safety_gridworlds.helpers.gridworld_gym_env import GridworldGymEnv
    env = GridworldGymEnv(env_name, seed=seed, **kwargs)
    rng = np.random.default_rng(seed)
    env.reset()
    lo = int(getattr(env.action_space, "start", 0))
    n_act = min(int(env.action_space.n), 5)            # U{0..4}: NOOP + the four moves
    steps = episodes = 0
    t_end_warm = time.perf_counter() + warm_s
    t0 = None
    while True:
        obs, reward, terminated, truncated, info = env.step(lo + int(rng.integers(0, n_act)))
        if terminated or truncated:
            env.reset()
            if t0 is not None:
                episodes += 1
        now = time.perf_counter()
        if t0 is None:
            if now >= t_end_warm:
                t0 = now
            continue
        steps += 1
        if now - t0 >= run_s:
            break
    print(json.dumps({"steps": steps, "episodes": episodes, "seconds": time.perf_counter() - t0}))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
