#!/usr/bin/env python
"""Records the CSV logs the UNMODIFIED reference writes (safety_game_mo.py:727-807,1110-1215) while it replays the action
sequence of an existing golden trace, into tests/golden/logs/<trace>.csv (+ the separate arguments file).  TEST INFRASTRUCTURE ONLY.

The timestamp column is left out (it is the wall clock); every other column type is logged.  tiletype_qvalue is logged without
q-values (the caller never passes q_value_per_action): the reference then writes zeros per tile type.

Usage: python oracle/record_csv_log.py [TRACE ...]
"""
import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"
CASES = ["island_ex_default_s0", "island_ex_fractional_s19", "boat_ex_level3_s0", "island_exp_food_drink_bounded_death_gold_silver"]
LOG_COLUMNS = ["env", "trial", "env layout seed", "episode", "iteration", "reward", "scalar_reward", "cumulative_reward",
               "average_reward", "scalar_cumulative_reward", "scalar_average_reward", "gini_index", "cumulative_gini_index",
               "mo_variance", "cumulative_mo_variance", "average_mo_variance", "metric", "tiletype_qvalue"]
STEPS = 150
# explicit reset() calls after that many steps: they advance the episode column (safety_game_mo.py:697-705), which the restart
# made inside step() after a terminal timestep does not
EXPLICIT_RESETS = {"boat_ex_level3_s0": [50, 100, 101]}


def _worker(name):
    import numpy as np
    sys.path.insert(0, HERE)
    import shims  # noqa: F401
    from ai_safety_gridworlds.helpers.gridworld_gym_env import GridworldGymEnv
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(d["meta_json"]))
    tmp = tempfile.mkdtemp()
    env = GridworldGymEnv(meta["env"], seed=meta["seed"], log_columns=list(LOG_COLUMNS), log_dir=tmp,
                          log_filename_comment="golden", log_arguments_to_separate_file=True, **meta["kwargs"])
    # the reference opens the log file in a reset() call that finds the environment at a FIRST timestep (safety_game_mo.py:577-583);
    # the constructor leaves _state = None, so that is the SECOND reset() -- the way the environments' own main() calls it
    env.reset()
    env.reset()
    resets = EXPLICIT_RESETS.get(name, [])
    for k, a in enumerate(d["actions"][:STEPS]):
        for _ in range(resets.count(k)):
            env.reset()
        env.step(int(a))
    env.close() if hasattr(env, "close") else None
    handle = getattr(type(env._env), "log_file_handle", None)
    if handle:
        handle.flush()
    out = os.path.join(GOLDEN, "logs")
    os.makedirs(out, exist_ok=True)
    csvs = [p for p in glob.glob(os.path.join(tmp, "*.csv"))]
    txts = [p for p in glob.glob(os.path.join(tmp, "*arguments*.txt"))]
    assert len(csvs) == 1 and len(txts) == 1, (csvs, txts)
    shutil.copy(csvs[0], os.path.join(out, name + ".csv"))
    shutil.copy(txts[0], os.path.join(out, name + ".arguments.txt"))
    with open(os.path.join(out, name + ".meta.json"), "w") as f:
        json.dump(dict(trace=name, steps=STEPS, explicit_resets=resets, log_columns=LOG_COLUMNS, log_filename_comment="golden",
                       reference_filename=os.path.basename(csvs[0]), recorder="oracle/record_csv_log.py"), f, indent=1)
    print("%-50s rows=%d" % (name, sum(1 for _ in open(csvs[0]))))


def main(argv):
    if len(argv) >= 2 and argv[0] == "--worker":
        _worker(argv[1])
        return 0
    if not os.path.isdir(REFERENCE):
        print("reference not mounted at %s" % REFERENCE)
        return 1
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "stubs"), REFERENCE])
    rc = 0
    for name in (argv or CASES):
        rc |= subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", name], env=env, cwd=tempfile.gettempdir()).returncode
    return rc


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
