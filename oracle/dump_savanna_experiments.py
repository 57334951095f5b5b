#!/usr/bin/env python
"""Reads the flag overlays of the reference's aintelope_savanna experiments (ai_safety_gridworlds/experiments/aintelope/*.py:
`init_experiment_flags()`) by importing them, and prints the values that differ from the environment's own defaults as Python
literals.  TEST / BUILD INFRASTRUCTURE: the output was pasted into ai_safety_gridworlds_b200/envs/savanna_experiments.py.
Run with PYTHONPATH=oracle/stubs:/root/reference, one experiment per fresh interpreter (absl flags are process globals)."""
import importlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["danger_tiles", "food_drink_homeostasis", "food_drink_homeostasis_danger_gold_silver", "food_drink_homeostasis_gold",
         "food_drink_homeostasis_gold_silver", "food_drink_homeostasis_predators_gold_silver", "food_homeostasis", "food_sharing",
         "food_sustainability", "food_unbounded", "predators", "savanna_demo"]


def worker(name):
    sys.path.insert(0, HERE)
    import shims  # noqa: F401
    from absl import flags as _absl_flags  # noqa: F401
    sys.path.insert(0, os.path.dirname(HERE))
    from ai_safety_gridworlds_b200.envs.aintelope_savanna import DEFAULT_FLAGS
    mod = importlib.import_module("ai_safety_gridworlds.experiments.aintelope." + name)
    flags = mod.init_experiment_flags()
    out = {}
    for key, default in DEFAULT_FLAGS.items():
        v = getattr(flags, key)
        if hasattr(v, "_reward_dimensions_dict"):
            v = dict(v._reward_dimensions_dict)
        if isinstance(default, bool):
            v = bool(v)
        elif isinstance(default, float) and v is not None:
            v = float(v)
        if v != default:
            out[key] = v
    classes = [n for n in dir(mod) if n.endswith("Experiment")]
    print(json.dumps({"name": name, "overlay": out, "classes": classes}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
    else:
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(HERE, "stubs"), "/root/reference"]))
        for n in NAMES:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", n], env=env)
