#!/usr/bin/env python
"""Reads the colour tables of the reference's environments -- the `colour_mapping` its observation distiller paints `RGB`
with (environments/shared/observation_distiller_ex.py:147-189; every game's GAME_BG_COLOURS updated with the shared ones) --
by constructing each environment through the reference's own factory, and writes them as constants to
ai_safety_gridworlds_b200/envs/colours.json ({factory name: {character: [r, g, b] on pycolab's 0..999 scale}}).
TEST / BUILD INFRASTRUCTURE (run in the build container, where /root/reference exists); one environment per fresh
interpreter because absl flags are process globals.

    python oracle/dump_colours.py
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "ai_safety_gridworlds_b200", "envs", "colours.json")


def worker(name):
    sys.path[:0] = ["/root/reference", os.path.join(HERE, "stubs"), HERE]
    import shims  # noqa: F401
    # firemaker_ex_ma cannot be constructed at this snapshot without the documented oracle-side shim (shim 4)
    from ai_safety_gridworlds.environments.shared.rl import pycolab_interface_ma as pim
    from ai_safety_gridworlds.environments.shared.ma_reward import ma_reward
    orig = pim.EnvironmentMa._update_for_game_step

    def patched(self, observations, reward, discount, *a, **k):
        if getattr(self, "_last_reward", None) is None:
            self._last_reward = ma_reward({})
        return orig(self, observations, reward, discount, *a, **k)
    pim.EnvironmentMa._update_for_game_step = patched
    from ai_safety_gridworlds.helpers import factory
    env = factory.get_environment_obj(name)
    conv = env._observation_distiller._array_converter
    mapping = getattr(conv, "_colour_mapping", None)
    if mapping is None:                                    # the original suite: pycolab's ObservationToArray pair (safety_game.py)
        mapping = {}
        rgb = conv._renderers["RGB"] if hasattr(conv, "_renderers") else None
        vm = getattr(rgb, "_value_mapping", None)
        if vm is not None:
            mapping = vm
    print(json.dumps({"name": name, "colours": {str(k): [int(x) for x in v] for k, v in mapping.items()}}))


def main():
    sys.path.insert(0, ROOT)
    from ai_safety_gridworlds_b200.envs import ENVIRONMENTS
    out, failed = {}, []
    for name in sorted(ENVIRONMENTS):
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", name], capture_output=True, text=True)
        try:
            rec = json.loads(p.stdout.strip().splitlines()[-1])
            out[rec["name"]] = rec["colours"]
        except Exception:
            failed.append((name, (p.stderr or p.stdout).strip().splitlines()[-1:] ))
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote %s: %d environments" % (OUT, len(out)))
    for name, why in failed:
        print("no colours for %s: %s" % (name, why))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
    else:
        main()
