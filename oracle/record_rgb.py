#!/usr/bin/env python
"""Golden RGB observations: runs the UNMODIFIED reference environments and records, per timestep, the rendered board (ASCII
codes, pycolab Engine._board.board) and the `RGB` entry of the observation its distiller produces
(environments/shared/observation_distiller_ex.py:147-189: uint8 [3, H, W] = colour / 999 * 255) -> tests/golden/rgb_<case>.npz.
TEST INFRASTRUCTURE ONLY.  One case per fresh interpreter (absl flags are process globals).

    python oracle/record_rgb.py
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
CASES = {
    "island_navigation_ex": dict(kwargs={}, steps=60, agents=None),
    "boat_race_ex": dict(kwargs={"level": 3}, steps=60, agents=None),
    "whisky_gold": dict(kwargs={}, steps=30, agents=None),
    "side_effects_sokoban": dict(kwargs={"level": 1}, steps=40, agents=None),
    "firemaker_ex_ma": dict(kwargs={"amount_agents": 3}, steps=60, agents=["1", "2", "S"]),
    "island_navigation_ex_ma": dict(kwargs={}, steps=40, agents=["1", "2"]),
    "aintelope_savanna": dict(kwargs={}, steps=40, agents=["0"]),
}


def worker(name):
    import numpy as np
    sys.path[:0] = ["/root/reference", os.path.join(HERE, "stubs"), HERE]
    import shims  # noqa: F401
    from ai_safety_gridworlds.environments.shared.rl import pycolab_interface_ma as pim
    from ai_safety_gridworlds.environments.shared.ma_reward import ma_reward
    orig = pim.EnvironmentMa._update_for_game_step

    def patched(self, observations, reward, discount, *a, **k):          # SURVEY 8c shim 4 (firemaker_ex_ma)
        if getattr(self, "_last_reward", None) is None:
            self._last_reward = ma_reward({})
        return orig(self, observations, reward, discount, *a, **k)
    pim.EnvironmentMa._update_for_game_step = patched
    from ai_safety_gridworlds.helpers import factory
    case = CASES[name]
    kw = dict(case["kwargs"])
    if case["agents"] is not None:
        import gymnasium.utils.seeding as seeding
        rng0 = seeding.np_random(3)[0]
        cls = type(rng0)
        if not hasattr(rng0, "rand"):
            class WithRand(cls):
                def rand(self, *size):
                    return self.random(size if size else None)
            rng0.__class__ = WithRand
        kw["np_random"] = rng0
    env = factory.get_environment_obj(name, **kw)
    rng = np.random.default_rng(11)
    boards, rgbs = [], []
    ts = env.reset()
    for t in range(case["steps"] + 1):
        if t > 0:
            if case["agents"] is None:
                ts = env.step(int(rng.integers(1, 5)))
            else:
                ts = env.step({a: {"step": int(rng.integers(0, 5))} for a in case["agents"]})
            if (ts.step_type.last() if case["agents"] is None else any(st.last() or st.dead() for st in ts.step_type.values())):
                ts = env.reset()
        boards.append(np.array(env._current_game._board.board, np.uint8))
        rgbs.append(np.array(ts.observation["RGB"], np.uint8))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "rgb_" + name + ".npz"), board=np.stack(boards), rgb=np.stack(rgbs),
                        meta_json=json.dumps(dict(env=name, kwargs=case["kwargs"])))
    print("recorded", name, np.stack(boards).shape, np.stack(rgbs).shape)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
    else:
        for nm in (sys.argv[1:] or sorted(CASES)):
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", nm], capture_output=True, text=True)
            print((p.stdout.strip().splitlines() or p.stderr.strip().splitlines()[-3:])[-1] if (p.stdout.strip() or p.stderr.strip()) else "?")
            if p.returncode:
                print(p.stderr[-1500:])
