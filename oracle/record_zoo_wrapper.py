#!/usr/bin/env python
"""Golden traces of the reference's PettingZoo parallel WRAPPER options: runs the UNMODIFIED
`GridworldZooParallelEnv('island_navigation_ex_ma', ...)` (helpers/gridworld_zoo_parallel_env.py) with
`layers_order_in_cube`, `layers_order_in_cube_per_agent`, the coordinate infos, `use_transitions` / `flatten_observations`,
the float board format and `test_death`, on stored actions, logging the generator calls the replay needs (the agents'
shuffle order per step, and every `np_random.random()` draw of the wrapper's test_death branch, :577-586).
Writes tests/golden/zoo_<case>.npz.  TEST INFRASTRUCTURE ONLY; one case per fresh interpreter.

    python oracle/record_zoo_wrapper.py [case ...]
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"
AGENTS = ["1", "2"]
NAMES = ["agent_1", "agent_2"]
CASES = {
    "zoo_islandma_layers_coords_s3": dict(seed=3, steps=45, env_kwargs={}, wrapper=dict(
        layers_order_in_cube=["W", "1", "Z", "2", "#"], layers_order_in_cube_per_agent={"agent_1": ["2", "Q", "1", " "]})),
    "zoo_islandma_test_death_s4": dict(seed=4, steps=70, env_kwargs={"max_iterations": 30}, wrapper=dict(test_death=True, test_death_probability=0.25)),
    "zoo_islandma_transitions_flat_s5": dict(seed=5, steps=25, env_kwargs={}, wrapper=dict(use_transitions=True, flatten_observations=True)),
    "zoo_islandma_board_format_s6": dict(seed=6, steps=25, env_kwargs={"level": 2}, wrapper=dict(ascii_observation_format=False, use_transitions=True)),
}


def _worker(name):
    import numpy as np
    sys.path.insert(0, HERE)
    import shims  # noqa: F401
    from ai_safety_gridworlds.helpers.gridworld_zoo_parallel_env import GridworldZooParallelEnv
    from ai_safety_gridworlds.environments.shared import safety_game_ma as _sgma, safety_game_moma as _sgmoma
    if not hasattr(_sgmoma, "safety_game_ma"):               # shim 6 of oracle/record_island_ma.py
        _sgmoma.safety_game_ma = _sgma
    from ai_safety_gridworlds.environments.shared.safety_game_ma import NP_RANDOM
    case = CASES[name]
    env = GridworldZooParallelEnv("island_navigation_ex_ma", seed=case["seed"], **case["wrapper"], **case["env_kwargs"])
    core = env._env
    log = {"order": None, "random": []}

    def hook(rng):
        if getattr(rng, "_gw_hooked", False):
            return
        cls = type(rng)

        class Logged(cls):
            def shuffle(self, x, *a, **k):
                super().shuffle(x, *a, **k)
                if isinstance(x, list):
                    log["order"] = [AGENTS.index(item[0]) for item in x]

            def random(self, *a, **k):
                v = super().random(*a, **k)
                if not a and not k:
                    log["random"].append(float(v))
                return v
        rng.__class__ = Logged
        rng._gw_hooked = True

    def hook_all():
        hook(core.environment_data[NP_RANDOM])
        hook(env._np_random)

    ascii_fmt = case["wrapper"].get("ascii_observation_format", True)

    def obs_array(o):
        if o is None:
            return None
        a = np.asarray(o)
        if a.dtype.kind == "U":
            return np.vectorize(lambda ch: ord(ch) if ch else 0)(a).astype(np.uint8)
        return a.astype(np.float32)

    steps = []

    def snapshot(obs, rewards, terms, infos, actions, order):
        rec = {"actions": actions, "order": order, "death_draws": list(log["random"]), "agents_after": list(env.agents), "per_agent": {}}
        for nm in NAMES:
            pa = {}
            o = obs_array(obs.get(nm)) if obs else None
            pa["obs"] = None if o is None else o.tolist()
            pa["obs_shape"] = None if o is None else list(o.shape)
            r = None if rewards is None else rewards.get(nm)
            pa["has_reward"] = rewards is not None and nm in rewards
            pa["reward"] = None if r is None or np.isscalar(r) else [float(x) for x in np.asarray(r).reshape(-1)]
            pa["has_done"] = terms is not None and nm in terms
            pa["done"] = None if terms is None or nm not in terms else bool(terms[nm])
            info = infos.get(nm) if infos else None
            if info is not None:
                pa["layers_order"] = list(info.get("info_observation_layers_order", []))
                pa["layers_cube"] = np.asarray(info["info_observation_layers_cube"]).astype(np.uint8).tolist() if "info_observation_layers_cube" in info else None
                pa["agent_layers_order"] = list(info.get("info_agent_observation_layers_order", []))
                pa["agent_layers_cube"] = np.asarray(info["info_agent_observation_layers_cube"]).astype(np.uint8).tolist() if "info_agent_observation_layers_cube" in info else None
                co = info.get("info_observation_coordinates")
                pa["coordinates"] = None if co is None else {k: [[int(a), int(b)] for a, b in v] for k, v in co.items()}
                ac = info.get("info_agent_observation_coordinates")
                pa["agent_coordinates"] = None if ac is None else ([] if isinstance(ac, list) else {k: [[int(a), int(b)] for a, b in v] for k, v in ac.items()})
                ld = info.get("info_agent_observation_layers_dict")
                pa["agent_layers_dict_keys"] = None if ld is None else sorted(ld.keys())
            rec["per_agent"][nm] = pa
        steps.append(rec)

    spaces = {nm: dict(shape=[int(x) for x in env.observation_space(nm).shape], dtype=str(env.observation_space(nm).dtype)) for nm in NAMES}
    obs, infos = env.reset(seed=case["seed"])
    hook_all()
    snapshot(obs, None, None, infos, None, None)
    rng = np.random.default_rng(5000 + case["seed"])
    for t in range(case["steps"]):
        if not env.agents:
            obs, infos = env.reset()
            hook_all()
            log["random"] = []
            snapshot(obs, None, None, infos, [-1, -1], [-1, -1])
            continue
        live = [nm in env.agents for nm in NAMES]
        a = [int(rng.integers(0, 5)) if live[i] else -1 for i in range(2)]
        log["order"], log["random"] = None, []
        obs, rewards, terms, truncs, infos = env.step({nm: a[i] for i, nm in enumerate(NAMES) if live[i]})
        hook_all()
        if log["order"] is not None:
            order = list(log["order"])
        else:
            acting = [i for i in range(2) if live[i]]
            order = acting + [-1] * (2 - len(acting))
        snapshot(obs, rewards, terms, infos, a, order)
    meta = dict(env="island_navigation_ex_ma", env_kwargs=case["env_kwargs"], wrapper=case["wrapper"], seed=case["seed"], spaces=spaces,
                recorder="oracle/record_zoo_wrapper.py", numpy=np.__version__)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), steps_json=np.array(json.dumps(steps)), meta_json=np.array(json.dumps(meta)))
    print("%-36s T=%d resets=%d death draws=%d" % (name, len(steps) - 1, sum(1 for s in steps[1:] if s["actions"] == [-1, -1]),
                                                   sum(len(s["death_draws"]) for s in steps)))


def main(argv):
    if len(argv) >= 2 and argv[0] == "--worker":
        _worker(argv[1])
        return 0
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "stubs"), REFERENCE])
    rc = 0
    for name in (argv or list(CASES)):
        rc |= subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", name], env=env).returncode
    return rc


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
