#!/usr/bin/env python
"""Golden-trace recorder for island_navigation_ex_ma: runs the UNMODIFIED reference
through its PettingZoo parallel wrapper and writes tests/golden/islandma_*.npz.
TEST INFRASTRUCTURE ONLY.

Per parallel step (index 0 = reset; a reset() the recorder had to call because every agent was done
is a step with actions -1) the trace holds, for agents in the fixed order ('1','2'):
  actions    int32  [T,2]       submitted step actions, -1 = agent not in env.agents (it is done)
  order      int8   [T,2]       execution order of this step as agent indices, -1 = no frame
                                (Generator.shuffle in rl/pycolab_interface_ma.py:177-180 when both agents act)
  board      uint8  [T+1,H,W]; cube uint8 [T+1,L,H,W]
  crop1/2    uint8  [T+1,5,5]; lcrop1/2 uint8 [T+1,L,5,5]   per-agent rotated views (zeros once the agent left `agents`)
  reward1/2  float64[T+1,R]     this step's reward vector (zeros at index 0 and for absent agents)
  cum        float64[T+1,2,R]   SafetyEnvironmentMoMa._episode_return per agent (includes what dead agents keep collecting)
  done       uint8  [T+1,2]; step_type int8 [T+1,2]; present uint8 [T+1,2] (agent was in the returned dicts)
  metrics    float64[T+1,M]; pos int16 [T+1,2,2]; adir / odir int8 [T+1,2]; frame int32 [T+1]
  maps       uint8  [T+1,H,W]   environment_data['ascii_art'] of the running game (differs from the level's art only under
                                map_randomization_frequency >= 1, safety_game_mo_base.py:943-1134)

Shims: gymnasium / pettingzoo stubs (oracle/stubs), the None-last-reward guard of
EnvironmentMa._update_for_game_step (SURVEY.md 8c shim 4), and the missing `safety_game_ma` module name in
safety_game_moma (shim 6, see _worker).
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"

CASES = {
    "islandma_default_s0": dict(seed=0, steps=160, kwargs={}),
    "islandma_default_s1": dict(seed=1, steps=120, kwargs={}),
    "islandma_homeostasis_s2": dict(seed=2, steps=160, kwargs=dict(sustainability_challenge=True, thirst_hunger_death=True,
                                                                   penalise_oversatiation=True)),
    "islandma_proportional_s3": dict(seed=3, steps=140, kwargs=dict(penalise_oversatiation=True, use_satiation_proportional_reward=True,
                                                                    sustainability_challenge=True)),
    "islandma_level0_s4": dict(seed=4, steps=100, kwargs=dict(level=0)),
    "islandma_level4_s5": dict(seed=5, steps=120, kwargs=dict(level=4, penalise_oversatiation=True, sustainability_challenge=True)),
    "islandma_level5_s6": dict(seed=6, steps=100, kwargs=dict(level=5)),
    "islandma_level10_s7": dict(seed=7, steps=140, kwargs=dict(level=10, thirst_hunger_death=True, penalise_oversatiation=True)),
    "islandma_fixeddir_s8": dict(seed=8, steps=100, kwargs=dict(observation_direction_mode=0, action_direction_mode=0)),
    "islandma_noshuffle_s9": dict(seed=9, steps=100, kwargs={}, no_shuffle=True),
    "islandma_maxiter15_s10": dict(seed=10, steps=80, kwargs=dict(max_iterations=15, level=6)),
    "islandma_randmap_every_game_s12": dict(seed=12, steps=200, kwargs=dict(map_randomization_frequency=3)),
    "islandma_randmap_level4_s13": dict(seed=13, steps=160, kwargs=dict(map_randomization_frequency=3, level=4, max_iterations=25,
                                                                         penalise_oversatiation=True, sustainability_challenge=True)),
    "islandma_randmap_once_s14": dict(seed=14, steps=120, kwargs=dict(map_randomization_frequency=1)),
    # map resizing (safety_game_mo_base.py:984-1036): a 6 x 7 / 7 x 9 board holding the two agents inside a water border
    "islandma_resized_6x7_s15": dict(seed=15, steps=160, kwargs=dict(map_randomization_frequency=3, map_width=7, map_height=6, max_iterations=30)),
    "islandma_resized_7x9_once_s16": dict(seed=16, steps=120, kwargs=dict(map_randomization_frequency=1, map_width=9, map_height=7, level=8,
                                                                          penalise_oversatiation=True, max_iterations=40)),
    # direction mode 2: TURN_* actions 5..8 turn, moves are relative to the kept direction (safety_game_ma.py:515-768)
    "islandma_turning_actions_s17": dict(seed=17, steps=160, kwargs=dict(observation_direction_mode=2, action_direction_mode=2)),
    "islandma_turning_actions_level4_s18": dict(seed=18, steps=160, kwargs=dict(observation_direction_mode=2, action_direction_mode=2, level=4,
                                                                               penalise_oversatiation=True, sustainability_challenge=True,
                                                                               map_randomization_frequency=3, max_iterations=40)),
    # remove_unused_tile_types_from_layers (safety_game_mo_base.py:1123-1129): the layers are the characters of the board only
    "islandma_remove_unused_s19": dict(seed=19, steps=120, kwargs=dict(remove_unused_tile_types_from_layers=True, level=6)),
    "islandma_remove_unused_randmap_s20": dict(seed=20, steps=140, kwargs=dict(remove_unused_tile_types_from_layers=True, map_randomization_frequency=3,
                                                                              level=2, max_iterations=30)),
    "islandma_level2_s11": dict(seed=11, steps=120, kwargs=dict(level=2, penalise_oversatiation=True, thirst_hunger_death=True,
                                                                 sustainability_challenge=True, max_iterations=60)),
}


def _fuzz_cases():
    """Randomised flag combinations (fixed seed) to widen the pin of the oracle beyond the hand-picked variants."""
    import random
    rnd = random.Random(7120261018)
    out = {}
    for k in range(10):
        kw = {"level": rnd.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 10]), "sustainability_challenge": rnd.random() < 0.5,
              "thirst_hunger_death": rnd.random() < 0.4, "penalise_oversatiation": rnd.random() < 0.6,
              "use_satiation_proportional_reward": rnd.random() < 0.4, "max_iterations": rnd.choice([14, 40, 100]),
              "observation_direction_mode": rnd.choice([0, 1, 1]), "action_direction_mode": rnd.choice([0, 1, 1]),
              "map_randomization_frequency": rnd.choice([0, 0, 3])}
        for prefix in ("DRINK", "FOOD"):
            if rnd.random() < 0.5:
                kw[prefix + "_DEFICIENCY_RATE"] = rnd.choice([-1, -0.5, -2])
            if rnd.random() < 0.5:
                kw[prefix + "_EXTRACTION_RATE"] = rnd.choice([10, 4, 2.5])
            if rnd.random() < 0.4:
                kw[prefix + "_OVERSATIATION_LIMIT"] = rnd.choice([4, 0, 8])
            if rnd.random() < 0.4:
                kw[prefix + "_DEFICIENCY_THRESHOLD"] = rnd.choice([-3, -1, -5.5])
            if rnd.random() < 0.4:
                kw[prefix + "_OVERSATIATION_THRESHOLD"] = rnd.choice([2, 0, 3])
            if rnd.random() < 0.3:
                kw[prefix + "_DEFICIENCY_LIMIT"] = rnd.choice([-20, -6])
        if rnd.random() < 0.3:
            kw["MOVEMENT_REWARD"] = rnd.choice(["{'MOVEMENT_REWARD': 0}", "{'MOVEMENT_REWARD': -2.5}"])
        out["islandma_fuzz_%02d" % k] = dict(seed=700 + k, steps=120, kwargs=kw, no_shuffle=rnd.random() < 0.25)
    return out


CASES.update(_fuzz_cases())
AGENTS = ["1", "2"]


def _worker(name):
    import numpy as np
    sys.path.insert(0, HERE)
    import shims  # noqa: F401
    from ai_safety_gridworlds.environments.shared.rl import pycolab_interface_ma as pim
    from ai_safety_gridworlds.environments.shared.ma_reward import ma_reward
    orig = pim.EnvironmentMa._update_for_game_step

    def patched(self, observations, reward, discount, *a, **k):
        if getattr(self, "_last_reward", None) is None:
            self._last_reward = ma_reward({})
        return orig(self, observations, reward, discount, *a, **k)
    pim.EnvironmentMa._update_for_game_step = patched
    # shim 6: safety_game_moma.AgentSafetySpriteMo.terminate_episode (:1636) calls `safety_game_ma.terminate_episode` but the
    # module only does `from ...safety_game_ma import <names>`: every sprite-initiated termination (thirst/hunger death,
    # the 'U' goal of level 0) dies with a NameError.  Binding the missing module name is the only reading of that line.
    from ai_safety_gridworlds.helpers.gridworld_zoo_parallel_env import GridworldZooParallelEnv
    from ai_safety_gridworlds.environments.shared import safety_game_ma as _sgma, safety_game_moma as _sgmoma
    if not hasattr(_sgmoma, "safety_game_ma"):
        _sgmoma.safety_game_ma = _sgma
    from ai_safety_gridworlds.environments.shared.safety_game_ma import NP_RANDOM

    case = CASES[name]
    env = GridworldZooParallelEnv("island_navigation_ex_ma", seed=case["seed"], **case["kwargs"])
    core = env._env
    if case.get("no_shuffle"):
        core._randomize_agent_actions_order = False          # rl/pycolab_interface_ma.py:177
    log = {"order": None}

    def hook_rng():
        rng = core.environment_data[NP_RANDOM]
        if getattr(rng, "_gw_hooked", False):
            return
        cls = type(rng)

        class Logged(cls):
            def shuffle(self, x, *a, **k):
                super().shuffle(x, *a, **k)
                if isinstance(x, list):                  # the agents' action list; the map randomiser shuffles a numpy array
                    log["order"] = [AGENTS.index(item[0]) for item in x]
        rng.__class__ = Logged
        rng._gw_hooked = True

    names = ["agent_1", "agent_2"]
    rec = {k: [] for k in ("board", "cube", "crop1", "crop2", "lcrop1", "lcrop2", "reward1", "reward2", "cum", "done", "step_type",
                           "present", "metrics", "pos", "adir", "odir", "frame", "maps")}
    meta = {}

    def codes(a):
        return np.vectorize(ord)(a).astype(np.uint8)

    def snapshot(obs, rewards, terms, infos, first):
        game = core._current_game
        any_info = next(iter(infos.values())) if infos else None
        if not meta:
            meta["layer_order"] = list(any_info["info_observation_layers_order"])
            meta["metric_names"] = list(any_info["metrics_dict"].keys())
            meta["reward_keys"] = {a: sorted({k for r in core.enabled_ma_rewards[a] for k, v in r._reward_dimensions_dict.items() if v != 0})
                                   for a in AGENTS}
        R = len(meta["reward_keys"]["1"])
        L = len(meta["layer_order"])
        # the global observation of a step in which every agent finished is not returned by the wrapper: read it from the core
        last = core.last_observations if hasattr(core, "last_observations") else None
        if any_info is not None:
            rec["board"].append(np.array(any_info["ascii_codes"], dtype=np.uint8))
            rec["cube"].append(np.array(any_info["info_observation_layers_cube"], dtype=np.uint8))
            rec["metrics"].append(np.array([float(v) for v in any_info["metrics_dict"].values()], dtype=np.float64))
        else:
            rec["board"].append(np.array(last["ascii_codes"], dtype=np.uint8))
            rec["cube"].append(np.zeros_like(rec["cube"][-1]))
            rec["metrics"].append(np.array([float(v) for v in core.environment_data["metrics_dict"].values()], dtype=np.float64))
        for a, nm in zip(AGENTS, names):
            o = obs.get(nm) if obs else None
            info = infos.get(nm) if infos else None
            rec["crop" + a].append(codes(o[0]) if o is not None else np.zeros((5, 5), np.uint8))
            rec["lcrop" + a].append(np.array(info["info_agent_observation_layers_cube"], dtype=np.uint8) if info is not None
                                    else np.zeros((L, 5, 5), np.uint8))
            r = rewards.get(nm) if rewards else None
            rec["reward" + a].append(np.zeros(R) if (first or r is None or np.isscalar(r)) else np.array(r, dtype=np.float64))
        ret = core._episode_return.tolist(core.enabled_ma_rewards) if getattr(core, "_episode_return", None) is not None else {}
        rec["cum"].append(np.array([np.array(ret.get(a, np.zeros(R)), dtype=np.float64) if not np.isscalar(ret.get(a, None)) else np.zeros(R)
                                    for a in AGENTS]))
        rec["done"].append(np.array([int(bool(terms.get(nm, True))) if terms else 0 for nm in names], dtype=np.uint8))
        rec["present"].append(np.array([int(bool(obs) and nm in obs) for nm in names], dtype=np.uint8))
        st = core._state
        rec["step_type"].append(np.array([int(st[a]) for a in AGENTS], dtype=np.int8))
        rec["pos"].append(np.array([[game.things[a].position.row, game.things[a].position.col] for a in AGENTS], dtype=np.int16))
        rec["adir"].append(np.array([int(game.things[a].action_direction) for a in AGENTS], dtype=np.int8))
        rec["odir"].append(np.array([int(game.things[a].observation_direction) for a in AGENTS], dtype=np.int8))
        rec["frame"].append(int(game.the_plot.frame))
        rec["maps"].append(np.array([[ord(ch) for ch in row] for row in core.environment_data["ascii_art"]], dtype=np.uint8))

    obs, infos = env.reset(seed=case["seed"])
    hook_rng()
    snapshot(obs, None, None, infos, True)
    rng = np.random.default_rng(9000 + case["seed"])
    lo, hi = (0, 8) if case["kwargs"].get("action_direction_mode") == 2 else (0, 4)
    actions, orders = [], []
    for t in range(case["steps"]):
        if not env.agents:                       # every agent is done: the reference needs a reset()
            obs, infos = env.reset()
            hook_rng()
            actions.append([-1, -1]); orders.append([-1, -1])
            snapshot(obs, None, None, infos, True)
            continue
        live = [nm in env.agents for nm in names]
        a = [int(rng.integers(lo, hi + 1)) if live[i] else -1 for i in range(2)]
        log["order"] = None
        obs, rewards, terms, truncs, infos = env.step({nm: a[i] for i, nm in enumerate(names) if live[i]})
        hook_rng()
        actions.append(a)
        if log["order"] is not None:
            orders.append(list(log["order"]))
        else:
            acting = [i for i in range(2) if live[i]]
            orders.append(acting + [-1] * (2 - len(acting)))
        snapshot(obs, rewards, terms, infos, False)

    out = {k: np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v) for k, v in rec.items()}
    out["actions"] = np.array(actions, dtype=np.int32)
    out["order"] = np.array(orders, dtype=np.int8)
    out["frame"] = out["frame"].astype(np.int32)
    kw = dict(case["kwargs"])
    if case.get("no_shuffle"):
        kw["randomize_agent_actions_order"] = False
    meta.update(env="island_navigation_ex_ma", kwargs=kw, seed=case["seed"], amount_agents=2,
                value_mapping={k: float(v) for k, v in core._value_mapping.items()},
                max_iterations=int(core._max_iterations), recorder="oracle/record_island_ma.py",
                reference="levitation-opensource/ai-safety-gridworlds @ /root/reference", numpy=np.__version__)
    out["meta_json"] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("%-28s T=%d episodes=%d max_frame=%d board=%s layers=%s R=%d" % (
        name, len(actions), int((out["actions"][:, 0] == -1).__and__(out["actions"][:, 1] == -1).sum()) + 1, int(out["frame"].max()),
        out["board"].shape[1:], "".join(meta["layer_order"]), out["reward1"].shape[1]))


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
    for name in (argv or list(CASES)):
        rc |= subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", name], env=env).returncode
    return rc


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
