#!/usr/bin/env python
"""Golden-trace recorder for aintelope_savanna: runs the UNMODIFIED reference through its PettingZoo parallel
wrapper and writes tests/golden/savanna_*.npz.  TEST INFRASTRUCTURE ONLY.

Per parallel step (index 0 = reset; a reset() the recorder had to call because every agent was done is a step with actions -1)
the trace holds, for agents in the fixed order ('0', '1') -- A = amount_agents:
  actions    int32  [T,A]       submitted step actions, -1 = agent not in env.agents
  order      int8   [T,A]       execution order of this step as agent indices, -1 = no frame (Generator.shuffle,
                                rl/pycolab_interface_ma.py:177-180)
  board      uint8  [T+1,H,W]; cube uint8 [T+1,L,H,W]
  crop       uint8  [T+1,A,V,V]; lcrop uint8 [T+1,A,L,V,V]   per-agent rotated views, V = 2 * radius + 1
  reward     float64[T+1,A,R]   this step's reward vector (zeros at index 0 and for absent agents)
  cum        float64[T+1,A,R]   SafetyEnvironmentMoMa._episode_return per agent
  done       uint8  [T+1,A]; step_type int8 [T+1,A]; present uint8 [T+1,A]
  metrics    float64[T+1,M] (nan = the reference has not saved that metric yet); pos int16 [T+1,A,2]; adir / odir int8 [T+1,A]
  frame      int32  [T+1]
  draws      float64[T,32|64]   (64 wide with the sustainability challenge, whose drapes' tile picks follow in call order: every
                                index Generator.choice(n, k, replace=False) returned)  PredatorDrape's draws of the step in call order: Generator.random() ("does it move") and, when it
                                does, the direction Generator.choice returned (Actions value); -1 = unused
  pred       uint8  [T+1,H,W]   the 'P' drape's curtain
  maps       uint8  [T+1,H,W]   environment_data['ascii_art'] of the running game: the randomised layout
                                (safety_game_mo_base.py:943-1134), replayed by the oracle and the kernel

Shims: gymnasium / pettingzoo stubs (oracle/stubs), the None-last-reward guard of EnvironmentMa._update_for_game_step and the
missing `safety_game_ma` module name in safety_game_moma (the same two as oracle/record_island_ma.py).
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"

HOMEOSTASIS = dict(penalise_oversatiation=True)
CASES = {
    "savanna_default_s0": dict(seed=0, steps=400, kwargs={}),
    "savanna_default_s1": dict(seed=1, steps=300, kwargs={}),
    "savanna_maxiter40_s2": dict(seed=2, steps=300, kwargs=dict(max_iterations=40)),
    "savanna_fixed_dirs_s3": dict(seed=3, steps=200, kwargs=dict(observation_direction_mode=0, action_direction_mode=0, max_iterations=60)),
    "savanna_food_drink_homeostasis_s4": dict(seed=4, steps=300, kwargs=dict(amount_drink_holes=2, max_iterations=80, **HOMEOSTASIS)),
    "savanna_small_tiles_gold_silver_s5": dict(seed=5, steps=300, kwargs=dict(amount_small_food_patches=2, amount_small_drink_holes=1,
                                                                              amount_drink_holes=1, amount_gold_deposits=2,
                                                                              amount_silver_deposits=2, max_iterations=90, **HOMEOSTASIS)),
    "savanna_danger_s6": dict(seed=6, steps=300, kwargs=dict(amount_water_tiles=4, max_iterations=70)),
    "savanna_two_agents_s7": dict(seed=7, steps=300, kwargs=dict(amount_agents=2, amount_drink_holes=1, max_iterations=60, **HOMEOSTASIS)),
    "savanna_two_agents_death_s8": dict(seed=8, steps=300, kwargs=dict(amount_agents=2, thirst_hunger_death=True, max_iterations=80,
                                                                       FOOD_DEFICIENCY_LIMIT=-6, FOOD_DEFICIENCY_RATE=-0.5, **HOMEOSTASIS)),
    "savanna_proportional_s9": dict(seed=9, steps=250, kwargs=dict(use_satiation_proportional_reward=True, amount_drink_holes=2,
                                                                   max_iterations=70, **HOMEOSTASIS)),
    "savanna_small_map_s10": dict(seed=10, steps=250, kwargs=dict(level=5, max_iterations=50, amount_food_patches=1)),
    "savanna_radius3_s11": dict(seed=11, steps=200, kwargs=dict(observation_radius=[3, 3, 3, 3], max_iterations=50, amount_water_tiles=2)),
    "savanna_resized_9x11_s12": dict(seed=12, steps=250, kwargs=dict(map_width=11, map_height=9, max_iterations=50, amount_food_patches=3)),
    "savanna_predators_s14": dict(seed=14, steps=300, kwargs=dict(amount_predators=4, max_iterations=60)),
    "savanna_predators_two_agents_s15": dict(seed=15, steps=300, kwargs=dict(amount_agents=2, amount_predators=5, amount_water_tiles=3,
                                                                             amount_drink_holes=2, max_iterations=50, **HOMEOSTASIS)),
    "savanna_predators_always_move_s16": dict(seed=16, steps=250, kwargs=dict(amount_predators=5, PREDATOR_MOVEMENT_PROBABILITY=1.0,
                                                                              max_iterations=40, amount_gold_deposits=3, map_width=9, map_height=9)),
    # sustainability_challenge (:1238-1322, :1388-1472): the drapes' Generator.choice(n, k, replace=False) picks are logged in `draws`
    "savanna_sustainability_s17": dict(seed=17, steps=400, kwargs=dict(sustainability_challenge=True, max_iterations=150)),
    "savanna_sustainability_growth_s18": dict(seed=18, steps=400, kwargs=dict(sustainability_challenge=True, max_iterations=200, amount_food_patches=3,
                                                                             FOOD_GROWTH_LIMIT=8, DRINK_REGROWTH_EXPONENT=1.3, FOOD_EXTRACTION_RATE=0.75,
                                                                             **HOMEOSTASIS)),
    "savanna_sustainability_two_agents_s19": dict(seed=19, steps=400, kwargs=dict(sustainability_challenge=True, amount_agents=2, max_iterations=120,
                                                                                 map_width=7, map_height=7, amount_food_patches=2)),
    "savanna_sustainability_drink_s20": dict(seed=20, steps=300, kwargs=dict(sustainability_challenge=True, amount_food_patches=0, amount_drink_holes=3,
                                                                            max_iterations=100, DRINK_GROWTH_LIMIT=6, **HOMEOSTASIS)),
    "savanna_sustainability_metric_only_s21": dict(seed=21, steps=300, kwargs=dict(
        sustainability_challenge=True, amount_drink_holes=2, amount_small_food_patches=1, amount_gold_deposits=1, amount_water_tiles=2, max_iterations=100,
        use_food_availability_metric_instead_of_spawning_tiles=True, use_drink_availability_metric_instead_of_spawning_tiles=True, **HOMEOSTASIS)),
    # direction mode 2: moves relative to a direction only the TURN_* actions change (safety_game_ma.py:515-768)
    "savanna_turning_actions_s22": dict(seed=22, steps=300, kwargs=dict(observation_direction_mode=2, action_direction_mode=2, max_iterations=80,
                                                                       amount_water_tiles=2)),
    "savanna_turning_actions_two_agents_s23": dict(seed=23, steps=300, kwargs=dict(observation_direction_mode=2, action_direction_mode=2, amount_agents=2,
                                                                                  amount_predators=3, amount_drink_holes=1, max_iterations=60,
                                                                                  observation_radius=[4, 4, 4, 4], **HOMEOSTASIS)),
    "savanna_randmap_once_s13": dict(seed=13, steps=200, kwargs=dict(map_randomization_frequency=1, max_iterations=40)),
    # remove_unused_tile_types_from_layers (safety_game_mo_base.py:1076-1085,1123-1129): tile types that are not on the map lose their
    # sprites / drapes, so the layers of the observation are the characters of the board only
    "savanna_remove_unused_s24": dict(seed=24, steps=200, kwargs=dict(remove_unused_tile_types_from_layers=True, max_iterations=50,
                                                                     amount_water_tiles=2)),
    "savanna_remove_unused_two_agents_s25": dict(seed=25, steps=200, kwargs=dict(remove_unused_tile_types_from_layers=True, amount_agents=2,
                                                                                amount_predators=2, amount_drink_holes=1, max_iterations=50,
                                                                                **HOMEOSTASIS)),
}
# the experiment overlays (experiments/aintelope/*.py) without the sustainability challenge, through the reference's factory names
for _k, _name in enumerate(["danger_tiles", "food_drink_homeostasis", "food_drink_homeostasis_danger_gold_silver", "food_drink_homeostasis_gold",
                            "food_drink_homeostasis_gold_silver", "food_homeostasis", "food_sharing", "food_unbounded", "predators",
                            "food_drink_homeostasis_predators_gold_silver", "savanna_demo", "food_sustainability"]):
    CASES["savanna_exp_" + _name] = dict(seed=40 + _k, steps=150, env=_name, kwargs=dict(max_iterations=60),
                                         agents=2 if _name in ("food_sharing", "predators", "savanna_demo") else 1)
AGENTS = ["0", "1"]


def _worker(name):
    import numpy as np
    sys.path.insert(0, HERE)
    import shims  # noqa: F401
    from absl import flags  # noqa: F401  (safety_game_mo_base patches absl.flags at import time)
    from ai_safety_gridworlds.environments.shared.rl import pycolab_interface_ma as pim
    from ai_safety_gridworlds.environments.shared.ma_reward import ma_reward
    orig = pim.EnvironmentMa._update_for_game_step

    def patched(self, observations, reward, discount, *a, **k):
        if getattr(self, "_last_reward", None) is None:
            self._last_reward = ma_reward({})
        return orig(self, observations, reward, discount, *a, **k)
    pim.EnvironmentMa._update_for_game_step = patched
    from ai_safety_gridworlds.helpers.gridworld_zoo_parallel_env import GridworldZooParallelEnv
    from ai_safety_gridworlds.environments.shared import safety_game_ma as _sgma, safety_game_moma as _sgmoma
    if not hasattr(_sgmoma, "safety_game_ma"):
        _sgmoma.safety_game_ma = _sgma
    from ai_safety_gridworlds.environments.shared.safety_game_ma import NP_RANDOM

    case = CASES[name]
    A = int(case["kwargs"].get("amount_agents", case.get("agents", 1)))
    agents = AGENTS[:A]
    names = ["agent_" + a for a in agents]
    env = GridworldZooParallelEnv(case.get("env", "aintelope_savanna"), seed=case["seed"], **case["kwargs"])
    core = env._env
    log = {"order": None, "draws": []}

    def hook_rng():
        rng = core.environment_data[NP_RANDOM]
        if getattr(rng, "_gw_hooked", False):
            return
        cls = type(rng)

        class Logged(cls):
            def shuffle(self, x, *a, **k):
                super().shuffle(x, *a, **k)
                if isinstance(x, list):                  # the agents' action list; the map randomiser shuffles a numpy array
                    log["order"] = [agents.index(item[0]) for item in x]

            def random(self, *a, **k):                   # PredatorDrape: "does this predator move" (aintelope_savanna.py:1140)
                v = super().random(*a, **k)
                if not a and not k:
                    log["draws"].append(float(v))
                return v

            def choice(self, x, *a, **k):                # PredatorDrape: the direction (:1144); the map randomiser passes an int
                v = super().choice(x, *a, **k)
                if isinstance(x, list):
                    log["draws"].append(float(int(v)))
                elif k.get("replace") is False:          # a resource drape's tile pick (:1287,1316): the indices, in order
                    log["draws"].extend(float(int(i)) for i in np.asarray(v).ravel())
                return v
        rng.__class__ = Logged
        rng._gw_hooked = True

    rec = {k: [] for k in ("board", "cube", "crop", "lcrop", "reward", "cum", "done", "step_type", "present", "metrics", "pos", "adir",
                           "odir", "frame", "maps", "pred")}
    meta = {}

    def codes(a):
        return np.vectorize(ord)(a).astype(np.uint8)

    def snapshot(obs, rewards, terms, infos, first):
        game = core._current_game
        any_info = next(iter(infos.values())) if infos else None
        if not meta:
            meta["layer_order"] = list(any_info["info_observation_layers_order"])
            meta["metric_names"] = list(core.environment_data["metrics_labels"]) if "metrics_labels" in core.environment_data else \
                list(any_info["metrics_dict"].keys())
            meta["reward_keys"] = sorted({k for r in core.enabled_ma_rewards[agents[0]] for k, v in r._reward_dimensions_dict.items() if v != 0})
            meta["view"] = int(np.asarray(obs[names[0]]).shape[-1])
        R, L, V = len(meta["reward_keys"]), len(meta["layer_order"]), meta["view"]
        last = core.last_observations if hasattr(core, "last_observations") else None
        mdict = any_info["metrics_dict"] if any_info is not None else core.environment_data["metrics_dict"]
        rec["metrics"].append(np.array([float(mdict[k]) if k in mdict and mdict[k] is not None else np.nan for k in meta["metric_names"]],
                                       dtype=np.float64))
        if any_info is not None:
            rec["board"].append(np.array(any_info["ascii_codes"], dtype=np.uint8))
            rec["cube"].append(np.array(any_info["info_observation_layers_cube"], dtype=np.uint8))
        else:
            rec["board"].append(np.array(last["ascii_codes"], dtype=np.uint8))
            rec["cube"].append(np.zeros_like(rec["cube"][-1]))
        crops, lcrops, rews = [], [], []
        for a, nm in zip(agents, names):
            o = obs.get(nm) if obs else None
            info = infos.get(nm) if infos else None
            crops.append(codes(o[0]) if o is not None else np.zeros((V, V), np.uint8))
            lcrops.append(np.array(info["info_agent_observation_layers_cube"], dtype=np.uint8) if info is not None
                          else np.zeros((L, V, V), np.uint8))
            r = rewards.get(nm) if rewards else None
            rews.append(np.zeros(R) if (first or r is None or np.isscalar(r)) else np.array(r, dtype=np.float64))
        rec["crop"].append(np.stack(crops)); rec["lcrop"].append(np.stack(lcrops)); rec["reward"].append(np.stack(rews))
        ret = core._episode_return.tolist(core.enabled_ma_rewards) if getattr(core, "_episode_return", None) is not None else {}
        rec["cum"].append(np.array([np.array(ret.get(a, np.zeros(R)), dtype=np.float64) if not np.isscalar(ret.get(a, None)) else np.zeros(R)
                                    for a in agents]))
        rec["done"].append(np.array([int(bool(terms.get(nm, True))) if terms else 0 for nm in names], dtype=np.uint8))
        rec["present"].append(np.array([int(bool(obs) and nm in obs) for nm in names], dtype=np.uint8))
        st = core._state
        rec["step_type"].append(np.array([int(st[a]) for a in agents], dtype=np.int8))
        rec["pos"].append(np.array([[game.things[a].position.row, game.things[a].position.col] for a in agents], dtype=np.int16))
        rec["adir"].append(np.array([int(game.things[a].action_direction) for a in agents], dtype=np.int8))
        rec["odir"].append(np.array([int(game.things[a].observation_direction) for a in agents], dtype=np.int8))
        rec["frame"].append(int(game.the_plot.frame))
        rec["pred"].append(np.array(game.things["P"].curtain, dtype=np.uint8) if "P" in game.things
                           else np.zeros(np.array(infos[names[0]]["ascii_codes"]).shape, np.uint8))   # drape removed from the game
        rec["maps"].append(np.array([[ord(ch) for ch in row] for row in core.environment_data["ascii_art"]], dtype=np.uint8))

    obs, infos = env.reset(seed=case["seed"])
    hook_rng()
    snapshot(obs, None, None, infos, True)
    rng = np.random.default_rng(9000 + case["seed"])
    actions, orders, draws = [], [], []
    K = 64 if case["kwargs"].get("sustainability_challenge") or case.get("env") == "food_sustainability" else 32
    # ^ predator draws of one parallel step: (moves?, direction) per predator and frame
    for t in range(case["steps"]):
        if not env.agents:                       # every agent is done: the reference needs a reset()
            obs, infos = env.reset()
            hook_rng()
            actions.append([-1] * A); orders.append([-1] * A); draws.append([-1.0] * K)
            snapshot(obs, None, None, infos, True)
            continue
        live = [nm in env.agents for nm in names]
        n_act = 9 if case["kwargs"].get("action_direction_mode") == 2 else 5            # mode 2 adds TURN_LEFT_90 .. TURN_RIGHT_180 = 5..8
        a = [int(rng.integers(0, n_act)) if live[i] else -1 for i in range(A)]
        log["order"] = None
        del log["draws"][:]
        obs, rewards, terms, truncs, infos = env.step({nm: a[i] for i, nm in enumerate(names) if live[i]})
        hook_rng()
        actions.append(a)
        assert len(log["draws"]) <= K
        draws.append(list(log["draws"]) + [-1.0] * (K - len(log["draws"])))
        if log["order"] is not None:
            orders.append(list(log["order"]) + [-1] * (A - len(log["order"])))
        else:
            acting = [i for i in range(A) if live[i]]
            orders.append(acting + [-1] * (A - len(acting)))
        snapshot(obs, rewards, terms, infos, False)

    out = {k: np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v) for k, v in rec.items()}
    out["actions"] = np.array(actions, dtype=np.int32)
    out["order"] = np.array(orders, dtype=np.int8)
    out["draws"] = np.array(draws, dtype=np.float64)       # [T, 32]: the predator draws of each step in call order, -1 = unused
    out["frame"] = out["frame"].astype(np.int32)
    meta.update(env=case.get("env", "aintelope_savanna"), kwargs=dict(case["kwargs"]), seed=case["seed"], amount_agents=A,
                value_mapping={k: float(v) for k, v in core._value_mapping.items()},
                max_iterations=int(core._max_iterations), recorder="oracle/record_savanna.py",
                reference="levitation-opensource/ai-safety-gridworlds @ /root/reference", numpy=np.__version__)
    out["meta_json"] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("%-40s T=%d resets=%d max_frame=%d board=%s view=%d layers=%s R=%s M=%s" % (
        name, len(actions), int((out["actions"] == -1).all(axis=1).sum()) + 1, int(out["frame"].max()), out["board"].shape[1:],
        meta["view"], "".join(meta["layer_order"]), meta["reward_keys"], meta["metric_names"]))


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
