#!/usr/bin/env python
"""Golden-trace recorder for firemaker_ex_ma (BASELINE config 4): runs the UNMODIFIED reference
through its PettingZoo parallel wrapper and writes tests/golden/firemaker_*.npz.
TEST INFRASTRUCTURE ONLY.

Per parallel step (index 0 = reset) the trace holds, for agents in the fixed order ('1','2','S'):
  actions    int32  [T,3]       the submitted step actions
  order      int8   [T,3]       the order the reference executed them in this step
                                (Generator.shuffle in rl/pycolab_interface_ma.py:177-180), as agent indices
  draw_ofs   int64  [T+1]       prefix offsets into `draws`
  draws      float64[D]         every FireDrape `rand()` value in call order (firemaker_ex_ma.py:615,621)
  board      uint8  [T+1,17,17] global rendered board (ASCII codes)
  cube       uint8  [T+1,9,17,17] global layers cube (info_observation_layers_cube)
  crop1/2    uint8  [T+1,5,5]   worker observations (ASCII codes of the '<U1' crop)
  cropS      uint8  [T+1,33,33] supervisor observation
  lcrop1/2   uint8  [T+1,9,5,5] per-agent layers cubes (info_agent_observation_layers_cube)
  lcropS     uint8  [T+1,9,33,33]
  reward1/2  float64[T+1,2]; rewardS float64[T+1,3]   (zeros at index 0)
  cum1/2/S   float64[...]       cumulative_reward per agent
  done       uint8  [T+1,3]     terminateds
  step_type  int8   [T+1,3]     0 FIRST 1 MID 2 LAST 3 DEAD
  metrics    float64[T+1,16]    metrics_dict values in `metric_names` order
  pos        int16  [T+1,3,2]; frame int32 [T+1]; countdown int32 [T+1]; ext_fires int32 [T+1]

Shims (documented in SURVEY.md 8c): gymnasium/pettingzoo stubs (oracle/stubs); the stub
generator's rand() = Generator.random(); EnvironmentMa._update_for_game_step tolerates a None
last reward.  None changes the semantics the traces pin.
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
    "firemaker_s0": dict(seed=0, steps=150, kwargs={}, lo=0, hi=4),
    "firemaker_s1": dict(seed=1, steps=100, kwargs={}, lo=0, hi=4),
    "firemaker_maxiter60_s2": dict(seed=2, steps=70, kwargs={"max_iterations": 60}, lo=0, hi=4),
    # the order flag is switched on the constructed core environment: passing it as a wrapper keyword makes the
    # reference drop amount_agents (it then builds a 2-agent game); the engine semantics are the same either way
    "firemaker_noshuffle_s3": dict(seed=3, steps=100, kwargs={}, no_shuffle=True, lo=0, hi=4),
    "firemaker_turns_s4": dict(seed=4, steps=80, kwargs={}, lo=0, hi=8),
    # the reference's DEFAULT game: amount_agents = 2 = one worker and the supervisor; the '2' tile is lifted from the map
    # (firemaker_ex_ma.py:160,304-363); recorded with columns of agent '2' left at zero
    "firemaker_2agents_s5": dict(seed=5, steps=450, kwargs={}, lo=0, hi=4, amount_agents=2),
    "firemaker_2agents_maxiter40_s6": dict(seed=6, steps=90, kwargs={"max_iterations": 40}, lo=0, hi=4, amount_agents=2),
    # direction modes (firemaker_ex_ma.py:224-226): 1 = actions and views relative to the last move, 2 = turned by the TURN_* actions
    "firemaker_dir1_s7": dict(seed=7, steps=120, kwargs={"observation_direction_mode": 1, "action_direction_mode": 1}, lo=0, hi=4),
    "firemaker_dir2_s8": dict(seed=8, steps=120, kwargs={"observation_direction_mode": 2, "action_direction_mode": 2}, lo=0, hi=8),
    "firemaker_dir1_2agents_s9": dict(seed=9, steps=100, kwargs={"observation_direction_mode": 1, "action_direction_mode": 1, "max_iterations": 90},
                                      lo=0, hi=4, amount_agents=2),
    "firemaker_dir2_maxiter45_s10": dict(seed=10, steps=60, kwargs={"observation_direction_mode": 2, "action_direction_mode": 2, "max_iterations": 45},
                                         lo=0, hi=8),
}
AGENTS = ["1", "2", "S"]


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
    from ai_safety_gridworlds.helpers.gridworld_zoo_parallel_env import GridworldZooParallelEnv
    from ai_safety_gridworlds.environments.shared.safety_game_ma import NP_RANDOM

    case = CASES[name]
    n_agents = case.get("amount_agents", 3)
    present = AGENTS if n_agents == 3 else ["1", "S"]
    env = GridworldZooParallelEnv("firemaker_ex_ma", amount_agents=n_agents, seed=case["seed"], **case["kwargs"])
    assert env.possible_agents == ["agent_" + a for a in present], env.possible_agents
    core = env._env
    if case.get("no_shuffle"):
        core._randomize_agent_actions_order = False          # rl/pycolab_interface_ma.py:177
    log = {"draws": [], "order": None}

    def hook_rng():
        rng = core.environment_data[NP_RANDOM]
        if getattr(rng, "_gw_hooked", False):
            return
        cls = type(rng)

        class Logged(cls):
            def rand(self, *size):
                v = super().rand(*size)
                log["draws"].append(float(v))
                return v

            def shuffle(self, x, *a, **k):
                super().shuffle(x, *a, **k)
                log["order"] = [AGENTS.index(item[0]) for item in x]
        rng.__class__ = Logged
        rng._gw_hooked = True

    rec = {k: [] for k in ("board", "cube", "crop1", "crop2", "cropS", "lcrop1", "lcrop2", "lcropS", "reward1", "reward2", "rewardS",
                           "cum1", "cum2", "cumS", "done", "step_type", "metrics", "pos", "frame", "countdown", "ext_fires", "dirs")}
    meta = {}
    names = ["agent_1", "agent_2", "agent_S"]

    def codes(a):
        return np.vectorize(ord)(a).astype(np.uint8)

    def snapshot(obs, rewards, terms, infos, first):
        i1 = infos["agent_1"]
        if not meta:
            meta["layer_order"] = list(i1["info_observation_layers_order"])
            meta["metric_names"] = list(i1["metrics_dict"].keys())
            # sorted enabled reward dimensions per agent (ma_reward / mo_reward.tolist order)
            meta["reward_keys"] = {a: sorted({k for r in core.enabled_ma_rewards[a] for k, v in r._reward_dimensions_dict.items() if v != 0})
                                   if hasattr(core.enabled_ma_rewards[a][0], "_reward_dimensions_dict") else None for a in present}
        rec["board"].append(np.array(i1["ascii_codes"], dtype=np.uint8))
        rec["cube"].append(np.array(i1["info_observation_layers_cube"], dtype=np.uint8))
        for a, nm in zip(AGENTS, names):
            if a not in present:
                R = 2
                rec["crop" + a].append(np.zeros((5, 5), np.uint8)); rec["lcrop" + a].append(np.zeros((9, 5, 5), np.uint8))
                rec["reward" + a].append(np.zeros(R)); rec["cum" + a].append(np.zeros(R))
                continue
            o = obs.get(nm)
            info = infos[nm]
            crop = codes(o[0]) if o is not None else np.zeros((33, 33) if a == "S" else (5, 5), np.uint8)
            rec["crop" + a].append(crop)
            rec["lcrop" + a].append(np.array(info["info_agent_observation_layers_cube"], dtype=np.uint8))
            R = 3 if a == "S" else 2
            r = rewards.get(nm) if rewards else None
            rec["reward" + a].append(np.zeros(R) if (first or r is None or np.isscalar(r)) else np.array(r, dtype=np.float64))
            rec["cum" + a].append(np.array(info["cumulative_reward"][a], dtype=np.float64))
        rec["done"].append(np.array([(int(bool(terms.get(nm, True))) if terms else 0) if a in present else 0 for a, nm in zip(AGENTS, names)], dtype=np.uint8))
        st = core._state
        rec["step_type"].append(np.array([int(st[a]) if a in present else 0 for a in AGENTS], dtype=np.int8))
        rec["metrics"].append(np.array([float(v) for v in i1["metrics_dict"].values()], dtype=np.float64))
        game = core._current_game
        rec["pos"].append(np.array([[game.things[a].position.row, game.things[a].position.col] if a in present else [-1, -1] for a in AGENTS], dtype=np.int16))
        rec["dirs"].append(np.array([[int(game.things[a].action_direction), int(game.things[a].observation_direction)] if a in present else [2, 2]
                                     for a in AGENTS], dtype=np.int8))    # Directions: LEFT 0, RIGHT 1, UP 2, DOWN 3
        rec["frame"].append(int(game.the_plot.frame))
        rec["countdown"].append(int(core.environment_data["stop_button_press_countdown"]))
        rec["ext_fires"].append(int(getattr(game.things["F"], "number_of_external_fires", 0)))

    obs, infos = env.reset(seed=case["seed"])
    hook_rng()
    snapshot(obs, None, None, infos, True)
    rng = np.random.default_rng(5000 + case["seed"])
    actions, orders, draw_ofs = [], [], [0]
    for t in range(case["steps"]):
        if not env.agents:                       # every agent is done: the next step() call restarts the game
            obs, infos = env.reset()
            hook_rng()
            a = [0, 0, 0]
            actions.append(a); orders.append([0, 1, 2] if n_agents == 3 else [0, 2, -1]); draw_ofs.append(len(log["draws"]))
            snapshot(obs, None, None, infos, True)
            continue
        a = [int(rng.integers(case["lo"], case["hi"] + 1)) for _ in AGENTS]
        log["order"] = None
        obs, rewards, terms, truncs, infos = env.step({nm: a[i] for i, nm in enumerate(names) if AGENTS[i] in present})
        hook_rng()
        actions.append(a)
        default_order = [0, 1, 2] if n_agents == 3 else [0, 2]
        o = log["order"] if log["order"] is not None else default_order
        orders.append(list(o) + [-1] * (3 - len(o)))
        draw_ofs.append(len(log["draws"]))
        snapshot(obs, rewards, terms, infos, False)

    out = {k: np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v) for k, v in rec.items()}
    out["actions"] = np.array(actions, dtype=np.int32)
    out["order"] = np.array(orders, dtype=np.int8)
    out["draw_ofs"] = np.array(draw_ofs, dtype=np.int64)
    out["draws"] = np.array(log["draws"], dtype=np.float64)
    out["frame"] = out["frame"].astype(np.int32)
    out["countdown"] = out["countdown"].astype(np.int32)
    out["ext_fires"] = out["ext_fires"].astype(np.int32)
    kw = dict(case["kwargs"])
    if case.get("no_shuffle"):
        kw["randomize_agent_actions_order"] = False
    if n_agents != 3:
        kw["amount_agents"] = n_agents
    meta.update(env="firemaker_ex_ma", kwargs=kw, seed=case["seed"], amount_agents=n_agents,
                value_mapping={k: float(v) for k, v in core._value_mapping.items()},
                max_iterations=int(core._max_iterations), recorder="oracle/record_firemaker.py",
                reference="levitation-opensource/ai-safety-gridworlds @ /root/reference", numpy=np.__version__)
    out["meta_json"] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("%-26s T=%d draws=%d max_fires=%d frames=%d" % (name, len(actions), len(log["draws"]), int(out["ext_fires"].max()), int(out["frame"].max())))


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
