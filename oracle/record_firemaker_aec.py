#!/usr/bin/env python
"""Golden-trace recorder for firemaker_ex_ma through the reference's AEC wrapper
(helpers/gridworld_zoo_aec_env.py:607-806): runs the UNMODIFIED reference's `GridworldZooAecEnv`
agent by agent and writes tests/golden/firemaker_aec_*.npz.  TEST INFRASTRUCTURE ONLY.

One AEC step = ONE single-agent `EnvironmentMa.step({agent: action})` = one engine frame
(rl/pycolab_interface_ma.py:173-246).  Index 0 of every per-step array is the state after reset().

  agent      int8   [T]         index ('1','2','S' -> 0,1,2) of `agent_selection` when step k was made
  action     int32  [T]         the submitted action, -1 = None (the "dead step" of a finished agent, :623-648)
  draw_ofs   int64  [T+1]; draws float64 [D]    FireDrape rand() values per step, call order
  board      uint8  [T+1,17,17]; cube uint8 [T+1,9,17,17]          global observation after the step
  crop1/2    uint8  [T+1,5,5]; cropS uint8 [T+1,33,33]             observe(agent) of every agent after the step
  lcrop1/2/S uint8  [T+1,9,h,w]                                    observe_info(agent) layers cubes
  reward1/2  float64[T+1,2]; rewardS float64 [T+1,3]               env.rewards after the step (zeros when 0.0 / removed)
  cum1/2/S   float64[...]                                          env._cumulative_rewards (the value last() returns)
  term       uint8  [T+1,3]     env.terminations (1 also for removed agents)
  alive      uint8  [T+1,3]     agent in env.agents
  selection  int8   [T+1]       index of agent_selection after the step (-1 = None)
  step_type  int8   [T+1,3]; frame int32 [T+1]; pos int16 [T+1,3,2]

Shims: as oracle/record_firemaker.py.
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
    "firemaker_aec_s0": dict(seed=0, steps=240, kwargs={}, lo=0, hi=4),
    "firemaker_aec_maxiter40_s1": dict(seed=1, steps=130, kwargs={"max_iterations": 40}, lo=0, hi=4),
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
    from ai_safety_gridworlds.helpers.gridworld_zoo_aec_env import GridworldZooAecEnv
    from ai_safety_gridworlds.environments.shared.safety_game_ma import NP_RANDOM

    case = CASES[name]
    env = GridworldZooAecEnv("firemaker_ex_ma", amount_agents=3, seed=case["seed"], **case["kwargs"])
    core = env._env
    log = {"draws": []}

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
        rng.__class__ = Logged
        rng._gw_hooked = True

    names = ["agent_1", "agent_2", "agent_S"]
    rec = {k: [] for k in ("board", "cube", "crop1", "crop2", "cropS", "lcrop1", "lcrop2", "lcropS", "reward1", "reward2", "rewardS",
                           "cum1", "cum2", "cumS", "term", "alive", "selection", "step_type", "frame", "pos")}
    meta = {}

    def codes(a):
        return np.vectorize(ord)(a).astype(np.uint8)

    def vec(x, R):
        if x is None or np.isscalar(x):
            return np.zeros(R)
        return np.array(x, dtype=np.float64)

    def snapshot():
        info1 = env.observe_info("agent_1")
        if not meta:
            meta["layer_order"] = list(info1["info_observation_layers_order"])
        rec["board"].append(np.array(info1["ascii_codes"], dtype=np.uint8))
        rec["cube"].append(np.array(info1["info_observation_layers_cube"], dtype=np.uint8))
        for a, nm in zip(AGENTS, names):
            R = 3 if a == "S" else 2
            rec["crop" + a].append(codes(env.observe(nm)[0]))
            rec["lcrop" + a].append(np.array(env.observe_info(nm)["info_agent_observation_layers_cube"], dtype=np.uint8))
            rec["reward" + a].append(vec(env.rewards.get(nm), R))
            rec["cum" + a].append(vec(env._cumulative_rewards.get(nm), R))
        rec["term"].append(np.array([int(bool(env.terminations.get(nm, True))) for nm in names], dtype=np.uint8))
        rec["alive"].append(np.array([int(nm in env.agents) for nm in names], dtype=np.uint8))
        rec["selection"].append(names.index(env.agent_selection) if env.agent_selection is not None else -1)
        st = core._state
        rec["step_type"].append(np.array([int(st[a]) for a in AGENTS], dtype=np.int8))
        game = core._current_game
        rec["pos"].append(np.array([[game.things[a].position.row, game.things[a].position.col] for a in AGENTS], dtype=np.int16))
        rec["frame"].append(int(game.the_plot.frame))

    env.reset(seed=case["seed"])
    hook_rng()
    snapshot()
    rng = np.random.default_rng(7000 + case["seed"])
    agents, actions, draw_ofs = [], [], [0]
    for t in range(case["steps"]):
        sel = env.agent_selection
        if sel is None:
            break
        dead = env.terminations[sel] or env.truncations[sel]
        a = None if dead else int(rng.integers(case["lo"], case["hi"] + 1))
        try:
            env.step(a)
        except KeyError as exc:
            # reference defect: once a finished agent has made its dead step and left `agents`, the next live
            # agent's step credits it a reward and fails (gridworld_zoo_aec_env.py:757-759); the trace ends here
            meta["ended_by"] = "KeyError(%s) at AEC step %d" % (exc, t)
            break
        hook_rng()
        agents.append(names.index(sel))
        actions.append(-1 if a is None else a)
        draw_ofs.append(len(log["draws"]))
        snapshot()

    out = {k: np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v) for k, v in rec.items()}
    out["agent"] = np.array(agents, dtype=np.int8)
    out["action"] = np.array(actions, dtype=np.int32)
    out["draw_ofs"] = np.array(draw_ofs, dtype=np.int64)
    out["draws"] = np.array(log["draws"], dtype=np.float64)
    out["frame"] = out["frame"].astype(np.int32)
    out["selection"] = out["selection"].astype(np.int8)
    meta.update(env="firemaker_ex_ma", wrapper="GridworldZooAecEnv", kwargs=dict(case["kwargs"]), seed=case["seed"], amount_agents=3,
                max_iterations=int(core._max_iterations), recorder="oracle/record_firemaker_aec.py",
                reference="levitation-opensource/ai-safety-gridworlds @ /root/reference", numpy=np.__version__)
    out["meta_json"] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("%-28s T=%d draws=%d frames=%d dead_steps=%d" % (name, len(actions), len(log["draws"]), int(out["frame"].max()),
                                                           int((out["action"] < 0).sum())))


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
