#!/usr/bin/env python
"""Golden-trace recorder: runs the UNMODIFIED reference and writes tests/golden/*.npz.

TEST INFRASTRUCTURE ONLY.  Nothing under ai_safety_gridworlds_b200/ may import this.

The reference (levitation-opensource/ai-safety-gridworlds, mounted read-only at
/root/reference) is pure Python, so it can be imported in the build container but
cannot travel to the GPU box.  This script drives it through its own public boundary
(`GridworldGymEnv.reset/step`, helpers/gridworld_gym_env.py:455,588) on explicit,
stored action arrays and records, per timestep, everything the CUDA path and the C
oracle must reproduce:

  board      uint8  [T+1,H,W]   info['ascii_codes']   (rendered board, pycolab/engine.py:737)
  obs        float32[T+1,H,W]   value-mapped board, the Gym observation (gridworld_gym_env.py:525-536)
  cube       uint8  [T+1,L,H,W] info['info_observation_layers_cube'] (safety_game_mo.py:487-506)
  reward     float64[T+1,R]     reward vector (safety_game_mo.py:1049-1066); zeros on FIRST
  step_type  int8   [T+1]       0 FIRST / 1 MID / 2 LAST (rl/environment.py)
  reason     int8   [T+1]       TerminationReason or -1 (termination_reason_enum.py:24-39)
  discount   float64[T+1]       nan where the reference returns None
  cumulative float64[T+1,R]     info['cumulative_reward'] (safety_game_mo.py:1027-1044)
  average    float64[T+1,R]     info['average_reward']
  scalars    float64[T+1,5]     gini, cumulative gini, mo_var, cum mo_var, avg mo_var (:1071-1084)
  metrics    float64[T+1,M]     info['metrics_dict'] values in `metric_names` order
  pos        int16  [T+1,2]     agent sprite (row, col)
  safety     int16  [T+1]       environment_data['safety'] (island_navigation_ex.py:461-469), -1 if absent
  frame      int32  [T+1]       the_plot.frame

Index 0 is the reset() timestep; index t>=1 is the result of step(actions[t-1]).
Stepping continues through episode ends exactly as the reference does when a caller
keeps calling step(): the call after a LAST timestep ignores its action, rebuilds the
game and returns a FIRST timestep (rl/pycolab_interface_mo.py:175-178).

Each case runs in a fresh interpreter because the reference keeps absl flags and
class-level statics as process globals (island_navigation_ex.py:227-337,
safety_game_mo.py:318-384).

Usage:  python oracle/record.py            # (re)generate every case
        python oracle/record.py NAME ...   # only the named cases
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"

# policy: "uniform" = U{lo..hi}; "safe" = uniform over {0..4} moves that do not enter
# `avoid` tiles (keeps episodes long enough to exercise regrowth / visit counters);
# "mixed" = safe with probability 0.85 else uniform over lo..hi.
CASES = {
    # ---- island_navigation_ex, config 1/3 -------------------------------
    "island_ex_default_s0": dict(env="island_navigation_ex", kwargs={}, steps=400, seed=0, policy="uniform", lo=0, hi=4),
    "island_ex_default_s1": dict(env="island_navigation_ex", kwargs={}, steps=400, seed=1, policy="uniform", lo=0, hi=4),
    "island_ex_default_safe_s2": dict(env="island_navigation_ex", kwargs={}, steps=600, seed=2, policy="safe", avoid="W", lo=0, hi=4),
    "island_ex_default_mixed_s3": dict(env="island_navigation_ex", kwargs={}, steps=600, seed=3, policy="mixed", avoid="W", lo=0, hi=4),
    "island_ex_allactions_s4": dict(env="island_navigation_ex", kwargs={}, steps=400, seed=4, policy="mixed", avoid="W", lo=0, hi=9),
    "island_ex_level2_s5": dict(env="island_navigation_ex", kwargs={"level": 2}, steps=300, seed=5, policy="uniform", lo=0, hi=4),
    "island_ex_level3_s6": dict(env="island_navigation_ex", kwargs={"level": 3}, steps=300, seed=6, policy="uniform", lo=0, hi=4),
    "island_ex_level4_s7": dict(env="island_navigation_ex", kwargs={"level": 4}, steps=300, seed=7, policy="uniform", lo=0, hi=4),
    "island_ex_level5_s8": dict(env="island_navigation_ex", kwargs={"level": 5}, steps=300, seed=8, policy="mixed", avoid="W", lo=0, hi=4),
    "island_ex_level6_s9": dict(env="island_navigation_ex", kwargs={"level": 6}, steps=300, seed=9, policy="mixed", avoid="W", lo=0, hi=4),
    "island_ex_level7_s10": dict(env="island_navigation_ex", kwargs={"level": 7}, steps=300, seed=10, policy="mixed", avoid="W", lo=0, hi=4),
    "island_ex_level8_s11": dict(env="island_navigation_ex", kwargs={"level": 8}, steps=300, seed=11, policy="mixed", avoid="W", lo=0, hi=4),
    "island_ex_nosustain_s12": dict(env="island_navigation_ex", kwargs={"sustainability_challenge": False}, steps=400, seed=12, policy="safe", avoid="W", lo=0, hi=4),
    "island_ex_death_s13": dict(env="island_navigation_ex", kwargs={"thirst_hunger_death": True}, steps=400, seed=13, policy="safe", avoid="W", lo=0, hi=4),
    "island_ex_nooversat_s14": dict(env="island_navigation_ex", kwargs={"penalise_oversatiation": False}, steps=400, seed=14, policy="safe", avoid="W", lo=0, hi=4),
    "island_ex_proportional_s15": dict(env="island_navigation_ex", kwargs={"use_satiation_proportional_reward": True}, steps=400, seed=15, policy="safe", avoid="W", lo=0, hi=4),
    "island_ex_maxiter20_s16": dict(env="island_navigation_ex", kwargs={"max_iterations": 20}, steps=300, seed=16, policy="safe", avoid="W", lo=0, hi=4),
    "island_ex_noops_off_s17": dict(env="island_navigation_ex", kwargs={"noops": False}, steps=300, seed=17, policy="mixed", avoid="W", lo=1, hi=4),
    # experiment-overlay style flag values (experiments/food_drink_bounded_death_gold_silver.py:32-150)
    "island_ex_overlay_s18": dict(env="island_navigation_ex", kwargs={
        "level": 4, "sustainability_challenge": False, "thirst_hunger_death": True, "penalise_oversatiation": False,
        "MOVEMENT_REWARD": "{'MOVEMENT_REWARD': 0}", "DRINK_REWARD": "{'DRINK_REWARD': 0}", "FOOD_REWARD": "{'FOOD_REWARD': 0}",
        "GAP_REWARD": "{'FOOD_REWARD': 0, 'DRINK_REWARD': 0}",
        "DRINK_EXTRACTION_RATE": 7, "FOOD_EXTRACTION_RATE": 7, "DRINK_OVERSATIATION_LIMIT": 0, "FOOD_OVERSATIATION_LIMIT": 0},
        steps=400, seed=18, policy="uniform", lo=0, hi=4),
    "island_ex_fractional_s19": dict(env="island_navigation_ex", kwargs={
        "use_satiation_proportional_reward": True, "DRINK_DEFICIENCY_RATE": -0.3, "FOOD_DEFICIENCY_RATE": -0.7,
        "DRINK_EXTRACTION_RATE": 3.5, "FOOD_EXTRACTION_RATE": 6.25, "DRINK_REGROWTH_EXPONENT": 1.3,
        "DRINK_AVAILABILITY_INITIAL": 12.5, "FOOD_GROWTH_LIMIT": 15, "FOOD_AVAILABILITY_INITIAL": 15},
        steps=500, seed=19, policy="safe", avoid="W", lo=0, hi=4),
    # ---- experiment overlays (ai_safety_gridworlds/experiments/*.py), through the reference's factory names ----
    "island_exp_food_bounded": dict(env="food_bounded", kwargs={}, steps=200, seed=30, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_bounded": dict(env="food_drink_bounded", kwargs={}, steps=200, seed=31, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_bounded_death": dict(env="food_drink_bounded_death", kwargs={}, steps=200, seed=32, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_bounded_death_gold": dict(env="food_drink_bounded_death_gold", kwargs={}, steps=200, seed=33, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_bounded_death_gold_silver": dict(env="food_drink_bounded_death_gold_silver", kwargs={}, steps=200, seed=34, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_bounded_gold": dict(env="food_drink_bounded_gold", kwargs={}, steps=200, seed=35, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_bounded_gold_silver": dict(env="food_drink_bounded_gold_silver", kwargs={}, steps=200, seed=36, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_rolf": dict(env="food_drink_rolf", kwargs={}, steps=200, seed=37, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_rolf_gold_as_gap": dict(env="food_drink_rolf_gold_as_gap", kwargs={}, steps=200, seed=38, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_rolf_gold_as_resource": dict(env="food_drink_rolf_gold_as_resource", kwargs={}, steps=200, seed=39, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_rolf_gold_as_resource_scaled": dict(env="food_drink_rolf_gold_as_resource_scaled", kwargs={}, steps=200, seed=40, policy="uniform", lo=0, hi=4),
    "island_exp_food_drink_unbounded": dict(env="food_drink_unbounded", kwargs={"max_iterations": 50}, steps=200, seed=41, policy="uniform", lo=0, hi=4),
    # ---- boat_race_ex, config 2 ---------------------------------------------
    "boat_ex_level3_s0": dict(env="boat_race_ex", kwargs={"level": 3}, steps=500, seed=0, policy="uniform", lo=0, hi=4),
    "boat_ex_level3_s1": dict(env="boat_race_ex", kwargs={"level": 3}, steps=500, seed=1, policy="uniform", lo=0, hi=4),
    "boat_ex_level3_allactions_s2": dict(env="boat_race_ex", kwargs={"level": 3}, steps=400, seed=2, policy="uniform", lo=0, hi=9),
    "boat_ex_level2_s3": dict(env="boat_race_ex", kwargs={"level": 2}, steps=400, seed=3, policy="uniform", lo=0, hi=4),
    "boat_ex_level1_s4": dict(env="boat_race_ex", kwargs={"level": 1}, steps=400, seed=4, policy="uniform", lo=0, hi=4),
    "boat_ex_level0_s5": dict(env="boat_race_ex", kwargs={"level": 0}, steps=400, seed=5, policy="uniform", lo=0, hi=4),
    "boat_ex_level3_norep_s6": dict(env="boat_race_ex", kwargs={"level": 3, "repetition_penalty": False}, steps=300, seed=6, policy="uniform", lo=0, hi=4),
    "boat_ex_level3_noiter_s7": dict(env="boat_race_ex", kwargs={"level": 3, "iterations_penalty": False}, steps=300, seed=7, policy="uniform", lo=0, hi=4),
    "boat_ex_level3_maxiter30_s8": dict(env="boat_race_ex", kwargs={"level": 3, "max_iterations": 30}, steps=300, seed=8, policy="uniform", lo=0, hi=4),
    "boat_ex_level2_noops_off_s9": dict(env="boat_race_ex", kwargs={"level": 2, "noops": False}, steps=300, seed=9, policy="uniform", lo=1, hi=4),
}


def _fuzz_cases():
    """Randomised flag combinations (fixed seed): levels, switches, fractional rates / limits / thresholds, multi-dimensional and
    zero reward flags, short cut-offs.  They widen the pin of the oracle beyond the hand-picked variants above."""
    import random
    rnd = random.Random(20261018)
    out = {}
    for k in range(16):
        kw = {"level": rnd.randrange(0, 10), "sustainability_challenge": rnd.random() < 0.5, "thirst_hunger_death": rnd.random() < 0.4,
              "penalise_oversatiation": rnd.random() < 0.6, "use_satiation_proportional_reward": rnd.random() < 0.4,
              "max_iterations": rnd.choice([12, 40, 100, 100])}
        for prefix in ("DRINK", "FOOD"):
            if rnd.random() < 0.6:
                kw[prefix + "_DEFICIENCY_RATE"] = rnd.choice([-1, -0.5, -2, -0.25])
            if rnd.random() < 0.6:
                kw[prefix + "_EXTRACTION_RATE"] = rnd.choice([10, 4, 7, 2.5, 20])
            if rnd.random() < 0.5:
                kw[prefix + "_OVERSATIATION_LIMIT"] = rnd.choice([4, 0, 8, 2.5])
            if rnd.random() < 0.4:
                kw[prefix + "_DEFICIENCY_LIMIT"] = rnd.choice([-20, -6, -10])
            if rnd.random() < 0.4:
                kw[prefix + "_DEFICIENCY_INITIAL"] = rnd.choice([0, -2, 3])
            if rnd.random() < 0.4:
                kw[prefix + "_AVAILABILITY_INITIAL"] = rnd.choice([20, 8, 12.5])
            if rnd.random() < 0.3:
                kw[prefix + "_GROWTH_LIMIT"] = rnd.choice([20, 15, 30])
        if rnd.random() < 0.4:
            kw["DRINK_REGROWTH_EXPONENT"] = rnd.choice([1.1, 1.3, 1.05])
        if rnd.random() < 0.3:
            kw["MOVEMENT_REWARD"] = rnd.choice(["{'MOVEMENT_REWARD': 0}", "{'MOVEMENT_REWARD': -2.5}", "{'MOVEMENT_REWARD': -1, 'DRINK_REWARD': 0.5}"])
        if rnd.random() < 0.3:
            kw["GOLD_REWARD"] = rnd.choice(["{'GOLD_REWARD': 15}", "{'GOLD_REWARD': 40, 'SILVER_REWARD': 1}"])
        if rnd.random() < 0.3:
            kw["DRINK_REWARD"] = rnd.choice(["{'DRINK_REWARD': 0}", "{'DRINK_REWARD': 7.5}"])
        if k == 14:
            # level 0 (no drink tile) with DRINK_DEFICIENCY_INITIAL = -2: the reference itself raises "Reward DRINK_DEFICIENCY_REWARD is
            # not enabled but is still included in mo_reward with nonzero value" at the first step (mo_reward.py:198); the spec compiler
            # raises the same error eagerly (tests/test_abi_and_host.py::test_reference_rejected_flags_are_rejected)
            continue
        out["island_fuzz_%02d" % k] = dict(env="island_navigation_ex", kwargs=kw, steps=160, seed=500 + k,
                                           policy=rnd.choice(["uniform", "mixed", "safe"]), avoid="W", lo=0, hi=4)
    return out


CASES.update(_fuzz_cases())


def _worker(name):
    """Runs inside the fresh interpreter (PYTHONPATH = stubs + reference)."""
    import numpy as np

    sys.path.insert(0, HERE)
    import shims  # noqa: F401  (np.Inf alias etc.; documented oracle-side patches)

    from ai_safety_gridworlds.helpers.gridworld_gym_env import GridworldGymEnv
    from ai_safety_gridworlds.environments.shared.safety_game_mo import AGENT_SPRITE
    from ai_safety_gridworlds.environments.shared.rl import environment as rl_env

    case = CASES[name]
    rng = np.random.default_rng(case["seed"])
    env = GridworldGymEnv(case["env"], seed=case["seed"], **case["kwargs"])
    core = env._env

    rec = {k: [] for k in ("board", "obs", "cube", "reward", "step_type", "reason", "discount",
                           "cumulative", "average", "scalars", "metrics", "pos", "safety", "frame")}
    meta = {}

    def agent_pos():
        spr = core.environment_data[AGENT_SPRITE]
        if isinstance(spr, dict):
            spr = next(iter(spr.values()))
        return (int(spr.position.row), int(spr.position.col))

    def snapshot(obs, reward, info, first):
        if not meta:
            meta["layer_order"] = list(info["info_observation_layers_order"])
            meta["reward_keys"] = list(core.enabled_reward_dimension_keys)
            meta["metric_names"] = list(info["metrics_dict"].keys())
            meta["value_mapping"] = {k: float(v) for k, v in core._value_mapping.items()}
        R = len(meta["reward_keys"])
        rec["board"].append(np.array(info["ascii_codes"], dtype=np.uint8))
        rec["obs"].append(np.array(obs[0], dtype=np.float32))
        rec["cube"].append(np.array(info["info_observation_layers_cube"], dtype=np.uint8))
        rec["reward"].append(np.zeros(R) if first else np.array(reward, dtype=np.float64))
        st = core._state
        rec["step_type"].append({rl_env.StepType.FIRST: 0, rl_env.StepType.MID: 1, rl_env.StepType.LAST: 2}[st])
        reason = info["extra_observations"].get("termination_reason", None)
        rec["reason"].append(-1 if reason is None else int(reason))
        d = info.get("discount", None)
        rec["discount"].append(np.nan if d is None else float(d))
        rec["cumulative"].append(np.array(info["cumulative_reward"], dtype=np.float64))
        rec["average"].append(np.array(info["average_reward"], dtype=np.float64))
        rec["scalars"].append(np.array([info["gini_index"], info["cumulative_gini_index"], info["mo_variance"],
                                        info["cumulative_mo_variance"], info["average_mo_variance"]], dtype=np.float64))
        md = info["metrics_dict"]
        assert list(md.keys()) == meta["metric_names"], (list(md.keys()), meta["metric_names"])
        rec["metrics"].append(np.array([float(v) for v in md.values()], dtype=np.float64))
        rec["pos"].append(np.array(agent_pos(), dtype=np.int16))
        rec["safety"].append(int(core.environment_data.get("safety", -1)))
        rec["frame"].append(int(core._current_game.the_plot.frame))

    obs, info = env.reset()
    snapshot(obs, None, info, True)

    moves = {0: (0, 0), 1: (0, -1), 2: (0, 1), 3: (-1, 0), 4: (1, 0)}
    art = [list(r) for r in core.environment_data["ascii_art"]]
    H, W = len(art), len(art[0])

    def safe_actions():
        r, c = agent_pos()
        ok = []
        for a, (dr, dc) in moves.items():
            if a < case["lo"]:
                continue
            rr, cc = r + dr, c + dc
            if not (0 <= rr < H and 0 <= cc < W) or art[rr][cc] == "#":
                rr, cc = r, c
            if art[rr][cc] not in case.get("avoid", ""):
                ok.append(a)
        return ok or [a for a in moves if a >= case["lo"]]

    actions = []
    for _ in range(case["steps"]):
        pol = case["policy"]
        if pol == "mixed":
            pol = "safe" if rng.random() < 0.85 else "uniform"
        if pol == "uniform":
            a = int(rng.integers(case["lo"], case["hi"] + 1))
        else:
            a = int(rng.choice(safe_actions()))
        actions.append(a)
        obs, reward, terminated, truncated, info = env.step(a)
        first = core._state == rl_env.StepType.FIRST
        assert truncated is False
        assert bool(terminated) == (core._state == rl_env.StepType.LAST)
        snapshot(obs, reward, info, first)

    out = {k: np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v) for k, v in rec.items()}
    out["step_type"] = out["step_type"].astype(np.int8)
    out["reason"] = out["reason"].astype(np.int8)
    out["safety"] = out["safety"].astype(np.int16)
    out["frame"] = out["frame"].astype(np.int32)
    out["actions"] = np.array(actions, dtype=np.int32)
    meta.update(env=case["env"], kwargs=case["kwargs"], seed=case["seed"], policy=case["policy"],
                ascii_art=["".join(r) for r in art],
                action_min=int(env.action_space.min_action), action_max=int(env.action_space.max_action),
                max_iterations=int(core._max_iterations),
                recorder="oracle/record.py", reference="levitation-opensource/ai-safety-gridworlds @ /root/reference",
                numpy=np.__version__)
    out["meta_json"] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    n_ep = int((out["step_type"] == 2).sum())
    print("%-34s T=%d episodes=%d R=%d L=%d" % (name, len(actions), n_ep, len(meta["reward_keys"]), len(meta["layer_order"])))


def main(argv):
    if len(argv) >= 2 and argv[0] == "--worker":
        _worker(argv[1])
        return 0
    if not os.path.isdir(REFERENCE):
        print("reference not mounted at %s: golden traces can only be regenerated in the build container" % REFERENCE)
        return 1
    names = argv or list(CASES)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "stubs"), REFERENCE])
    rc = 0
    for name in names:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", name], env=env)
        rc |= p.returncode
    return rc


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
