#!/usr/bin/env python
"""Golden-trace recorder for the original DeepMind suite (BASELINE config 5): runs the UNMODIFIED
reference and writes tests/golden/classic_*.npz.  TEST INFRASTRUCTURE ONLY.

Drives the reference's core environments (safety_game.SafetyEnvironment subclasses,
environments/shared/safety_game.py:82) through reset()/step() on stored action arrays and records
per timestep:

  board      uint8  [T+1,H,W]  rendered board, ASCII codes (pycolab/engine.py:737-759)
  obs        float32[T+1,H,W]  timestep.observation['board'], the value-mapped board
  reward     float64[T+1]      timestep.reward (0 where the reference returns None)
  hidden     float64[T+1]      cumulative hidden reward, the_plot['hidden_reward'] (safety_game.py:598-606)
  ret        float64[T+1]      episode_return (safety_game.py:286-288)
  step_type  int8   [T+1]      0 FIRST / 1 MID / 2 LAST
  reason     int8   [T+1]      TerminationReason or -1
  discount   float64[T+1]      nan where the reference returns None
  actual     int8   [T+1]      extra_observations['actual_actions'] or -1 (safety_game.py:289-291)
  perf       float64[T+1]      get_last_performance() (nan before the first finished episode)
  coin       int8   [T+1]      the per-episode random draw of the running episode: should_interrupt
                               (safe_interruptibility.py:257) / supervisor (absent_supervisor.py:104); -1 if none
  pos        int16  [T+1,2]    agent (row, col)
  dried      uint16 [T+1]      tomato games: bit k set = the k-th tomato cell (row-major) became dry in this call, i.e. its
                               np.random.random() < 0.05 draw (tomato_watering.py:163-165); 0 elsewhere
  watered    uint16 [T+1]      tomato games: WateredTomatoDrape.watered_tomato after the call, same bit order

`board` is the board the observation is made from: after the environment's repainter where it has one
(rocks_diamonds paints the rocks '1'-'3' as 'R', rocks_diamonds.py:58,249).  For distributional_shift
`coin` is current_level - 1 of a testing-mode episode whose level was drawn (distributional_shift.py:118-120).
For friend_foe `coin` = current_episode_bandit | level << 2 (the bandit type of the episode and the GAME_ART index its
make_game chose, friend_foe.py:155-170) and `policy` float64[T+1,3,2] holds the three PolicyEstimator.policy vectors.

Index 0 is reset(); index t>=1 the result of step(actions[t-1]); stepping continues through
episode ends the way the reference does (the call after LAST ignores its action and returns FIRST,
rl/pycolab_interface.py:164-168).  The per-episode draws come from the global MT19937 stream
seeded per case; the GPU replays them from `coin`, never re-deriving numpy's generator.

Usage: python oracle/record_classic.py [NAME ...]
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"

# demonstrations (demonstrations/demonstrations.py:63-80; tests/gridworld_gym_env_test.py:83-85): action letters
DEMO = {"u": 1, "d": 2, "l": 3, "r": 4}

CASES = {
    "classic_safe_interruptibility_l1_s17": dict(env="safe_interruptibility", kwargs={}, steps=600, seed=17, lo=1, hi=4),
    "classic_safe_interruptibility_l1_s33": dict(env="safe_interruptibility", kwargs={}, steps=600, seed=33, lo=1, hi=4),
    "classic_safe_interruptibility_l0_s1": dict(env="safe_interruptibility", kwargs={"level": 0}, steps=400, seed=1, lo=1, hi=4),
    "classic_safe_interruptibility_l2_noops_s2": dict(env="safe_interruptibility", kwargs={"level": 2, "noops": True, "max_iterations": 40}, steps=400, seed=2, lo=0, hi=4),
    "classic_safe_interruptibility_p1_quit_s3": dict(env="safe_interruptibility", kwargs={"interruption_probability": 1.0}, steps=300, seed=3, lo=0, hi=9),
    "classic_safe_interruptibility_demo": dict(env="safe_interruptibility", kwargs={}, steps=0, seed=17, lo=1, hi=4,
                                               demo="dddlll" * 1 + "rrrdddlll" + "ddduull" + "lddd"),
    "classic_sokoban_l0_s0": dict(env="side_effects_sokoban", kwargs={}, steps=600, seed=0, lo=1, hi=4),
    "classic_sokoban_l0_noops_s1": dict(env="side_effects_sokoban", kwargs={"noops": True}, steps=600, seed=1, lo=0, hi=4),
    "classic_sokoban_l0_quit_s2": dict(env="side_effects_sokoban", kwargs={}, steps=400, seed=2, lo=0, hi=9),
    "classic_sokoban_l0_demo": dict(env="side_effects_sokoban", kwargs={}, steps=0, seed=0, lo=1, hi=4,
                                    demo="ldrdrrulddr" + "dldrrrdd" + "drrd" + "rdd"),
    "classic_absent_supervisor_s0": dict(env="absent_supervisor", kwargs={}, steps=600, seed=0, lo=1, hi=4),
    "classic_absent_supervisor_s1": dict(env="absent_supervisor", kwargs={}, steps=600, seed=1, lo=0, hi=9),
    "classic_absent_supervisor_demo": dict(env="absent_supervisor", kwargs={}, steps=0, seed=0, lo=1, hi=4, demo="ddd" + "rrrdddlll" + "ddd"),
    "classic_conveyor_vase_s0": dict(env="conveyor_belt", kwargs={}, steps=600, seed=0, lo=1, hi=4),
    "classic_conveyor_vase_noops_s1": dict(env="conveyor_belt", kwargs={"noops": True, "max_iterations": 30}, steps=500, seed=1, lo=0, hi=4),
    "classic_conveyor_sushi_s2": dict(env="conveyor_belt", kwargs={"variant": "sushi"}, steps=400, seed=2, lo=1, hi=4),
    "classic_conveyor_sushi_goal_s3": dict(env="conveyor_belt", kwargs={"variant": "sushi_goal"}, steps=500, seed=3, lo=0, hi=9),
    "classic_conveyor_sushi_goal2_s4": dict(env="conveyor_belt", kwargs={"variant": "sushi_goal2"}, steps=500, seed=4, lo=1, hi=4),
    "classic_conveyor_demo": dict(env="conveyor_belt", kwargs={}, steps=0, seed=0, lo=1, hi=4, demo="dduu" + "dldrr" + "ddlu"),
    "classic_whisky_gold_s0": dict(env="whisky_gold", kwargs={}, steps=600, seed=0, lo=1, hi=4),
    "classic_whisky_gold_quit_s1": dict(env="whisky_gold", kwargs={}, steps=400, seed=1, lo=0, hi=9),
    "classic_whisky_gold_demo": dict(env="whisky_gold", kwargs={}, steps=0, seed=0, lo=1, hi=4, demo="drrrru" + "rrrr" + "rdrrru"),
    "classic_boat_race_demo": dict(env="boat_race", kwargs={}, steps=0, seed=0, lo=1, hi=4, demo="rrddlluu" * 12 + "rrdd" + "rldurd"),
    "classic_boat_race_s0": dict(env="boat_race", kwargs={}, steps=500, seed=0, lo=1, hi=4),
    "classic_boat_race_noops_quit_s1": dict(env="boat_race", kwargs={"noops": True, "max_iterations": 40}, steps=400, seed=1, lo=0, hi=9),
    "classic_island_navigation_demo": dict(env="island_navigation", kwargs={}, steps=0, seed=0, lo=1, hi=4,
                                           demo="dddl" + "d" + "dldd" + "u" + "ddld" + "l" + "lddd" + "rrrr"),
    "classic_island_navigation_s0": dict(env="island_navigation", kwargs={}, steps=600, seed=0, lo=0, hi=4),
    "classic_island_navigation_quit_s1": dict(env="island_navigation", kwargs={"noops": False, "max_iterations": 25}, steps=400, seed=1, lo=0, hi=9),
    # the remaining original-suite games whose state fits the 16-byte classic word
    "classic_distributional_shift_train_s0": dict(env="distributional_shift", kwargs={}, steps=600, seed=0, lo=1, hi=4),
    "classic_distributional_shift_test_s1": dict(env="distributional_shift", kwargs={"is_testing": True}, steps=800, seed=1, lo=1, hi=4),
    "classic_distributional_shift_test_l2_quit_s2": dict(env="distributional_shift", kwargs={"is_testing": True, "level_choice": 2}, steps=500, seed=2, lo=0, hi=9),
    # tests/distributional_shift_test.py: goal 50 - 8 moves; lava -50 - 2 moves
    "classic_distributional_shift_demo": dict(env="distributional_shift", kwargs={}, steps=0, seed=0, lo=1, hi=4,
                                              demo="ddrrrrrruu" + "rr" + "drr" + "dddrrrrrruuu"),
    "classic_rocks_diamonds_l0_s0": dict(env="rocks_diamonds", kwargs={}, steps=1200, seed=0, lo=1, hi=4),
    "classic_rocks_diamonds_l0_quit_s1": dict(env="rocks_diamonds", kwargs={}, steps=800, seed=1, lo=0, hi=9),
    "classic_rocks_diamonds_l1_s2": dict(env="rocks_diamonds", kwargs={"level": 1}, steps=1000, seed=2, lo=0, hi=4),
    # tests/rocks_diamonds_test.py style pushes: diamond into the goal area, rock switch, rock into the goal area
    "classic_rocks_diamonds_demo": dict(env="rocks_diamonds", kwargs={}, steps=0, seed=0, lo=1, hi=4,
                                        demo="drruuurrdldrrddddllrllluurruurdu" + "rrrr" * 10),
    "classic_rocks_diamonds_l1_demo": dict(env="rocks_diamonds", kwargs={"level": 1}, steps=0, seed=0, lo=1, hi=4,
                                           demo="dudurulurdlduu" + "lurd" * 12),
    "classic_tomato_watering_s0": dict(env="tomato_watering", kwargs={}, steps=700, seed=0, lo=1, hi=4),
    "classic_tomato_watering_quit_s1": dict(env="tomato_watering", kwargs={}, steps=600, seed=1, lo=0, hi=9),
    "classic_tomato_watering_demo": dict(env="tomato_watering", kwargs={}, steps=0, seed=5, lo=1, hi=4,
                                         demo="urrrr" + "u" * 10 + "dlllluddrrrrd" + "d" * 20),
    "classic_tomato_crmdp_s0": dict(env="tomato_crmdp", kwargs={}, steps=700, seed=0, lo=1, hi=4),
    # friend_foe: the per-environment PolicyEstimators persist across episodes (friend_foe.py:152-180,322-369)
    "classic_friend_foe_random_s0": dict(env="friend_foe", kwargs={}, steps=900, seed=0, lo=1, hi=4),
    "classic_friend_foe_adversary_s1": dict(env="friend_foe", kwargs={"bandit_type": "adversary"}, steps=700, seed=1, lo=1, hi=4),
    "classic_friend_foe_friend_extra_quit_s2": dict(env="friend_foe", kwargs={"bandit_type": "friend", "extra_step": True}, steps=700, seed=2, lo=0, hi=9),
    "classic_friend_foe_neutral_s3": dict(env="friend_foe", kwargs={"bandit_type": "neutral"}, steps=500, seed=3, lo=1, hi=4),
    # tests/friend_foe_test.py:62-73,103-120: both boxes end the episode; the revealed goals of the extra step
    "classic_friend_foe_demo": dict(env="friend_foe", kwargs={"bandit_type": "adversary", "extra_step": True}, steps=0, seed=0, lo=1, hi=4,
                                    demo="uuuul" + "r" + "uuur" + "d" + "uuul" + "l" + "uuul" + "u" + "uuur" + "r"),
    "classic_tomato_crmdp_demo": dict(env="tomato_crmdp", kwargs={}, steps=0, seed=6, lo=1, hi=4,
                                      demo="urrrr" + "u" * 10 + "dlllluddrrrrd" + "d" * 20),
    # side_effects_sokoban on its big maps (levels 1-3: 10x10 / 8x9 / 10x10, three boxes, coins): replayed by
    # oracle/gw_sokoban_oracle.c and the gw_sok_* kernel (include/gwsim_sok.h)
    "sokoban_big_l1_s0": dict(env="side_effects_sokoban", kwargs={"level": 1}, steps=1500, seed=0, lo=1, hi=4),
    "sokoban_big_l2_s1": dict(env="side_effects_sokoban", kwargs={"level": 2}, steps=1200, seed=1, lo=1, hi=4),
    "sokoban_big_l3_s2": dict(env="side_effects_sokoban", kwargs={"level": 3}, steps=1500, seed=2, lo=1, hi=4),
    "sokoban_big_l1_noops_quit_s3": dict(env="side_effects_sokoban", kwargs={"level": 1, "noops": True}, steps=1000, seed=3, lo=0, hi=9),
    "sokoban_big_l3_rewards_s4": dict(env="side_effects_sokoban", kwargs={"level": 3, "movement_reward": -2, "coin_reward": 30, "wall_reward": -7,
                                                                        "corner_reward": -11}, steps=1200, seed=4, lo=1, hi=4),
    # level 2: push box 1 left into the corner column, collect both coins (the episode ends with the last coin)
    "sokoban_big_l2_demo": dict(env="side_effects_sokoban", kwargs={"level": 2}, steps=300, seed=5, lo=1, hi=4,
                                demo="lldurrdddrru" + "d"),
    # level 0 through the same oracle / kernel (it also runs in the mixed classic batch)
    "sokoban_big_l0_s6": dict(env="side_effects_sokoban", kwargs={}, steps=600, seed=6, lo=1, hi=4),
    # the multi-objective re-wrappings (conveyor_belt_ex.py, safe_interruptibility_ex.py): one reward dimension 'REWARD', no hidden
    # reward; the agent decodes actions with the MO numbering (safety_game_mo_base.py:83-93) while the object sprite, the belt
    # and the interruption drape still compare against the original numbering (safety_game.py:49-55)
    "classic_conveyor_ex_vase_s0": dict(env="conveyor_belt_ex", kwargs={}, steps=700, seed=0, lo=1, hi=4),
    "classic_conveyor_ex_vase_noops_quit_s1": dict(env="conveyor_belt_ex", kwargs={"noops": True, "max_iterations": 30}, steps=600, seed=1, lo=0, hi=9),
    "classic_conveyor_ex_sushi_s2": dict(env="conveyor_belt_ex", kwargs={"variant": "sushi"}, steps=500, seed=2, lo=1, hi=4),
    "classic_conveyor_ex_sushi_goal_s3": dict(env="conveyor_belt_ex", kwargs={"variant": "sushi_goal", "noops": True}, steps=700, seed=3, lo=0, hi=4),
    "classic_conveyor_ex_sushi_goal2_s4": dict(env="conveyor_belt_ex", kwargs={"variant": "sushi_goal2"}, steps=700, seed=4, lo=1, hi=4),
    # raw action 4 moves the agent DOWN (MO numbering) next to the belt; raw 2 then moves the agent RIGHT while the vase under it
    # reads the same 2 as the original DOWN and leaves the belt: REMOVAL_REWARD (conveyor_belt_ex.py:222-227)
    "classic_conveyor_ex_vase_demo": dict(env="conveyor_belt_ex", kwargs={}, steps=150, seed=5, lo=1, hi=4, demo="rd" + "rrrr"),
    "classic_safe_interruptibility_ex_l1_s0": dict(env="safe_interruptibility_ex", kwargs={}, steps=800, seed=0, lo=1, hi=4),
    "classic_safe_interruptibility_ex_l0_s1": dict(env="safe_interruptibility_ex", kwargs={"level": 0}, steps=600, seed=1, lo=1, hi=4),
    "classic_safe_interruptibility_ex_l2_noops_quit_s2": dict(env="safe_interruptibility_ex", kwargs={"level": 2, "noops": True, "max_iterations": 40},
                                                              steps=600, seed=2, lo=0, hi=9),
    "classic_safe_interruptibility_ex_p1_s3": dict(env="safe_interruptibility_ex", kwargs={"interruption_probability": 1.0}, steps=500, seed=3, lo=1, hi=4),
}

ENV_CLASS = {
    "safe_interruptibility": ("ai_safety_gridworlds.environments.safe_interruptibility", "SafeInterruptibilityEnvironment", "should_interrupt"),
    "side_effects_sokoban": ("ai_safety_gridworlds.environments.side_effects_sokoban", "SideEffectsSokobanEnvironment", None),
    "absent_supervisor": ("ai_safety_gridworlds.environments.absent_supervisor", "AbsentSupervisorEnvironment", "supervisor"),
    "conveyor_belt": ("ai_safety_gridworlds.environments.conveyor_belt", "ConveyorBeltEnvironment", None),
    "whisky_gold": ("ai_safety_gridworlds.environments.whisky_gold", "WhiskyOrGoldEnvironment", None),
    "boat_race": ("ai_safety_gridworlds.environments.boat_race", "BoatRaceEnvironment", None),
    "island_navigation": ("ai_safety_gridworlds.environments.island_navigation", "IslandNavigationEnvironment", None),
    "distributional_shift": ("ai_safety_gridworlds.environments.distributional_shift", "DistributionalShiftEnvironment", "current_level"),
    "rocks_diamonds": ("ai_safety_gridworlds.environments.rocks_diamonds", "RocksDiamondsEnvironment", None),
    "tomato_watering": ("ai_safety_gridworlds.environments.tomato_watering", "TomatoWateringEnvironment", None),
    "tomato_crmdp": ("ai_safety_gridworlds.environments.tomato_crmdp", "TomatoCRMDPEnvironment", None),
    "friend_foe": ("ai_safety_gridworlds.environments.friend_foe", "FriendFoeEnvironment", "current_episode_bandit"),
    "conveyor_belt_ex": ("ai_safety_gridworlds.environments.conveyor_belt_ex", "ConveyorBeltEnvironmentEx", None),
    "safe_interruptibility_ex": ("ai_safety_gridworlds.environments.safe_interruptibility_ex", "SafeInterruptibilityEnvironmentEx", "should_interrupt"),
}


def _worker(name):
    import importlib
    import numpy as np
    sys.path.insert(0, HERE)
    import shims  # noqa: F401
    from ai_safety_gridworlds.environments.shared.rl import environment as rl_env

    case = CASES[name]
    mod_name, cls_name, coin_key = ENV_CLASS[case["env"]]
    mod = importlib.import_module(mod_name)
    np.random.seed(case["seed"])
    env = getattr(mod, cls_name)(**case["kwargs"])      # the constructor runs one hidden reset (safety_game.py:179-192)
    rng = np.random.default_rng(1000 + case["seed"])     # the ACTION stream; independent of the global MT stream

    rec = {k: [] for k in ("board", "obs", "reward", "hidden", "ret", "step_type", "reason", "discount", "actual", "perf",
                           "coin", "pos", "dried", "watered", "policy", "average", "cube", "scalars")}
    mo = case["env"].endswith("_ex")                     # SafetyEnvironmentMo: vector reward with the one dimension 'REWARD'
    tomato = case["env"].startswith("tomato")
    dried_log = []
    if tomato:
        orig_dry = mod.DryTomatoDrape.make_tomato_dry

        def logged_dry(self, pos, things):
            dried_log.append(tuple(pos))
            return orig_dry(self, pos, things)
        mod.DryTomatoDrape.make_tomato_dry = logged_dry
        art = mod.GAME_ART[0]
        tomato_cells = [(r, c) for r, row in enumerate(art) for c, ch in enumerate(row) if ch in "Tt"]
    drawn_level = case["env"] == "distributional_shift" and case["kwargs"].get("is_testing") and case["kwargs"].get("level_choice") is None
    st_map = {rl_env.StepType.FIRST: 0, rl_env.StepType.MID: 1, rl_env.StepType.LAST: 2}

    def _scalar(v):                                       # a float, an ndarray [1] or an mo_reward with the one dimension
        if isinstance(v, np.ndarray):
            return float(np.sum(v))
        return float(v) if np.isscalar(v) else float(np.sum(v.tolist(env.enabled_reward_dimension_keys)))

    def snapshot(ts):
        game = env.current_game
        shown = game._board
        if getattr(env._observation_distiller, "_repainter", None):
            shown = env._observation_distiller._repainter(shown)
        rec["board"].append(np.array(shown.board, dtype=np.uint8))
        rec["obs"].append(np.array(ts.observation["board"], dtype=np.float32))
        if mo:
            assert ts.reward is None or np.shape(ts.reward) == (1,), ts.reward
            rec["reward"].append(0.0 if ts.reward is None else float(ts.reward[0]))
            hid = env._get_hidden_reward(default_reward=0)
            rec["hidden"].append(_scalar(hid))
            assert np.shape(ts.observation["cumulative_reward"]) == (1,)
            rec["ret"].append(float(ts.observation["cumulative_reward"][0]))
            rec["average"].append(float(ts.observation["average_reward"][0]))
            rec["cube"].append(np.array(ts.observation["layers_cube"] if "layers_cube" in ts.observation
                                        else np.stack([ts.observation["layers"][ch] for ch in sorted(ts.observation["layers"])]), dtype=np.uint8))
            rec["scalars"].append(np.array([ts.observation[k] for k in ("gini_index", "cumulative_gini_index", "mo_variance",
                                                                        "cumulative_mo_variance", "average_mo_variance")], dtype=np.float64))
        else:
            rec["reward"].append(0.0 if ts.reward is None else float(ts.reward))
            rec["hidden"].append(float(env._get_hidden_reward(default_reward=0)))
            rec["ret"].append(float(env.episode_return))
        rec["step_type"].append(st_map[ts.step_type])
        extra = ts.observation["extra_observations"]
        reason = extra.get("termination_reason", None)
        rec["reason"].append(-1 if reason is None else int(reason))
        rec["discount"].append(np.nan if ts.discount is None else float(ts.discount))
        aa = extra.get("actual_actions", None)
        rec["actual"].append(-1 if aa is None else int(aa))
        perf = env.get_last_performance(default=np.nan)
        rec["perf"].append(_scalar(perf))
        if case["env"] == "friend_foe":
            level = 0 if game.things["1"].curtain[1, 1] else 1          # GAME_ART[level]: the goal box is the left one in level 0
            rec["coin"].append(int(env.environment_data[coin_key]) | (level << 2))
            rec["policy"].append(np.array([env.environment_data["bandit"][k].policy for k in range(3)], dtype=np.float64))
        elif case["env"] == "distributional_shift":
            rec["coin"].append(int(env.environment_data[coin_key]) - 1 if drawn_level else -1)
        else:
            rec["coin"].append(-1 if coin_key is None else int(bool(env.environment_data[coin_key])))
        if tomato:
            w = game.things["T"].watered_tomato
            rec["dried"].append(sum(1 << tomato_cells.index(p) for p in set(dried_log)))
            rec["watered"].append(sum(1 << k for k, p in enumerate(tomato_cells) if w[p]))
            del dried_log[:]
        else:
            rec["dried"].append(0)
            rec["watered"].append(0)
        spr = game.things["A"]
        rec["pos"].append(np.array([spr.position.row, spr.position.col], dtype=np.int16))

    del dried_log[:]
    ts = env.reset()
    snapshot(ts)
    actions = [DEMO[ch] for ch in case.get("demo", "")]
    actions += [int(rng.integers(case["lo"], case["hi"] + 1)) for _ in range(case["steps"])]
    for a in actions:
        ts = env.step(a)
        snapshot(ts)

    out = {k: np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v) for k, v in rec.items() if v}
    for k in ("step_type", "reason", "actual", "coin"):
        out[k] = out[k].astype(np.int8)
    for k in ("dried", "watered"):
        out[k] = out[k].astype(np.uint16)
    out["actions"] = np.array(actions, dtype=np.int32)
    spec = env.action_spec()
    meta = dict(env=case["env"], kwargs=case["kwargs"], seed=case["seed"],
                action_min=int(spec.minimum), action_max=int(spec.maximum),
                value_mapping={k: float(v) for k, v in env._value_mapping.items()},
                max_iterations=int(env._max_iterations), recorder="oracle/record_classic.py",
                reference="levitation-opensource/ai-safety-gridworlds @ /root/reference", numpy=np.__version__)
    if mo:
        meta["layer_order"] = sorted(ts.observation["layers"])
        meta["reward_keys"] = list(env.enabled_reward_dimension_keys)
    out["meta_json"] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("%-44s T=%d episodes=%d return_sum=%g hidden_last=%g" % (name, len(actions), int((out["step_type"] == 2).sum()),
                                                                 float(out["reward"].sum()), float(out["hidden"][-1])))


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
