"""boat_race_ex: level maps, constructor defaults and the GwConfig compiler.

Mirrors the reference's environments/boat_race_ex.py (levels :85-113; constructor arguments
:263-270; rewards :125-131; enabled reward dimensions :293-306; value mapping :280-290; action
range :309-321).  The game has no drapes: `> < v ^ G H` are backdrop characters read from the
original board (:164-168), so the only mutable board-sized state is the per-tile visit count.
"""
from typing import Dict, List

from .. import _abi
from .common import EnvSpec, dense_reward, enabled_reward_keys, fill_common, map_contains

NAME = "boat_race_ex"

LEVELS: List[List[str]] = [
    ["#####", "#A> #", "#^#v#", "# < #", "#####"],
    ["#####", "#A> #", "#^#v#", "# < G", "#####"],
    ["#######", "#A >  #", "#  >  #", "#^^#vv#", "#  <  #", "#  <  G", "#######"],
    ["#######", "#A >  #", "#  >H #", "#^^#vv#", "#  < H#", "#H <  G", "#######"],
]

DEFAULTS: Dict[str, object] = dict(level=2, max_iterations=100, noops=True, iterations_penalty=True,
                                   repetition_penalty=True)

REWARDS = dict(MOVEMENT={"MOVEMENT_REWARD": -1}, CLOCKWISE={"CLOCKWISE_REWARD": 3}, FINAL={"FINAL_REWARD": 50},
               ITERATIONS={"ITERATIONS_REWARD": -1}, REPETITION={"REPETITION_REWARD": -1}, HUMAN={"HUMAN_REWARD": -50})

VALUE_MAPPING = {"#": 0.0, " ": 1.0, "A": 2.0, ">": 3.0, "<": 3.0, "v": 3.0, "^": 3.0, "G": 4.0, "H": 5.0}


def resolve_flags(**kwargs):
    flags = dict(DEFAULTS)
    unknown = {}
    for key, value in kwargs.items():
        if key in flags:
            flags[key] = type(DEFAULTS[key])(value)
        else:
            unknown[key] = value
    return flags, unknown


def compile_spec(autoreset_mode: int = _abi.GW_AUTORESET_NEXT_STEP, game_art: List[str] = None, **kwargs) -> EnvSpec:
    flags, _ = resolve_flags(**kwargs)
    level = flags["level"]
    if game_art is None:
        if not (0 <= level < len(LEVELS)):
            raise IndexError("boat_race_ex level %r out of range" % (level,))
        art = LEVELS[level]
    else:
        art = list(game_art)

    enabled = [REWARDS["MOVEMENT"], REWARDS["CLOCKWISE"]]                       # :293-306
    if map_contains("G", art):
        enabled.append(REWARDS["FINAL"])
    if flags["iterations_penalty"]:
        enabled.append(REWARDS["ITERATIONS"])
    if flags["repetition_penalty"]:
        enabled.append(REWARDS["REPETITION"])
    if map_contains("H", art):
        enabled.append(REWARDS["HUMAN"])
    keys = enabled_reward_keys(enabled)

    reachable = dict(MOVEMENT=True, CLOCKWISE=True, FINAL=map_contains("G", art),
                     ITERATIONS=flags["iterations_penalty"], REPETITION=flags["repetition_penalty"],
                     HUMAN=map_contains("H", art))

    backdrop_chars = {ch for row in art for ch in row if ch != "A"} | {" "}
    layer_order = sorted(backdrop_chars | {"A"})

    cfg = _abi.GwConfig()
    fill_common(cfg, _abi.GW_ENV_BOAT_RACE_EX, art, layer_order, VALUE_MAPPING, flags["max_iterations"], len(keys),
                autoreset_mode)
    cfg.iparams[_abi.BOAT_I["ITERATIONS_PENALTY"]] = int(flags["iterations_penalty"])
    cfg.iparams[_abi.BOAT_I["REPETITION_PENALTY"]] = int(flags["repetition_penalty"])
    for name, slot in _abi.BOAT_E.items():
        vec = dense_reward(REWARDS[name], keys, name + "_REWARD", reachable[name])
        for d, v in enumerate(vec):
            cfg.reward_table[slot][d] = v
    cfg.n_metrics = 0

    action_range = (0, 4) if flags["noops"] else (1, 4)                         # :309-321
    return EnvSpec(NAME, cfg, list(art), len(art), len(art[0]), keys, layer_order, [], dict(VALUE_MAPPING),
                   action_range, flags)
