"""EnvSpec: one environment TYPE compiled to the POD the CUDA library consumes.

This is the host-side "config compiler" (SURVEY.md section 7 step 2): it turns a game name, level
and the reference's flag names into a `GwConfig` plus the Python-visible metadata the reference
exposes (sorted reward-dimension keys, layer order, metric names, action range, value mapping).
"""
import ast
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

from .. import _abi


def parse_reward(value) -> Dict[str, float]:
    """Accepts what the reference's reward flags accept: a dict, or the `str(mo_reward)` /
    `str(dict)` text form parsed with literal_eval (shared/mo_reward.py:109-117)."""
    if isinstance(value, dict):
        return dict(value)
    if value is None or value == "":
        return {}
    if isinstance(value, str):
        obj = ast.literal_eval(value)
        if not isinstance(obj, dict):
            raise ValueError("reward flag must be a dict literal, got %r" % (value,))
        return obj
    raise TypeError("cannot interpret %r as a multi-objective reward" % (value,))


def enabled_reward_keys(enabled_rewards: List[Dict[str, float]]) -> List[str]:
    """mo_reward.get_enabled_reward_dimension_keys (shared/mo_reward.py:120-146): the union of the
    keys whose unit value is non-zero, sorted."""
    keys = set()
    for r in enabled_rewards:
        keys |= {k for k, v in r.items() if v != 0}
    return sorted(keys)


def dense_reward(reward: Dict[str, float], keys: List[str], event_name: str, can_fire: bool) -> List[float]:
    """mo_reward.tolist (shared/mo_reward.py:184-203): a posted reward with a non-zero value in a
    dimension that is not enabled raises ValueError.  The reference raises it at the first step
    that posts the reward; the batched engine cannot raise per environment, so the same error is
    raised here, eagerly, for every event that the flags and the map make reachable."""
    for k, v in reward.items():
        if v != 0 and k not in keys and can_fire:
            raise ValueError("Reward %s is not enabled but is still included in mo_reward with nonzero value" % k)
    return [float(reward.get(k, 0)) for k in keys]


def map_contains(ch: str, art: List[str]) -> bool:
    """shared/safety_ui_ex.py:662-666"""
    return any(ch in row for row in art)


@dataclass
class EnvSpec:
    name: str
    config: _abi.GwConfig
    art: List[str]
    height: int
    width: int
    reward_keys: List[str]
    layer_order: List[str]
    metric_names: List[str]
    value_mapping: Dict[str, float]
    action_range: Tuple[int, int]
    flags: Dict[str, object] = field(default_factory=dict)

    @property
    def n_rewards(self):
        return len(self.reward_keys)

    @property
    def n_layers(self):
        return len(self.layer_order)

    @property
    def cells(self):
        return self.height * self.width

    def with_autoreset(self, mode: int) -> "EnvSpec":
        cfg = _abi.GwConfig()
        C.memmove(C.byref(cfg), C.byref(self.config), C.sizeof(cfg))
        cfg.autoreset_mode = int(mode)
        return EnvSpec(self.name, cfg, self.art, self.height, self.width, self.reward_keys, self.layer_order,
                       self.metric_names, self.value_mapping, self.action_range, dict(self.flags))


def fill_common(cfg: _abi.GwConfig, env_type: int, art: List[str], layer_order: List[str],
                value_mapping: Dict[str, float], max_iterations: int, n_rewards: int, autoreset_mode: int):
    height, width = len(art), len(art[0])
    if any(len(r) != width for r in art):
        raise ValueError("ragged game art")
    if height * width > _abi.GW_MAX_CELLS:
        raise ValueError("board %dx%d exceeds GW_MAX_CELLS=%d" % (height, width, _abi.GW_MAX_CELLS))
    if len(layer_order) > _abi.GW_MAX_LAYERS:
        raise ValueError("too many layers")
    if n_rewards > _abi.GW_MAX_REWARDS:
        raise ValueError("too many reward dimensions")
    if not (1 <= int(max_iterations) <= 65535):
        raise ValueError("max_iterations must be in 1..65535")
    if sum(row.count("A") for row in art) != 1:
        raise ValueError("game art must contain exactly one agent 'A'")
    cfg.abi_version = _abi.GW_ABI_VERSION
    cfg.env_type = env_type
    cfg.height, cfg.width = height, width
    cfg.n_layers = len(layer_order)
    cfg.n_rewards = n_rewards
    cfg.max_iterations = int(max_iterations)
    cfg.autoreset_mode = int(autoreset_mode)
    flat = "".join(art)
    for i, ch in enumerate(flat):
        cfg.art[i] = ord(ch)
    for i, ch in enumerate(layer_order):
        cfg.layer_chars[i] = ord(ch)
    for ch, v in value_mapping.items():
        cfg.value_map[ord(ch)] = float(v)
    return height, width
