"""Checkpoint / resume of the batched environments.

The reference makes its environments picklable (safety_game_mo.py:406-419, safety_game_moma.py:414-427) so that a run can be
moved between processes.  Here everything a batch's future depends on is (a) the caller-owned state blob -- plus, for the games
with per-environment layouts, the maps / resource tensors -- and (b) the handle's call counter, which keys the Philox streams
(shuffle orders, in-game draws).  `state_dict(env)` copies both to host tensors (together with the current output tensors, so that
a restored batch shows the same observation before its next step); `load_state_dict(env, d)` puts them into an environment object
constructed with the same game, flags, batch size, seed and env_index_base, after which the run continues bit for bit.  The
result is a plain dict of tensors and ints: `torch.save` / `torch.load` round-trip it.  Rollout statistics are not part of a
checkpoint (read them with stats() before saving).
"""
import torch

from . import _abi

FORMAT = 1
# class name -> (persistent tensors, output tensors, (get, set) call-count symbols or None)
_LAYOUT = {
    "VectorEnv": (("state",), ("board", "cube", "value_board", "reward", "terminated", "step_type", "reason"), ("gw_call_count", "gw_set_call_count")),
    "ClassicVectorEnv": (("state",), ("board", "value_board", "reward", "terminated", "step_type", "reason", "actual"),
                         ("gw_call_count", "gw_set_call_count")),
    "SokobanVectorEnv": (("state",), ("board", "value_board", "reward", "terminated", "step_type", "reason", "actual"), None),
    "FiremakerVectorEnv": (("state",), ("board", "cube", "crop_workers", "crop_supervisor", "lcrop_workers", "lcrop_supervisor",
                                        "reward_workers", "reward_supervisor", "terminated", "step_type"),
                           ("gw_fm_call_count", "gw_fm_set_call_count")),
    "IslandMaVectorEnv": (("state", "maps"), ("board", "cube", "crop", "lcrop", "reward", "terminated", "step_type"),
                          ("gw_ima_call_count", "gw_ima_set_call_count")),
    "SavannaVectorEnv": (("state", "maps", "availability", "live_maps"),
                         ("_board_buf", "_cube_buf", "_crop_buf", "_lcrop_buf", "reward", "terminated", "step_type"),
                         ("gw_sav_call_count", "gw_sav_set_call_count")),
}


def _layout(env):
    for cls in type(env).__mro__:
        if cls.__name__ in _LAYOUT:
            return cls.__name__, _LAYOUT[cls.__name__]
    raise TypeError("%s is not a checkpointable environment batch" % type(env).__name__)


def state_dict(env):
    name, (persistent, outputs, calls) = _layout(env)
    torch.cuda.current_stream(env.device).synchronize()
    out = {"format": FORMAT, "class": name, "num_envs": int(env.num_envs), "env_index_base": int(getattr(env, "env_index_base", 0)),
           "call_count": int(getattr(_abi.load(), calls[0])(env._h)) if calls else 0, "tensors": {}, "outputs": {}}
    for key, names in (("tensors", persistent), ("outputs", outputs)):
        for nm in names:
            t = getattr(env, nm, None)
            if t is not None:
                out[key][nm] = t.detach().to("cpu", copy=True)
    return out


def load_state_dict(env, d):
    name, (persistent, outputs, calls) = _layout(env)
    if d.get("format") != FORMAT or d.get("class") != name:
        raise ValueError("checkpoint of %r (format %r) cannot be loaded into a %s" % (d.get("class"), d.get("format"), name))
    if int(d["num_envs"]) != int(env.num_envs):
        raise ValueError("checkpoint holds %d environments, this batch %d" % (d["num_envs"], env.num_envs))
    for key, names in (("tensors", persistent), ("outputs", outputs)):
        for nm in names:
            t = getattr(env, nm, None)
            saved = d[key].get(nm)
            if (t is None) != (saved is None):
                if key == "outputs":
                    continue                             # an optional output the other side did not request
                raise ValueError("checkpoint and batch disagree about the tensor %r" % nm)
            if t is not None:
                if tuple(t.shape) != tuple(saved.shape) or t.dtype != saved.dtype:
                    raise ValueError("tensor %r: checkpoint %s %s, batch %s %s" % (nm, tuple(saved.shape), saved.dtype, tuple(t.shape), t.dtype))
                t.copy_(saved)
    if calls:
        _abi.check(getattr(_abi.load(), calls[1])(env._h, int(d["call_count"])))
    torch.cuda.current_stream(env.device).synchronize()
