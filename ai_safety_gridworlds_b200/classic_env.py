"""ClassicVectorEnv: a MIXED batch of original-DeepMind-suite environments on one B200.

The batched counterpart of `safety_game.SafetyEnvironment.step` (environments/shared/safety_game.py:314)
for safe_interruptibility, side_effects_sokoban (level 0), absent_supervisor, conveyor_belt,
whisky_gold, boat_race, island_navigation, distributional_shift, rocks_diamonds, tomato_watering and
tomato_crmdp, friend_foe: environments [sum(counts[:t]), sum(counts[:t+1])) are of type specs[t], all stepped by
ONE launch of the fused kernel in csrc/gwsim_classic.cuh (BASELINE config 5).  Actions use the
original numbering (NOOP 0, UP 1, DOWN 2, LEFT 3, RIGHT 4, QUIT 9).  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from .envs import make_spec
from .envs.common import EnvSpec
from .vector_env import _ptr

SIDE = 8          # GW_CLASSIC_SIDE: one 64-entry board row per environment, [8, 8] with a pitch of 8 (maps wider than 8: dense)


def crop_board(rows, spec):
    """[..., 8, 8] board rows of environments of type `spec` -> their [..., H, W] boards (a view)."""
    H, W = spec.height, spec.width
    if W > SIDE:                                   # the 7 x 9 maps are laid out densely in the 64 entries
        return rows.reshape(rows.shape[:-2] + (SIDE * SIDE,))[..., :H * W].reshape(rows.shape[:-2] + (H, W))
    return rows[..., :H, :W]


class ClassicVectorEnv(object):
    """Tensors (on `device`, reused between calls):
      board        uint8  [N, 8, 8]  rendered board, ASCII codes, zero outside each type's H x W (crop_board / boards_of
                                     give the H x W view; maps wider than 8 fill their 64 entries densely)
      value_board  float32[N, 8, 8]  value-mapped board (the Gym observation of the original suite)
      reward       float32[N, 2]     (reward, hidden-reward delta of this step)
      terminated / step_type uint8 [N]; reason / actual int8 [N]
    """

    def __init__(self, envs, counts, device=None, env_index_base=0, seed=0, autoreset_mode=_abi.GW_AUTORESET_SAME_STEP,
                 want_board=True, want_value_board=True):
        self._h = None
        lib = _abi.load()
        if not torch.cuda.is_available():
            raise _abi.GwError("no CUDA device: the batched simulator has no CPU fallback")
        specs = []
        for e in envs:
            if isinstance(e, EnvSpec):
                specs.append(e.with_autoreset(autoreset_mode))
            elif isinstance(e, str):
                specs.append(make_spec(e, autoreset_mode=autoreset_mode))
            elif isinstance(e, (tuple, list)) and len(e) == 2:
                name, kwargs = e
                specs.append(make_spec(name, autoreset_mode=autoreset_mode, **kwargs))
            else:
                specs.append(e)
        for s in specs:
            if not isinstance(s, EnvSpec):
                raise ValueError("%s is not a mixed-batch type (side_effects_sokoban levels 1-3 run in SokobanVectorEnv)"
                                 % getattr(s, "name", type(s).__name__))
        self.specs, self.counts = specs, [int(c) for c in counts]
        if len(self.specs) != len(self.counts) or not self.specs:
            raise ValueError("one count per environment type")
        self.num_envs = sum(self.counts)
        self.type_start = np.concatenate([[0], np.cumsum(self.counts)]).astype(np.int64)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self._lib = lib
        cfgs = (_abi.GwConfig * len(specs))(*[s.config for s in specs])
        cnt = (C.c_int64 * len(specs))(*self.counts)
        handle = C.c_void_p()
        _abi.check(lib.gw_create_mixed(cfgs, len(specs), cnt, dev_index, int(env_index_base), int(seed), C.byref(handle)))
        self._h = handle
        N, dev = self.num_envs, self.device
        # opaque state: one 16-byte word per environment, plane-major [words, ceil32(N)]; only friend_foe (its three
        # PolicyEstimators) needs more than one plane
        self._state_words = max(int(lib.gw_state_words(C.byref(s.config))) for s in specs)
        self.state = torch.zeros((self._state_words, (N + 31) // 32 * 32, 4), dtype=torch.int32, device=dev)
        self.board = torch.zeros((N, SIDE, SIDE), dtype=torch.uint8, device=dev) if want_board else None
        self.value_board = torch.zeros((N, SIDE, SIDE), dtype=torch.float32, device=dev) if want_value_board else None
        self.reward = torch.zeros((N, 2), dtype=torch.float32, device=dev)
        self.terminated = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self.step_type = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self.reason = torch.full((N,), -1, dtype=torch.int8, device=dev)
        self.actual = torch.full((N,), -1, dtype=torch.int8, device=dev)
        self._obs = _abi.GwObs(_ptr(self.board), None, _ptr(self.value_board))
        self._out = _abi.GwStepOut(_ptr(self.reward), _ptr(self.terminated), _ptr(self.step_type), _ptr(self.reason), _ptr(self.actual))
        self._raw_dev = torch.zeros((_abi.GW_STATS_RAW_LEN,), dtype=torch.float64, device=dev)
        self._coins = None
        self.reset()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def state_dict(self):
        """Checkpoint of this batch as a dict of host tensors and ints (checkpoint.py); torch.save-able."""
        from . import checkpoint
        return checkpoint.state_dict(self)

    def load_state_dict(self, d):
        """Restores a checkpoint made by state_dict() of a batch built with the same game, flags, size, seed and index base."""
        from . import checkpoint
        checkpoint.load_state_dict(self, d)

    def set_coin_override(self, coins):
        """coins: uint8 CUDA tensor [N] (0/1 force the next episode's draw of that environment, 255 = draw) or None."""
        if coins is not None and (coins.dtype != torch.uint8 or not coins.is_cuda or coins.shape != (self.num_envs,)):
            raise ValueError("coins must be a uint8 CUDA tensor of shape [num_envs]")
        self._coins = coins
        _abi.check(self._lib.gw_set_coin_override(self._h, _ptr(coins)))

    def set_dried_override(self, dried):
        """dried: uint16 CUDA tensor [N], the tomato games' per-frame draws of the NEXT calls as a bit mask over each
        environment's tomato cells in row-major order (0xFFFF = draw from Philox), or None."""
        if dried is not None and (dried.dtype != torch.uint16 or not dried.is_cuda or dried.shape != (self.num_envs,)):
            raise ValueError("dried must be a uint16 CUDA tensor of shape [num_envs]")
        self._dried = dried
        _abi.check(self._lib.gw_set_dried_override(self._h, _ptr(dried)))

    def boards_of(self, t, which="board"):
        """The [counts[t], H, W] boards of type t (a view of `board` or `value_board`)."""
        lo, hi = int(self.type_start[t]), int(self.type_start[t + 1])
        return crop_board(getattr(self, which)[lo:hi], self.specs[t])

    def reset(self, mask=None):
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _abi.check(self._lib.gw_reset(self._h, _ptr(m), _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation()

    def step(self, actions):
        if actions.dtype != torch.int32 or not actions.is_cuda or not actions.is_contiguous() or actions.shape != (self.num_envs,):
            raise ValueError("actions must be a contiguous int32 CUDA tensor of shape [num_envs]")
        _abi.check(self.step_raw(_ptr(actions)))
        return self.observation(), self.reward, self.terminated, self.step_type, self.reason

    def step_raw(self, actions_ptr):
        return self._lib.gw_step(self._h, actions_ptr, _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream())

    def observation(self):
        return {"board": self.board, "value_board": self.value_board}

    # host-buffer (end-to-end) path: actions H2D, kernel, observation + reward + terminated D2H
    def step_host(self, actions_host, observation="value_board"):
        if getattr(self, "_host", None) is None:
            N, pin = self.num_envs, dict(pin_memory=True)
            self._host = dict(value_board=torch.zeros((N, SIDE, SIDE), dtype=torch.float32, **pin),
                              board=torch.zeros((N, SIDE, SIDE), dtype=torch.uint8, **pin),
                              reward=torch.zeros((N, 2), dtype=torch.float32, **pin),
                              terminated=torch.zeros((N,), dtype=torch.uint8, **pin))
            self._dev_actions = torch.zeros((N,), dtype=torch.int32, device=self.device)
        hb = self._host
        self._dev_actions.copy_(actions_host, non_blocking=True)
        _abi.check(self.step_raw(_ptr(self._dev_actions)))
        src = self.value_board if observation == "value_board" else self.board
        hb[observation].copy_(src, non_blocking=True)
        hb["reward"].copy_(self.reward, non_blocking=True)
        hb["terminated"].copy_(self.terminated, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return hb[observation], hb["reward"], hb["terminated"]

    def host_bytes_per_step(self, observation="value_board"):
        return self.num_envs * 4, self.num_envs * ((256 if observation == "value_board" else 64) + 8 + 1)

    @property
    def state_words(self):
        return self._state_words

    def policies(self):
        """friend_foe: float64 [N, 3, 2] view of the PolicyEstimator.policy vectors (friend, neutral, adversary); a vector that
        was never updated reads (0, 0) and stands for the initial (0.5, 0.5)."""
        if self._state_words < 4:
            raise ValueError("no friend_foe type in this batch")
        return self.state[1:4, :self.num_envs].contiguous().view(torch.float64).reshape(3, self.num_envs, 2).permute(1, 0, 2)

    def finalize_stats(self, raw_host):
        raw = np.ascontiguousarray(raw_host, np.float64)
        out = np.zeros(_abi.GW_STATS_LEN, np.float64)
        _abi.check(self._lib.gw_stats_finalize(C.byref(self.specs[0].config), raw.ctypes.data_as(C.POINTER(C.c_double)),
                                               out.ctypes.data_as(C.POINTER(C.c_double))))
        ep = out[_abi.GW_STAT_EPISODES]
        return dict(env_steps=int(out[_abi.GW_STAT_ENV_STEPS]), episodes=int(ep), length_sum=int(out[_abi.GW_STAT_LENGTH_SUM]),
                    mean_length=float(out[_abi.GW_STAT_LENGTH_SUM] / ep) if ep else float("nan"),
                    return_sum=float(out[_abi.GW_STAT_RETURN_SUM]), hidden_sum=float(out[_abi.GW_STAT_RETURN_SUM + 1]),
                    performance_sum=float(out[_abi.GW_STAT_PERFORMANCE_SUM]))

    def random_actions(self, seed, step, out=None, lo=1, hi=4):
        if out is None:
            out = torch.empty((self.num_envs,), dtype=torch.int32, device=self.device)
        _abi.check(self._lib.gw_random_actions(self._h, int(seed), int(step), int(lo), int(hi), _ptr(out), self._stream()))
        return out

    def observe(self, layers=False):
        """Per-environment quantities read from the state; `layers=True` adds the un-occluded layers of the MO re-wrappings
        (conveyor_belt_ex, safe_interruptibility_ex) as uint8 [N, GW_MAX_LAYERS, 8, 8] in the board-row layout."""
        N, dev = self.num_envs, self.device
        out = dict(cumulative=torch.zeros((N, 2), dtype=torch.float32, device=dev), frame=torch.zeros((N,), dtype=torch.int32, device=dev),
                   pos=torch.zeros((N, 2), dtype=torch.int16, device=dev), coin=torch.zeros((N,), dtype=torch.int8, device=dev))
        if layers:
            out["layers"] = torch.zeros((N, _abi.GW_MAX_LAYERS, SIDE, SIDE), dtype=torch.uint8, device=dev)
        ex = _abi.GwExtras(None, _ptr(out["cumulative"]), _ptr(out["frame"]), _ptr(out["pos"]), None, None, None, None, _ptr(out["coin"]),
                           _ptr(out.get("layers")))
        _abi.check(self._lib.gw_observe(self._h, _ptr(self.state), C.byref(ex), self._stream()))
        return out

    def stats_raw_device(self):
        _abi.check(self._lib.gw_stats_device(self._h, _ptr(self._raw_dev), self._stream()))
        return self._raw_dev

    def stats(self, group=None):
        """Whole-batch episode statistics (all types together); `group` all-reduces the raw vector first."""
        raw = self.stats_raw_device()
        if group is not None:
            import torch.distributed as dist
            raw = raw.clone()
            dist.all_reduce(raw, op=dist.ReduceOp.SUM, group=None if group is True else group)
        raw = raw.cpu().numpy()
        out = np.zeros(_abi.GW_STATS_LEN, np.float64)
        _abi.check(self._lib.gw_stats_finalize(C.byref(self.specs[0].config), raw.ctypes.data_as(C.POINTER(C.c_double)),
                                               out.ctypes.data_as(C.POINTER(C.c_double))))
        ep = out[_abi.GW_STAT_EPISODES]
        return dict(env_steps=int(out[_abi.GW_STAT_ENV_STEPS]), episodes=int(ep), length_sum=int(out[_abi.GW_STAT_LENGTH_SUM]),
                    reasons=dict(terminated=int(out[_abi.GW_STAT_REASON0]), max_steps=int(out[_abi.GW_STAT_REASON0 + 1]),
                                 interrupted=int(out[_abi.GW_STAT_REASON0 + 2]), quit=int(out[_abi.GW_STAT_REASON0 + 3])),
                    return_sum=float(out[_abi.GW_STAT_RETURN_SUM]), hidden_sum=float(out[_abi.GW_STAT_RETURN_SUM + 1]),
                    performance_sum=float(out[_abi.GW_STAT_PERFORMANCE_SUM]),
                    overall_performance=float(out[_abi.GW_STAT_PERFORMANCE_SUM] / ep) if ep else float("nan"))

    def clear_stats(self):
        _abi.check(self._lib.gw_stats_clear(self._h, self._stream()))

    @property
    def launch_count(self):
        return int(self._lib.gw_launch_count(self._h))

    def bytes_per_env_step(self):
        """action + state in/out (16 B) + padded board (+ value board) + reward row + 4 flag bytes"""
        b = 4 + 2 * 16 + 8 + 4
        if self.board is not None:
            b += 64
        if self.value_board is not None:
            b += 256
        return b
