"""SokobanVectorEnv: side_effects_sokoban on its big maps (levels 1-3; level 0 too) on one B200.

The batched counterpart of `SideEffectsSokobanEnvironment.step` (environments/side_effects_sokoban.py:320-370 over
shared/safety_game.py:82-300): N environments of one level, one launch of `gw_sok_kernel` (csrc/gwsim_sok.cuh) per step behind
the C ABI of include/gwsim_sok.h.  Actions use the original numbering (NOOP 0, UP 1, DOWN 2, LEFT 3, RIGHT 4, QUIT 9).  Exposes
the tensors and methods of ClassicVectorEnv that the Gym wrapper reads.  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from .envs import make_spec
from .envs.classic import SokSpec
from .vector_env import _ptr

ROW = _abi.GW_SOK_MAX_CELLS       # one dense 128-entry board row per environment (pitch = width)


def crop_rows(rows, spec):
    """[..., 128] board rows -> their [..., H, W] boards (a view)."""
    H, W = spec.height, spec.width
    return rows[..., :H * W].reshape(rows.shape[:-1] + (H, W))


class SokobanVectorEnv(object):
    """Tensors (on `device`, reused between calls):
      board uint8 [N, 128], value_board float32 [N, 128] (zero past H*W; `boards()` gives [N, H, W]),
      reward float32 [N, 2] (reward, hidden-reward delta), terminated / step_type uint8 [N], reason / actual int8 [N]."""

    def __init__(self, spec, num_envs, device=None, autoreset_mode=_abi.GW_AUTORESET_SAME_STEP, want_board=True, want_value_board=True,
                 **kwargs):
        self._h = None
        lib = _abi.load()
        if not torch.cuda.is_available():
            raise _abi.GwError("no CUDA device: the batched simulator has no CPU fallback")
        if isinstance(spec, str):
            spec = make_spec(spec, autoreset_mode=autoreset_mode, **kwargs)
        if not isinstance(spec, SokSpec):
            raise ValueError("SokobanVectorEnv serves side_effects_sokoban levels 1-3 (and 0 through compile_sokoban_big)")
        self.spec = spec.with_autoreset(autoreset_mode)
        self.specs = [self.spec]
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self._lib = lib
        handle = C.c_void_p()
        _abi.check(lib.gw_sok_create(C.byref(self.spec.config), self.num_envs, dev_index, C.byref(handle)))
        self._h = handle
        N, dev = self.num_envs, self.device
        self.state = torch.zeros((int(lib.gw_sok_state_bytes(N)) // 4,), dtype=torch.int32, device=dev)
        self.board = torch.zeros((N, ROW), dtype=torch.uint8, device=dev) if want_board else None
        self.value_board = torch.zeros((N, ROW), dtype=torch.float32, device=dev) if want_value_board else None
        self.reward = torch.zeros((N, 2), dtype=torch.float32, device=dev)
        self.terminated = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self.step_type = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self.reason = torch.full((N,), -1, dtype=torch.int8, device=dev)
        self.actual = torch.full((N,), -1, dtype=torch.int8, device=dev)
        self._obs = _abi.GwSokObs(_ptr(self.board), _ptr(self.value_board))
        self._out = _abi.GwSokOut(_ptr(self.reward), _ptr(self.terminated), _ptr(self.step_type), _ptr(self.reason), _ptr(self.actual))
        self._raw_dev = torch.zeros((_abi.GW_SOK_STATS_LEN,), dtype=torch.float64, device=dev)
        self.reset()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gw_sok_destroy(self._h)
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

    def boards(self, which="board"):
        return crop_rows(getattr(self, which), self.spec)

    def reset(self, mask=None):
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _abi.check(self._lib.gw_sok_reset(self._h, _ptr(m), _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation()

    def step(self, actions):
        if actions.dtype != torch.int32 or not actions.is_cuda or not actions.is_contiguous() or actions.shape != (self.num_envs,):
            raise ValueError("actions must be a contiguous int32 CUDA tensor of shape [num_envs]")
        _abi.check(self._lib.gw_sok_step(self._h, _ptr(actions), _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation(), self.reward, self.terminated, self.step_type, self.reason

    def step_raw(self, actions_ptr):
        return self._lib.gw_sok_step(self._h, actions_ptr, _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream())

    def observation(self):
        return {"board": self.board, "value_board": self.value_board}

    def random_actions(self, seed, step, out=None, lo=1, hi=4):
        g = torch.Generator(device=self.device)
        g.manual_seed((int(seed) << 32) ^ int(step))
        a = torch.randint(lo, hi + 1, (self.num_envs,), generator=g, device=self.device, dtype=torch.int32)
        if out is not None:
            out.copy_(a)
            return out
        return a

    # host-buffer (end-to-end) path: actions H2D, kernel, observation + reward + terminated D2H
    def step_host(self, actions_host, observation="value_board"):
        if getattr(self, "_host", None) is None:
            N, pin = self.num_envs, dict(pin_memory=True)
            self._host = dict(value_board=torch.zeros((N, ROW), dtype=torch.float32, **pin), board=torch.zeros((N, ROW), dtype=torch.uint8, **pin),
                              reward=torch.zeros((N, 2), dtype=torch.float32, **pin), terminated=torch.zeros((N,), dtype=torch.uint8, **pin))
            self._dev_actions = torch.zeros((N,), dtype=torch.int32, device=self.device)
        hb = self._host
        self._dev_actions.copy_(actions_host, non_blocking=True)
        _abi.check(self.step_raw(_ptr(self._dev_actions)))
        hb[observation].copy_(self.value_board if observation == "value_board" else self.board, non_blocking=True)
        hb["reward"].copy_(self.reward, non_blocking=True)
        hb["terminated"].copy_(self.terminated, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return hb[observation], hb["reward"], hb["terminated"]

    def host_bytes_per_step(self, observation="value_board"):
        return self.num_envs * 4, self.num_envs * ((4 * ROW if observation == "value_board" else ROW) + 8 + 1)

    state_words = 1

    def stats_raw_device(self):
        _abi.check(self._lib.gw_sok_stats_device(self._h, _ptr(self._raw_dev), self._stream()))
        return self._raw_dev

    def finalize_stats(self, raw_host):
        v, S = np.asarray(raw_host, np.float64), _abi.SOK_STAT
        ep = v[S["EPISODES"]]
        return dict(env_steps=int(v[S["ENV_STEPS"]]), episodes=int(ep), length_sum=int(v[S["LENGTH_SUM"]]),
                    mean_length=float(v[S["LENGTH_SUM"]] / ep) if ep else float("nan"),
                    return_sum=float(v[S["RETURN_SUM"]]), hidden_sum=float(v[S["HIDDEN_SUM"]]), performance_sum=float(v[S["HIDDEN_SUM"]]))

    def set_coin_override(self, coins):
        raise ValueError("side_effects_sokoban makes no per-episode random draw")

    def observe(self, layers=False):
        N, dev = self.num_envs, self.device
        raw = torch.zeros((N, 2), dtype=torch.int32, device=dev)
        out = dict(frame=torch.zeros((N,), dtype=torch.int32, device=dev), pos=torch.zeros((N, 2), dtype=torch.int16, device=dev),
                   boxes=torch.zeros((N, _abi.GW_SOK_MAX_BOXES), dtype=torch.uint8, device=dev),
                   coins=torch.zeros((N,), dtype=torch.uint8, device=dev), coin=torch.zeros((N,), dtype=torch.int8, device=dev))
        ex = _abi.GwSokExtras(_ptr(raw), _ptr(out["frame"]), _ptr(out["pos"]), _ptr(out["boxes"]), _ptr(out["coins"]))
        _abi.check(self._lib.gw_sok_observe(self._h, _ptr(self.state), C.byref(ex), self._stream()))
        out["cumulative"] = raw.float()            # integers: exact
        return out

    def stats(self, group=None):
        """Episode statistics since the last clear; `group` all-reduces the (integer-valued) raw vector first."""
        _abi.check(self._lib.gw_sok_stats_device(self._h, _ptr(self._raw_dev), self._stream()))
        raw = self._raw_dev
        if group is not None:
            import torch.distributed as dist
            raw = raw.clone()
            dist.all_reduce(raw, op=dist.ReduceOp.SUM, group=None if group is True else group)
        v = raw.cpu().numpy()
        S = _abi.SOK_STAT
        ep = v[S["EPISODES"]]
        return dict(env_steps=int(v[S["ENV_STEPS"]]), episodes=int(ep), length_sum=int(v[S["LENGTH_SUM"]]),
                    return_sum=float(v[S["RETURN_SUM"]]), hidden_sum=float(v[S["HIDDEN_SUM"]]), performance_sum=float(v[S["HIDDEN_SUM"]]),
                    reasons=dict(terminated=int(v[S["REASON0"]]), max_steps=int(v[S["REASON0"] + 1]), interrupted=int(v[S["REASON0"] + 2]),
                                 quit=int(v[S["REASON0"] + 3])),
                    overall_performance=float(v[S["HIDDEN_SUM"]] / ep) if ep else float("nan"))

    def clear_stats(self):
        _abi.check(self._lib.gw_sok_stats_clear(self._h, self._stream()))

    @property
    def launch_count(self):
        return int(self._lib.gw_sok_launch_count(self._h))

    def bytes_per_env_step(self):
        """action + state in/out (16 B) + board row (+ value row) + reward row + 4 flag bytes"""
        return 4 + 2 * 16 + 8 + 4 + (ROW if self.board is not None else 0) + (4 * ROW if self.value_board is not None else 0)
