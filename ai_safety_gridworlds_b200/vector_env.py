"""VectorEnv: N lock-stepped instances of one environment type on one B200.

The host-side counterpart of the reference's per-instance stack
`SafetyEnvironmentMo.step -> EnvironmentMo.step -> Engine.play -> _process_timestep`
(shared/safety_game_mo.py:810, shared/rl/pycolab_interface_mo.py:157, pycolab/engine.py:583,
shared/safety_game_mo.py:971), batched: every quantity the reference returns per step is a torch
CUDA tensor with the environment index outermost, produced by ONE launch of the fused kernel in
csrc/gwsim.cu through the C ABI of include/gwsim.h.  torch is used for device memory, streams
and torch.distributed only.  There is no CPU fallback: without the built library or without a
CUDA device construction raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from .envs import make_spec
from .envs.common import EnvSpec


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class VectorEnv(object):
    """`num_envs` environments of one EnvSpec, stepped in lockstep on `device`.

    Tensors (all on `device`, reused between calls -- clone what you keep):
      board        uint8  [N, H, W]     rendered board as ASCII codes  (info['ascii_codes'])
      cube         uint8  [N, L, H, W]  0/1 layers cube, channel order spec.layer_order
      value_board  float32[N, H, W]     value-mapped board (the Gym observation)
      reward       float32[N, R]        reward vector, dimension order spec.reward_keys
      terminated   uint8  [N]; step_type uint8 [N]; reason int8 [N]
    """

    def __init__(self, env, num_envs, device=None, env_index_base=0, autoreset_mode=_abi.GW_AUTORESET_SAME_STEP,
                 want_board=True, want_cube=True, want_value_board=True, **kwargs):
        self._h = None
        lib = _abi.load()
        if not torch.cuda.is_available():
            raise _abi.GwError("no CUDA device: the batched simulator has no CPU fallback")
        if isinstance(env, EnvSpec):
            spec = env.with_autoreset(autoreset_mode)
        else:
            spec = make_spec(env, autoreset_mode=autoreset_mode, **kwargs)
        self.spec = spec
        self.num_envs = int(num_envs)
        if self.num_envs <= 0:
            raise ValueError("num_envs must be positive")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _abi.GwError("VectorEnv needs a CUDA device, got %s" % (self.device,))
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self.env_index_base = int(env_index_base)
        self._lib = lib

        handle = C.c_void_p()
        _abi.check(lib.gw_create(C.byref(spec.config), self.num_envs, dev_index, self.env_index_base, C.byref(handle)))
        self._h = handle
        N, Hh, Ww, L, R = self.num_envs, spec.height, spec.width, spec.n_layers, spec.n_rewards
        self.state_words = int(lib.gw_state_words(C.byref(spec.config)))
        nbytes = int(lib.gw_state_bytes(C.byref(spec.config), N))
        chunks = (N + 31) // 32
        assert nbytes == self.state_words * _abi.GW_STATE_WORD_BYTES * 32 * chunks
        dev = self.device
        # opaque state blob: [chunks of 32 environments][state_words][32] 16-byte words
        self.state = torch.zeros((chunks, self.state_words, 32, 4), dtype=torch.int32, device=dev)
        self.board = torch.empty((N, Hh, Ww), dtype=torch.uint8, device=dev) if want_board else None
        self.cube = torch.empty((N, L, Hh, Ww), dtype=torch.uint8, device=dev) if want_cube else None
        self.value_board = torch.empty((N, Hh, Ww), dtype=torch.float32, device=dev) if want_value_board else None
        self.reward = torch.zeros((N, R), dtype=torch.float32, device=dev)
        self.terminated = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self.step_type = torch.zeros((N,), dtype=torch.uint8, device=dev)
        self.reason = torch.full((N,), -1, dtype=torch.int8, device=dev)
        self._obs = _abi.GwObs(_ptr(self.board), _ptr(self.cube), _ptr(self.value_board))
        self._out = _abi.GwStepOut(_ptr(self.reward), _ptr(self.terminated), _ptr(self.step_type), _ptr(self.reason), None)
        self._raw_dev = torch.zeros((_abi.GW_STATS_RAW_LEN,), dtype=torch.float64, device=dev)
        self._host = None
        self.reset()

    # ------------------------------------------------------------------ lifetime
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

    # ------------------------------------------------------------------ stepping
    def reset(self, mask=None):
        """New episode in every environment (or where `mask` is non-zero); renders the observation.
        safety_game_mo.py:526-724 + pycolab/engine.py:520-581."""
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            if m.shape != (self.num_envs,):
                raise ValueError("reset mask must have shape [num_envs]")
        _abi.check(self._lib.gw_reset(self._h, _ptr(m), _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation()

    def step(self, actions):
        """actions: int32 CUDA tensor [N] in the MO action numbering (safety_game_mo_base.py:76-93).
        Returns (observation dict, reward [N,R], terminated [N], step_type [N], reason [N])."""
        if actions.dtype != torch.int32 or not actions.is_cuda or not actions.is_contiguous() or actions.shape != (self.num_envs,):
            raise ValueError("actions must be a contiguous int32 CUDA tensor of shape [num_envs]")
        _abi.check(self._lib.gw_step(self._h, _ptr(actions), _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation(), self.reward, self.terminated, self.step_type, self.reason

    def step_raw(self, actions_ptr):
        """The bare C-ABI call on a raw device pointer (bench inner loop)."""
        return self._lib.gw_step(self._h, actions_ptr, _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream())

    def observation(self):
        return {"board": self.board, "cube": self.cube, "value_board": self.value_board}

    def random_actions(self, seed, step, out=None, lo=None, hi=None):
        """Philox4x32-10 keyed by (seed, global env index, step): U{lo..hi} in the MO numbering."""
        if out is None:
            out = torch.empty((self.num_envs,), dtype=torch.int32, device=self.device)
        lo = 0 if lo is None else lo
        hi = 4 if hi is None else hi
        _abi.check(self._lib.gw_random_actions(self._h, int(seed), int(step), int(lo), int(hi), _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ host-buffer (end-to-end) path
    def _host_buffers(self):
        if self._host is None:
            N, s = self.num_envs, self.spec
            pin = dict(pin_memory=True)
            self._host = dict(
                actions=torch.zeros((N,), dtype=torch.int32, **pin),
                value_board=torch.zeros((N, s.height, s.width), dtype=torch.float32, **pin) if self.value_board is not None else None,
                board=torch.zeros((N, s.height, s.width), dtype=torch.uint8, **pin) if self.board is not None else None,
                reward=torch.zeros((N, s.n_rewards), dtype=torch.float32, **pin),
                terminated=torch.zeros((N,), dtype=torch.uint8, **pin),
            )
            self._dev_actions = torch.zeros((N,), dtype=torch.int32, device=self.device)
        return self._host

    def step_host(self, actions_host, observation="value_board"):
        """One step with HOST buffers: actions (pinned int32 [N]) are copied to the device, the fused
        kernel runs, and the Gym result tuple -- observation, reward, terminated -- is copied back
        into pinned host tensors.  Returns (obs_host, reward_host, terminated_host) after the
        stream has been synchronised.  This is the call bench.py times as `e2e`."""
        hb = self._host_buffers()
        self._dev_actions.copy_(actions_host, non_blocking=True)
        _abi.check(self.step_raw(_ptr(self._dev_actions)))
        src = self.value_board if observation == "value_board" else self.board
        dst = hb[observation]
        dst.copy_(src, non_blocking=True)
        hb["reward"].copy_(self.reward, non_blocking=True)
        hb["terminated"].copy_(self.terminated, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return dst, hb["reward"], hb["terminated"]

    def host_bytes_per_step(self, observation="value_board"):
        s = self.spec
        obs_b = s.cells * (4 if observation == "value_board" else 1)
        return self.num_envs * 4, self.num_envs * (obs_b + 4 * s.n_rewards + 1)

    # ------------------------------------------------------------------ extras / statistics
    def observe(self, f64=False):
        """metrics_dict values, cumulative / average reward, frame, agent position, safety and the
        Gini / variance scalars of _process_timestep (columns: gini_index, cumulative_gini_index,
        mo_variance, cumulative_mo_variance, average_mo_variance) -- computed from the state.
        `f64=True` adds `cumulative_f64`, the episode return before it is rounded to float32 (the CSV log needs it)."""
        N, M, R = self.num_envs, len(self.spec.metric_names), self.spec.n_rewards
        dev = self.device
        out = dict(metrics=torch.zeros((N, max(M, 1)), dtype=torch.float64, device=dev)[:, :M].contiguous() if M else None,
                   cumulative=torch.zeros((N, R), dtype=torch.float32, device=dev),
                   frame=torch.zeros((N,), dtype=torch.int32, device=dev),
                   pos=torch.zeros((N, 2), dtype=torch.int16, device=dev),
                   safety=torch.zeros((N,), dtype=torch.int16, device=dev),
                   average=torch.zeros((N, R), dtype=torch.float32, device=dev),
                   scalars=torch.zeros((N, 5), dtype=torch.float64, device=dev))
        if f64:
            out["cumulative_f64"] = torch.zeros((N, R), dtype=torch.float64, device=dev)
        ex = _abi.GwExtras(_ptr(out["metrics"]), _ptr(out["cumulative"]), _ptr(out["frame"]), _ptr(out["pos"]), _ptr(out["safety"]),
                           _ptr(out["average"]), _ptr(out["scalars"]), _ptr(self.reward), None, None, _ptr(out.get("cumulative_f64")))
        _abi.check(self._lib.gw_observe(self._h, _ptr(self.state), C.byref(ex), self._stream()))
        return out

    def peek_fractions(self):
        d = torch.zeros((self.num_envs,), dtype=torch.float64, device=self.device)
        f = torch.zeros((self.num_envs,), dtype=torch.float64, device=self.device)
        _abi.check(self._lib.gw_peek_fractions(self._h, _ptr(self.state), _ptr(d), _ptr(f), self._stream()))
        return d, f

    def stats_raw_device(self):
        """Raw statistics vector (float64 [GW_STATS_RAW_LEN], integer-valued except 4 slots) on the
        device, asynchronous: the tensor a sharded job all-reduces."""
        _abi.check(self._lib.gw_stats_device(self._h, _ptr(self._raw_dev), self._stream()))
        return self._raw_dev

    def finalize_stats(self, raw_host):
        return finalize_stats(self.spec, raw_host)

    def stats(self, group=None):
        """Episode statistics since construction / clear_stats.  With a torch.distributed `group`
        (or the default group when initialised and group is True) the raw vector is all-reduced
        (SUM, NCCL over NVLink) first, so every rank returns the whole-job statistics."""
        raw = self.stats_raw_device()
        if group is not None:
            import torch.distributed as dist
            raw = raw.clone()
            dist.all_reduce(raw, op=dist.ReduceOp.SUM, group=None if group is True else group)
        return self.finalize_stats(raw.cpu().numpy())

    def clear_stats(self):
        _abi.check(self._lib.gw_stats_clear(self._h, self._stream()))

    @property
    def launch_count(self):
        return int(self._lib.gw_launch_count(self._h))

    def bytes_per_env_step(self):
        """Algorithmic HBM bytes one environment step moves: action + state in +
        state out + every emitted output tensor."""
        s = self.spec
        b = 4 + 2 * self.state_words * _abi.GW_STATE_WORD_BYTES
        if self.board is not None:
            b += s.cells
        if self.cube is not None:
            b += s.n_layers * s.cells
        if self.value_board is not None:
            b += 4 * s.cells
        b += 4 * s.n_rewards + 3          # reward row + terminated + step_type + reason
        return b


def finalize_stats(spec, raw_host):
    """Raw statistics vector (e.g. the all-reduced one) -> dict; pure host arithmetic."""
    raw = np.ascontiguousarray(raw_host, np.float64)
    assert raw.shape == (_abi.GW_STATS_RAW_LEN,)
    out = np.zeros(_abi.GW_STATS_LEN, np.float64)
    _abi.check(_abi.load().gw_stats_finalize(C.byref(spec.config), raw.ctypes.data_as(C.POINTER(C.c_double)),
                                             out.ctypes.data_as(C.POINTER(C.c_double))))
    return stats_dict(out, spec)


def stats_dict(vec, spec):
    R = spec.n_rewards
    episodes = vec[_abi.GW_STAT_EPISODES]
    ret = vec[_abi.GW_STAT_RETURN_SUM:_abi.GW_STAT_RETURN_SUM + R]
    return dict(
        env_steps=int(vec[_abi.GW_STAT_ENV_STEPS]), episodes=int(episodes), length_sum=int(vec[_abi.GW_STAT_LENGTH_SUM]),
        reasons=dict(terminated=int(vec[_abi.GW_STAT_REASON0]), max_steps=int(vec[_abi.GW_STAT_REASON0 + 1]),
                     interrupted=int(vec[_abi.GW_STAT_REASON0 + 2]), quit=int(vec[_abi.GW_STAT_REASON0 + 3])),
        return_sum=dict(zip(spec.reward_keys, (float(x) for x in ret))),
        mean_return=dict(zip(spec.reward_keys, (float(x / episodes) if episodes else float("nan") for x in ret))),
        mean_length=float(vec[_abi.GW_STAT_LENGTH_SUM] / episodes) if episodes else float("nan"),
    )
