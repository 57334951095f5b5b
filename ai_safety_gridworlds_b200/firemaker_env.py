"""FiremakerVectorEnv: N lock-stepped firemaker_ex_ma games (3 agents each) on one B200.

The batched counterpart of `SafetyEnvironmentMoMa.step` / `EnvironmentMa.step`
(environments/shared/safety_game_moma.py:984, rl/pycolab_interface_ma.py:173) for BASELINE config 4:
one launch of the warp-per-environment kernel in csrc/gwsim_fm.cuh runs the three sequential
per-agent engine frames of a parallel step and renders the global and per-agent observations.
Agents are indexed 0 = '1', 1 = '2', 2 = 'S'.  No CPU fallback.
"""
import ctypes as C

import torch

from . import _abi
from .envs import make_spec
from .envs.firemaker_ex_ma import FiremakerSpec
from .parallel import MultiAgentStatsMixin
from .vector_env import _ptr


class FiremakerVectorEnv(MultiAgentStatsMixin):
    """Tensors (on `device`, reused between calls):
      board uint8 [N,17,17]; cube uint8 [N,9,17,17]; crop_workers uint8 [N,2,5,5]; crop_supervisor uint8 [N,33,33];
      lcrop_workers uint8 [N,2,9,5,5]; lcrop_supervisor uint8 [N,9,33,33];
      reward_workers float32 [N,2,2] (ENERGY, WORKSHOP); reward_supervisor float32 [N,3] (ENERGY, EXTERNAL_FIRE, TRESPASSING);
      terminated / step_type uint8 [N,3]
    """

    def __init__(self, num_envs, device=None, env_index_base=0, seed=0, autoreset_mode=_abi.GW_AUTORESET_SAME_STEP,
                 want_cube=True, want_crops=True, want_layer_crops=True, spec=None, **kwargs):
        self._h = None
        lib = _abi.load()
        if not torch.cuda.is_available():
            raise _abi.GwError("no CUDA device: the batched simulator has no CPU fallback")
        if spec is None:
            kwargs.setdefault("amount_agents", 3)              # BASELINE config 4; make_spec's own default is the reference's (2)
            spec = make_spec("firemaker_ex_ma", autoreset_mode=autoreset_mode, **kwargs)
        assert isinstance(spec, FiremakerSpec)
        spec.config.autoreset_mode = int(autoreset_mode)
        self.spec = spec
        self.num_envs = N = int(num_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = dev = torch.device("cuda", dev_index)
        self._lib = lib
        handle = C.c_void_p()
        _abi.check(lib.gw_fm_create(C.byref(spec.config), N, dev_index, int(env_index_base), int(seed), C.byref(handle)))
        self._h = handle
        u8 = dict(dtype=torch.uint8, device=dev)
        self.state = torch.zeros((N, _abi.GW_FM_STATE_WORDS, 4), dtype=torch.int32, device=dev)
        self.board = torch.zeros((N, 17, 17), **u8)
        self.cube = torch.zeros((N, 9, 17, 17), **u8) if want_cube else None
        self.crop_workers = torch.zeros((N, 2, 5, 5), **u8) if want_crops else None
        self.crop_supervisor = torch.zeros((N, 33, 33), **u8) if want_crops else None
        self.lcrop_workers = torch.zeros((N, 2, 9, 5, 5), **u8) if want_layer_crops else None
        self.lcrop_supervisor = torch.zeros((N, 9, 33, 33), **u8) if want_layer_crops else None
        self.reward_workers = torch.zeros((N, 2, 2), dtype=torch.float32, device=dev)
        self.reward_supervisor = torch.zeros((N, 3), dtype=torch.float32, device=dev)
        self.terminated = torch.zeros((N, 3), **u8)
        self.step_type = torch.zeros((N, 3), **u8)
        self._obs = _abi.GwFmObs(_ptr(self.board), _ptr(self.cube), _ptr(self.crop_workers), _ptr(self.crop_supervisor),
                                 _ptr(self.lcrop_workers), _ptr(self.lcrop_supervisor))
        self._out = _abi.GwFmOut(_ptr(self.reward_workers), _ptr(self.reward_supervisor), _ptr(self.terminated), _ptr(self.step_type))
        self._raw_dev = torch.zeros((_abi.GW_MA_STATS_LEN,), dtype=torch.float64, device=dev)
        self._stats_fns = (lib.gw_fm_stats_device, lib.gw_fm_stats_clear)
        from .envs.firemaker_ex_ma import REWARD_KEYS
        # raw layout: worker 1 [2], worker 2 [2], supervisor [3]; an absent worker '2' (amount_agents = 2) keeps its zero columns
        self._stats_columns = [(a if a in spec.reward_keys else None, list(REWARD_KEYS[a])) for a in ("1", "2", "S")]
        self.reset()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gw_fm_destroy(self._h)
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

    def reset(self, mask=None):
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _abi.check(self._lib.gw_fm_reset(self._h, _ptr(m), _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation()

    def step(self, actions, order=None, draws=None):
        """actions int32 [N,3] (MO numbering); order int32 [N,3] and draws float64 [N,K] replay a recorded
        reference run (tests); by default both come from the Philox streams."""
        N = self.num_envs
        if actions.dtype != torch.int32 or not actions.is_cuda or not actions.is_contiguous() or actions.shape != (N, 3):
            raise ValueError("actions must be a contiguous int32 CUDA tensor of shape [num_envs, 3]")
        if order is not None and (order.dtype != torch.int32 or order.shape != (N, 3) or not order.is_cuda):
            raise ValueError("order must be an int32 CUDA tensor of shape [num_envs, 3]")
        stride = 0
        if draws is not None:
            if draws.dtype != torch.float64 or draws.dim() != 2 or draws.shape[0] != N or not draws.is_cuda or not draws.is_contiguous():
                raise ValueError("draws must be a contiguous float64 CUDA tensor of shape [num_envs, K]")
            stride = draws.shape[1]
        _abi.check(self._lib.gw_fm_step(self._h, _ptr(actions), _ptr(order), _ptr(draws), stride, _ptr(self.state),
                                        C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation(), (self.reward_workers, self.reward_supervisor), self.terminated, self.step_type

    def step_raw(self, actions_ptr):
        return self._lib.gw_fm_step(self._h, actions_ptr, None, None, 0, _ptr(self.state), C.byref(self._obs), C.byref(self._out),
                                    self._stream())

    def observation(self):
        return dict(board=self.board, cube=self.cube, crop_workers=self.crop_workers, crop_supervisor=self.crop_supervisor,
                    lcrop_workers=self.lcrop_workers, lcrop_supervisor=self.lcrop_supervisor)

    def observe(self):
        N, dev = self.num_envs, self.device
        out = dict(metrics=torch.zeros((N, 16), dtype=torch.float64, device=dev), cumulative=torch.zeros((N, 7), dtype=torch.float32, device=dev),
                   frame=torch.zeros((N,), dtype=torch.int32, device=dev), pos=torch.zeros((N, 3, 2), dtype=torch.int16, device=dev),
                   ext_fires=torch.zeros((N,), dtype=torch.int32, device=dev), directions=torch.zeros((N, 3, 2), dtype=torch.int8, device=dev))
        ex = _abi.GwFmExtras(_ptr(out["metrics"]), _ptr(out["cumulative"]), _ptr(out["frame"]), _ptr(out["pos"]), _ptr(out["ext_fires"]),
                             _ptr(out["directions"]))
        _abi.check(self._lib.gw_fm_observe(self._h, _ptr(self.state), C.byref(ex), self._stream()))
        return out

    @property
    def launch_count(self):
        return int(self._lib.gw_fm_launch_count(self._h))

    def bytes_per_env_step(self):
        """3 actions + state in/out (160 B each) + every emitted tensor of one parallel step"""
        b = 12 + 2 * 160 + 289 + 28 + 6
        if self.cube is not None:
            b += 9 * 289
        if self.crop_workers is not None:
            b += 50 + 1089
        if self.lcrop_workers is not None:
            b += 9 * (50 + 1089)
        return b
