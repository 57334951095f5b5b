"""IslandMaVectorEnv: N lock-stepped island_navigation_ex_ma games (2 agents each) on one B200.

The batched counterpart of `SafetyEnvironmentMoMa.step` / `EnvironmentMa.step`
(environments/shared/safety_game_moma.py:984, rl/pycolab_interface_ma.py:173) for
environments/island_navigation_ex_ma.py: one launch of the lane-per-environment kernel in
csrc/gwsim_ima.cuh runs the sequential per-agent engine frames of a parallel step and renders the
global observation and the two rotated agent views.  Agents are indexed 0 = '1', 1 = '2'.
No CPU fallback.
"""
import ctypes as C

import torch

from . import _abi
from .envs import make_spec
from .parallel import MultiAgentStatsMixin
from .vector_env import _ptr


class IslandMaVectorEnv(MultiAgentStatsMixin):
    """Tensors (on `device`, reused between calls):
      board uint8 [N,H,W]; cube uint8 [N,L,H,W]; crop uint8 [N,2,5,5]; lcrop uint8 [N,2,L,5,5];
      reward float32 [N,2,R] (sorted reward-dimension keys); terminated / step_type uint8 [N,2]
    """

    def __init__(self, num_envs, device=None, env_index_base=0, seed=0, autoreset_mode=_abi.GW_AUTORESET_SAME_STEP,
                 want_cube=True, want_crops=True, want_layer_crops=True, spec=None, **kwargs):
        self._h = None
        lib = _abi.load()
        if not torch.cuda.is_available():
            raise _abi.GwError("no CUDA device: the batched simulator has no CPU fallback")
        if spec is None:
            spec = make_spec("island_navigation_ex_ma", autoreset_mode=autoreset_mode, **kwargs)
        if spec.config.env_type != _abi.GW_ENV_ISLAND_NAVIGATION_EX_MA:
            raise ValueError("IslandMaVectorEnv needs an island_navigation_ex_ma spec")
        spec.config.autoreset_mode = int(autoreset_mode)
        self.spec = spec
        self.num_envs = N = int(num_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = dev = torch.device("cuda", dev_index)
        self._lib = lib
        handle = C.c_void_p()
        _abi.check(lib.gw_ima_create(C.byref(spec.config), N, dev_index, int(env_index_base), int(seed), C.byref(handle)))
        self._h = handle
        H, W, L, R = spec.height, spec.width, spec.n_layers, spec.n_rewards
        u8 = dict(dtype=torch.uint8, device=dev)
        self.state = torch.zeros((lib.gw_ima_state_bytes(C.byref(spec.config), N) // 4,), dtype=torch.int32, device=dev)
        self.board = torch.zeros((N, H, W), **u8)
        self.cube = torch.zeros((N, L, H, W), **u8) if want_cube else None
        self.crop = torch.zeros((N, 2, 5, 5), **u8) if want_crops else None
        self.lcrop = torch.zeros((N, 2, L, 5, 5), **u8) if want_layer_crops else None
        self.reward = torch.zeros((N, 2, R), dtype=torch.float32, device=dev)
        self.terminated = torch.zeros((N, 2), **u8)
        self.step_type = torch.zeros((N, 2), **u8)
        self._obs = _abi.GwImaObs(_ptr(self.board), _ptr(self.cube), _ptr(self.crop), _ptr(self.lcrop))
        self._out = _abi.GwImaOut(_ptr(self.reward), _ptr(self.terminated), _ptr(self.step_type))
        self._raw_dev = torch.zeros((_abi.GW_MA_STATS_LEN,), dtype=torch.float64, device=dev)
        self._stats_fns = (lib.gw_ima_stats_device, lib.gw_ima_stats_clear)
        self._stats_columns = [(a, list(spec.reward_keys)) for a in ("1", "2")]
        # map randomisation (island_navigation_ex_ma.py:67): every environment plays its own layout.  Frequency 3 = a fresh
        # layout for every game; 1 / 2 = a fresh layout at every explicit reset() (once per experiment / env-seed update).
        self.maps = None
        freq = int(spec.flags.get("map_randomization_frequency", 0))
        if freq:
            art = torch.tensor([ord(ch) for row in spec.art for ch in row], dtype=torch.uint8, device=dev)
            mode = _abi.GW_IMA_MAPS_SHUFFLE_EVERY_GAME if freq == 3 else _abi.GW_IMA_MAPS_SHUFFLE_ON_RESET
            self.set_maps(art.reshape(1, H, W).repeat(N, 1, 1).contiguous(), mode)
        self.reset()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gw_ima_destroy(self._h)
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

    def set_maps(self, maps, mode=_abi.GW_IMA_MAPS_STATIC):
        """maps: uint8 CUDA tensor [N, H, W], the ascii art of every environment's game (kept by reference: rewrite it between
        calls to replay given layouts; the library rewrites it in the shuffle modes), or None for the type's art everywhere."""
        if maps is not None:
            s = self.spec
            if maps.dtype != torch.uint8 or not maps.is_cuda or not maps.is_contiguous() or tuple(maps.shape) != (self.num_envs, s.height, s.width):
                raise ValueError("maps must be a contiguous uint8 CUDA tensor of shape [num_envs, H, W]")
        self.maps = maps
        _abi.check(self._lib.gw_ima_set_maps(self._h, _ptr(maps), int(mode)))

    def reset(self, mask=None):
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _abi.check(self._lib.gw_ima_reset(self._h, _ptr(m), _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation()

    def step(self, actions, order=None):
        """actions int32 [N,2] (MO numbering; entries of finished agents are ignored); order int32 [N,2] = execution order as
        agent indices, -1 = no frame (replays a recorded shuffle; {agent, -1} is the AEC single-agent frame); by default the
        live agents act in Philox-shuffled order."""
        N = self.num_envs
        if actions.dtype != torch.int32 or not actions.is_cuda or not actions.is_contiguous() or actions.shape != (N, 2):
            raise ValueError("actions must be a contiguous int32 CUDA tensor of shape [num_envs, 2]")
        if order is not None and (order.dtype != torch.int32 or order.shape != (N, 2) or not order.is_cuda or not order.is_contiguous()):
            raise ValueError("order must be a contiguous int32 CUDA tensor of shape [num_envs, 2]")
        _abi.check(self._lib.gw_ima_step(self._h, _ptr(actions), _ptr(order), _ptr(self.state), C.byref(self._obs), C.byref(self._out),
                                         self._stream()))
        return self.observation(), self.reward, self.terminated, self.step_type

    def step_raw(self, actions_ptr):
        return self._lib.gw_ima_step(self._h, actions_ptr, None, _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream())

    def observation(self):
        return dict(board=self.board, cube=self.cube, crop=self.crop, lcrop=self.lcrop)

    def observe(self):
        N, dev, R = self.num_envs, self.device, self.spec.n_rewards
        out = dict(metrics=torch.zeros((N, 16), dtype=torch.float64, device=dev), cumulative=torch.zeros((N, 2, R), dtype=torch.float32, device=dev),
                   frame=torch.zeros((N,), dtype=torch.int32, device=dev), pos=torch.zeros((N, 2, 2), dtype=torch.int16, device=dev),
                   directions=torch.zeros((N, 2, 2), dtype=torch.int8, device=dev))
        ex = _abi.GwImaExtras(_ptr(out["metrics"]), _ptr(out["cumulative"]), _ptr(out["frame"]), _ptr(out["pos"]), _ptr(out["directions"]))
        _abi.check(self._lib.gw_ima_observe(self._h, _ptr(self.state), C.byref(ex), self._stream()))
        slots = [self.spec.config.metric_slots[i] for i in range(self.spec.config.n_metrics)]
        out["metrics"] = out["metrics"][:, slots]                  # the columns this level activates, metrics_dict order
        return out

    @property
    def launch_count(self):
        return int(self._lib.gw_ima_launch_count(self._h))

    def bytes_per_env_step(self):
        """2 actions + state in/out (192 B each) + every emitted tensor of one parallel step"""
        s = self.spec
        b = 8 + 2 * 16 * _abi.GW_IMA_STATE_WORDS + s.cells + 2 * s.n_rewards * 4 + 4
        if self.maps is not None:
            b += s.cells                                           # the environment's own layout is read every step
        if self.cube is not None:
            b += s.n_layers * s.cells
        if self.crop is not None:
            b += 50
        if self.lcrop is not None:
            b += 2 * s.n_layers * 25
        return b
