"""SavannaVectorEnv: N lock-stepped aintelope_savanna games (one or two agents each) on one B200.

The batched counterpart of `SafetyEnvironmentMoMa.step` / `EnvironmentMa.step` (environments/shared/safety_game_moma.py:984,
rl/pycolab_interface_ma.py:173) for environments/aintelope/aintelope_savanna.py: one launch of the warp-per-environment kernel in
csrc/gwsim_sav.cuh runs the sequential per-agent engine frames of a parallel step and renders the global observation and the
agents' rotated views.  Agents are indexed 0 = '0', 1 = '1'; with amount_agents = 1 column 1 of the per-agent tensors is unused
(step type 3, zero views).  No CPU fallback.
"""
import ctypes as C

import torch

from . import _abi
from .envs import make_spec
from .envs.aintelope_savanna import SavSpec
from .parallel import MultiAgentStatsMixin
from .vector_env import _ptr


class SavannaVectorEnv(MultiAgentStatsMixin):
    """Tensors (on `device`, reused between calls), V = 2 * observation radius + 1:
      board uint8 [N,H,W]; cube uint8 [N,L,H,W]; crop uint8 [N,2,V,V]; lcrop uint8 [N,2,L,V,V] (views of row-padded buffers);
      reward float32 [N,2,R] (sorted reward-dimension keys); terminated / step_type uint8 [N,2];
      maps uint8 [N,H,W]: every environment's own layout (rewritten by the library when it draws the layouts)."""

    def __init__(self, num_envs, device=None, env_index_base=0, seed=0, autoreset_mode=_abi.GW_AUTORESET_SAME_STEP,
                 want_cube=True, want_crops=True, want_layer_crops=True, spec=None, **kwargs):
        self._h = None
        lib = _abi.load()
        if not torch.cuda.is_available():
            raise _abi.GwError("no CUDA device: the batched simulator has no CPU fallback")
        if spec is None:
            spec = make_spec("aintelope_savanna", autoreset_mode=autoreset_mode, **kwargs)
        if not isinstance(spec, SavSpec):
            raise ValueError("SavannaVectorEnv needs an aintelope_savanna spec")
        self.spec = spec = spec.with_autoreset(autoreset_mode)
        self.num_envs = N = int(num_envs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = dev = torch.device("cuda", dev_index)
        self._lib = lib
        handle = C.c_void_p()
        _abi.check(lib.gw_sav_create(C.byref(spec.config), N, dev_index, int(env_index_base), int(seed), C.byref(handle)))
        self._h = handle
        H, W, L, R, V = spec.height, spec.width, spec.n_layers, spec.n_rewards, spec.view
        u8 = dict(dtype=torch.uint8, device=dev)
        self.state = torch.zeros((int(lib.gw_sav_state_bytes(N)) // 4,), dtype=torch.int32, device=dev)
        # the kernel stores 16 bytes per lane: every row of the output tensors is padded to a multiple of 16 bytes (GW_SAV_PITCH);
        # the public tensors are views of the padded buffers with the reference's shapes
        cp, vp = (H * W + 15) // 16 * 16, (V * V + 15) // 16 * 16
        self._board_buf = torch.zeros((N, cp), **u8)
        self._cube_buf = torch.zeros((N, L, cp), **u8) if want_cube else None
        self._crop_buf = torch.zeros((N, 2, vp), **u8) if want_crops else None
        self._lcrop_buf = torch.zeros((N, 2, L, vp), **u8) if want_layer_crops else None
        self.board = self._board_buf[:, :H * W].unflatten(-1, (H, W))
        self.cube = self._cube_buf[..., :H * W].unflatten(-1, (H, W)) if want_cube else None
        self.crop = self._crop_buf[..., :V * V].unflatten(-1, (V, V)) if want_crops else None
        self.lcrop = self._lcrop_buf[..., :V * V].unflatten(-1, (V, V)) if want_layer_crops else None
        self.reward = torch.zeros((N, 2, R), dtype=torch.float32, device=dev)
        self.terminated = torch.zeros((N, 2), **u8)
        self.step_type = torch.zeros((N, 2), **u8)
        self._obs = _abi.GwSavObs(_ptr(self._board_buf), _ptr(self._cube_buf), _ptr(self._crop_buf), _ptr(self._lcrop_buf))
        self._out = _abi.GwSavOut(_ptr(self.reward), _ptr(self.terminated), _ptr(self.step_type))
        self._raw_dev = torch.zeros((_abi.GW_MA_STATS_LEN,), dtype=torch.float64, device=dev)
        self._stats_fns = (lib.gw_sav_stats_device, lib.gw_sav_stats_clear)
        self._stats_columns = [(a, list(spec.reward_keys)) for a in ("0", "1")[:spec.n_agents]]
        if spec.config.sustainability & _abi.GW_SAV_SUST_ON:
            # sustainability_challenge: the shared availabilities of 'D', 'd', 'F', 'f' and the running game's tiles are state
            self.availability = torch.zeros((N, 4), dtype=torch.float64, device=dev)
            self.live_maps = torch.zeros((N, H, W), **u8)
            _abi.check(lib.gw_sav_set_resources(self._h, _ptr(self.availability), _ptr(self.live_maps)))
        # every environment plays its own layout (map_randomization_frequency, aintelope_savanna.py:67): 3 = a fresh layout for
        # every game, 1 / 2 = a fresh layout at every explicit reset(), 0 = the level's map as it is
        art = torch.tensor([ord(ch) for row in spec.art for ch in row], dtype=torch.uint8, device=dev)
        freq = int(spec.flags.get("map_randomization_frequency", 0))
        mode = _abi.GW_IMA_MAPS_SHUFFLE_EVERY_GAME if freq == 3 else _abi.GW_IMA_MAPS_SHUFFLE_ON_RESET if freq else _abi.GW_IMA_MAPS_STATIC
        self.set_maps(art.reshape(1, H, W).repeat(N, 1, 1).contiguous(), mode)
        self.reset()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gw_sav_destroy(self._h)
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
        """maps: uint8 CUDA tensor [N, H, W], the ascii art of every environment's game, kept by reference: rewrite it between
        calls to replay given layouts (mode STATIC); the library rewrites it in the shuffle modes."""
        s = self.spec
        if maps.dtype != torch.uint8 or not maps.is_cuda or not maps.is_contiguous() or tuple(maps.shape) != (self.num_envs, s.height, s.width):
            raise ValueError("maps must be a contiguous uint8 CUDA tensor of shape [num_envs, H, W]")
        self.maps = maps
        _abi.check(self._lib.gw_sav_set_maps(self._h, _ptr(maps), int(mode)))

    def reset(self, mask=None):
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _abi.check(self._lib.gw_sav_reset(self._h, _ptr(m), _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation()

    def step(self, actions, order=None, draws=None):
        """actions int32 [N,2] (MO numbering; entries of finished or absent agents are ignored); order int32 [N,2] = execution
        order as agent indices, -1 = no frame; by default the live agents act in Philox-shuffled order.  draws float64 [N, K >= 32]
        replays the predators' draws of a recorded reference run and, with the sustainability challenge, the resource drapes' tile
        picks (by default both come from Philox)."""
        N = self.num_envs
        if actions.dtype != torch.int32 or not actions.is_cuda or not actions.is_contiguous() or actions.shape != (N, 2):
            raise ValueError("actions must be a contiguous int32 CUDA tensor of shape [num_envs, 2]")
        if order is not None and (order.dtype != torch.int32 or order.shape != (N, 2) or not order.is_cuda or not order.is_contiguous()):
            raise ValueError("order must be a contiguous int32 CUDA tensor of shape [num_envs, 2]")
        if draws is not None and (draws.dtype != torch.float64 or draws.dim() != 2 or draws.shape[0] != N or not draws.is_cuda or not draws.is_contiguous()):
            raise ValueError("draws must be a contiguous float64 CUDA tensor of shape [num_envs, K]")
        _abi.check(self._lib.gw_sav_step(self._h, _ptr(actions), _ptr(order), _ptr(draws), 0 if draws is None else int(draws.shape[1]),
                                         _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream()))
        return self.observation(), self.reward, self.terminated, self.step_type

    def step_raw(self, actions_ptr):
        return self._lib.gw_sav_step(self._h, actions_ptr, None, None, 0, _ptr(self.state), C.byref(self._obs), C.byref(self._out), self._stream())

    def observation(self):
        return dict(board=self.board, cube=self.cube, crop=self.crop, lcrop=self.lcrop)

    def observe(self, all_slots=False):
        N, dev, R = self.num_envs, self.device, self.spec.n_rewards
        out = dict(metrics=torch.zeros((N, _abi.GW_SAV_METRICS), dtype=torch.float64, device=dev),
                   cumulative=torch.zeros((N, 2, R), dtype=torch.float32, device=dev), frame=torch.zeros((N,), dtype=torch.int32, device=dev),
                   pos=torch.zeros((N, 2, 2), dtype=torch.int16, device=dev), directions=torch.zeros((N, 2, 2), dtype=torch.int8, device=dev))
        ex = _abi.GwSavExtras(_ptr(out["metrics"]), _ptr(out["cumulative"]), _ptr(out["frame"]), _ptr(out["pos"]), _ptr(out["directions"]))
        _abi.check(self._lib.gw_sav_observe(self._h, _ptr(self.state), C.byref(ex), self._stream()))
        if not all_slots:
            out["metrics"] = out["metrics"][:, self.spec.metric_slots]      # the labels this map activates, metrics_labels order
        return out

    @property
    def launch_count(self):
        return int(self._lib.gw_sav_launch_count(self._h))

    def bytes_per_env_step(self):
        """2 actions + state in/out (192 B each) + the environment's map + every emitted tensor of one parallel step"""
        s = self.spec
        cp, vp = (s.cells + 15) // 16 * 16, (s.view * s.view + 15) // 16 * 16
        b = 8 + 2 * _abi.GW_SAV_STATE_BYTES + s.cells + cp + 2 * s.n_rewards * 4 + 4
        if self.cube is not None:
            b += s.n_layers * cp
        if self.crop is not None:
            b += s.n_agents * vp                              # the columns of an absent agent are never written
        if self.lcrop is not None:
            b += s.n_agents * s.n_layers * vp
        return b
