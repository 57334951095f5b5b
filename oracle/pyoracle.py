"""ctypes wrapper of oracle/libgw_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  It consumes the same GwConfig POD the CUDA library does, so a parity test feeds one
spec to both sides.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgw_oracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libgw_oracle.so"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.or_create.restype = C.c_void_p
        _lib.or_create.argtypes = [C.c_void_p, C.c_int64]
        _lib.or_destroy.argtypes = [C.c_void_p]
        _lib.or_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        _lib.or_step.argtypes = [C.c_void_p] + [C.c_void_p] * 8 + [C.c_int]
        _lib.or_observe.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        _lib.or_peek_fractions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.or_random_actions.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]
        _lib.or_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle(object):
    """N independent environments of one EnvSpec, stepped by the scalar C restatement."""

    def __init__(self, spec, n_envs, want_value_board=True, want_cube=True):
        self.spec = spec
        self.n = int(n_envs)
        self._h = lib().or_create(C.byref(spec.config), self.n)
        if not self._h:
            raise ValueError("oracle rejected the configuration")
        cells, L, R = spec.cells, spec.n_layers, spec.n_rewards
        self.board = np.zeros((self.n, spec.height, spec.width), np.uint8)
        self.cube = np.zeros((self.n, L, spec.height, spec.width), np.uint8) if want_cube else None
        self.value_board = np.zeros((self.n, spec.height, spec.width), np.float32) if want_value_board else None
        self.reward = np.zeros((self.n, R), np.float32)
        self.terminated = np.zeros(self.n, np.uint8)
        self.step_type = np.zeros(self.n, np.uint8)
        self.reason = np.full(self.n, -1, np.int8)

    def close(self):
        if self._h:
            lib().or_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _outs(self):
        return [_p(self.board), _p(self.cube), _p(self.value_board), _p(self.reward), _p(self.terminated),
                _p(self.step_type), _p(self.reason)]

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().or_reset(self._h, _p(m), *self._outs())

    def step(self, actions, n_threads=1):
        a = np.ascontiguousarray(actions, np.int32)
        assert a.shape == (self.n,)
        lib().or_step(self._h, _p(a), *self._outs(), int(n_threads))

    def observe(self):
        M, R = len(self.spec.metric_names), self.spec.n_rewards
        out = dict(metrics=np.zeros((self.n, M), np.float64), cumulative=np.zeros((self.n, R), np.float32),
                   frame=np.zeros(self.n, np.int32), pos=np.zeros((self.n, 2), np.int16), safety=np.zeros(self.n, np.int16))
        lib().or_observe(self._h, _p(out["metrics"]), _p(out["cumulative"]), _p(out["frame"]), _p(out["pos"]), _p(out["safety"]))
        return out

    def fractions(self):
        d = np.zeros(self.n, np.float64)
        f = np.zeros(self.n, np.float64)
        lib().or_peek_fractions(self._h, _p(d), _p(f))
        return d, f


def random_actions(seed, step, env_index_base, lo, hi, n):
    out = np.zeros(n, np.int32)
    lib().or_random_actions(seed, step, env_index_base, lo, hi, _p(out), n)
    return out


def philox(seed, env, step):
    out = np.zeros(4, np.uint32)
    lib().or_philox(seed, env, step, _p(out))
    return out


# ------------------------------------------------------------------------------------------------
# classic suite (BASELINE config 5): oracle/gw_classic_oracle.c
def _bind_classic():
    L = lib()
    if getattr(L, "_classic_bound", False):
        return L
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_uint64]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_coin_override.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_set_dried_override.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_policies.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_shape.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 8
    L.orc_step.argtypes = [C.c_void_p] + [C.c_void_p] * 8
    L.orc_observe.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.orc_layers.argtypes = [C.c_void_p, C.c_void_p]
    L._classic_bound = True
    return L


class ClassicOracle(object):
    """A mixed batch of classic-suite environments: specs[t] repeated counts[t] times."""

    def __init__(self, specs, counts, env_index_base=0, seed=0):
        L = _bind_classic()
        self.specs, self.counts = list(specs), [int(c) for c in counts]
        self.n = sum(self.counts)
        cfg_t = type(self.specs[0].config) * len(self.specs)
        self._cfgs = cfg_t(*[s.config for s in self.specs])
        cnt = (C.c_int64 * len(self.counts))(*self.counts)
        self._h = L.orc_create(C.byref(self._cfgs), len(self.specs), cnt, int(env_index_base), int(seed))
        if not self._h:
            raise ValueError("classic oracle rejected the configuration")
        h, w = C.c_int32(), C.c_int32()
        L.orc_shape(self._h, C.byref(h), C.byref(w))
        self.hmax, self.wmax = h.value, w.value
        n = self.n
        self.board = np.zeros((n, self.hmax, self.wmax), np.uint8)
        self.value_board = np.zeros((n, self.hmax, self.wmax), np.float32)
        self.reward = np.zeros((n, 2), np.float32)
        self.terminated = np.zeros(n, np.uint8)
        self.step_type = np.zeros(n, np.uint8)
        self.reason = np.full(n, -1, np.int8)
        self.actual = np.full(n, -1, np.int8)
        self._coins = None

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _outs(self):
        return [_p(self.board), _p(self.value_board), _p(self.reward), _p(self.terminated), _p(self.step_type), _p(self.reason),
                _p(self.actual)]

    def set_coin_override(self, coins):
        self._coins = None if coins is None else np.ascontiguousarray(coins, np.uint8)
        lib().orc_set_coin_override(self._h, _p(self._coins))

    def set_dried_override(self, dried):
        """uint16 [n]: the tomato games' per-frame draws of the next call as a mask over the tomato cells (0xFFFF = Philox)."""
        self._dried = None if dried is None else np.ascontiguousarray(dried, np.uint16)
        lib().orc_set_dried_override(self._h, _p(self._dried))

    def policies(self):
        """friend_foe: float64 [n, 3, 2], the PolicyEstimator.policy vectors (friend, neutral, adversary) of every environment."""
        out = np.zeros((self.n, 3, 2), np.float64)
        lib().orc_policies(self._h, _p(out))
        return out

    def crop(self, which, i, spec):
        """The H x W board of environment i out of its 64-entry row (pitch 8, or dense for maps wider than 8)."""
        row = getattr(self, which)[i].reshape(-1)
        H, W = spec.height, spec.width
        if W > self.wmax:
            assert not row[H * W:].any()
            return row[:H * W].reshape(H, W)
        full = row.reshape(self.hmax, self.wmax)
        assert not full[H:, :].any() and not full[:, W:].any()
        return full[:H, :W]

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orc_reset(self._h, _p(m), *self._outs())

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.int32)
        assert a.shape == (self.n,)
        lib().orc_step(self._h, _p(a), *self._outs())

    def observe(self):
        n = self.n
        out = dict(ret=np.zeros(n, np.int32), hidden=np.zeros(n, np.int32), frame=np.zeros(n, np.int32),
                   pos=np.zeros((n, 2), np.int16), coin=np.zeros(n, np.int8))
        lib().orc_observe(self._h, _p(out["ret"]), _p(out["hidden"]), _p(out["frame"]), _p(out["pos"]), _p(out["coin"]))
        return out


    def layers(self):
        """uint8 [n, 16, hmax, wmax]: obs['layers'] of the MO re-wrappings in layer_order, board-row layout (crop_layers)."""
        out = np.zeros((self.n, 16, self.hmax, self.wmax), np.uint8)
        lib().orc_layers(self._h, _p(out))
        return out

    def crop_layers(self, layers, i, spec):
        """[L, H, W] cube of environment i (maps no wider than 8)."""
        return layers[i, :len(spec.layer_order), :spec.height, :spec.width]


# ------------------------------------------------------------------------------------------------
# aintelope_savanna: oracle/gw_savanna_oracle.c
class SavannaOracle(object):
    def __init__(self, spec, n_envs, env_index_base=0, seed=0):
        L = lib()
        L.orv_create.restype = C.c_void_p
        L.orv_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64]
        L.orv_destroy.argtypes = [C.c_void_p]
        L.orv_set_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orv_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.orv_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 7
        L.orv_observe.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        self.spec, self.n = spec, int(n_envs)
        self._h = L.orv_create(C.byref(spec.config), self.n, int(env_index_base), int(seed))
        if not self._h:
            raise ValueError("savanna oracle rejected the configuration")
        n, cells, Lr, R, V = self.n, spec.cells, spec.n_layers, spec.n_rewards, spec.view
        self.board = np.zeros((n, spec.height, spec.width), np.uint8)
        self.cube = np.zeros((n, Lr, spec.height, spec.width), np.uint8)
        self.crop = np.zeros((n, 2, V, V), np.uint8)
        self.lcrop = np.zeros((n, 2, Lr, V, V), np.uint8)
        self.reward = np.zeros((n, 2, R), np.float32)
        self.terminated = np.zeros((n, 2), np.uint8)
        self.step_type = np.zeros((n, 2), np.uint8)
        # every environment plays its own layout; default: the canonical art, static
        self.maps = np.tile(np.frombuffer("".join(spec.art).encode(), np.uint8), (n, 1)).copy()
        self.set_maps(self.maps, 0)

    def set_maps(self, maps, mode):
        self.maps = np.ascontiguousarray(maps, np.uint8)
        assert self.maps.shape == (self.n, self.spec.cells)
        lib().orv_set_maps(self._h, _p(self.maps), int(mode))

    def close(self):
        if self._h:
            lib().orv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _outs(self):
        return [_p(self.board), _p(self.cube), _p(self.crop), _p(self.lcrop), _p(self.reward), _p(self.terminated), _p(self.step_type)]

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orv_reset(self._h, _p(m), *self._outs())

    def step(self, actions, order=None, draws=None):
        a = np.ascontiguousarray(actions, np.int32)
        assert a.shape == (self.n, 2)
        o = None if order is None else np.ascontiguousarray(order, np.int32)
        d = None if draws is None else np.ascontiguousarray(draws, np.float64)
        lib().orv_step(self._h, _p(a), _p(o), _p(d), 0 if d is None else d.shape[1], *self._outs())

    def observe(self):
        n, R = self.n, self.spec.n_rewards
        out = dict(metrics=np.zeros((n, 24), np.float64), cumulative=np.zeros((n, 2, R), np.float32), frame=np.zeros(n, np.int32),
                   pos=np.zeros((n, 2, 2), np.int16), directions=np.zeros((n, 2, 2), np.int8))
        lib().orv_observe(self._h, _p(out["metrics"]), _p(out["cumulative"]), _p(out["frame"]), _p(out["pos"]), _p(out["directions"]))
        return out


# ------------------------------------------------------------------------------------------------
# side_effects_sokoban on its big maps: oracle/gw_sokoban_oracle.c
class SokobanOracle(object):
    def __init__(self, spec, n_envs):
        L = lib()
        L.ors_create.restype = C.c_void_p
        L.ors_create.argtypes = [C.c_void_p, C.c_int64]
        L.ors_destroy.argtypes = [C.c_void_p]
        L.ors_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.ors_step.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.ors_observe.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        self.spec, self.n = spec, int(n_envs)
        self._h = L.ors_create(C.byref(spec.config), self.n)
        if not self._h:
            raise ValueError("sokoban oracle rejected the configuration")
        n = self.n
        self.board = np.zeros((n, 128), np.uint8)
        self.value_board = np.zeros((n, 128), np.float32)
        self.reward = np.zeros((n, 2), np.float32)
        self.terminated = np.zeros(n, np.uint8)
        self.step_type = np.zeros(n, np.uint8)
        self.reason = np.full(n, -1, np.int8)
        self.actual = np.full(n, -1, np.int8)

    def close(self):
        if self._h:
            lib().ors_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _outs(self):
        return [_p(self.board), _p(self.value_board), _p(self.reward), _p(self.terminated), _p(self.step_type), _p(self.reason),
                _p(self.actual)]

    def crop(self, which, i):
        H, W = self.spec.height, self.spec.width
        row = getattr(self, which)[i]
        assert not row[H * W:].any()
        return row[:H * W].reshape(H, W)

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().ors_reset(self._h, _p(m), *self._outs())

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.int32)
        assert a.shape == (self.n,)
        lib().ors_step(self._h, _p(a), *self._outs())

    def observe(self):
        n = self.n
        out = dict(cumulative=np.zeros((n, 2), np.int32), frame=np.zeros(n, np.int32), pos=np.zeros((n, 2), np.int16),
                   boxes=np.zeros((n, 3), np.uint8), coins=np.zeros(n, np.uint8))
        lib().ors_observe(self._h, _p(out["cumulative"]), _p(out["frame"]), _p(out["pos"]), _p(out["boxes"]), _p(out["coins"]))
        return out


# ------------------------------------------------------------------------------------------------
# firemaker_ex_ma (BASELINE config 4): oracle/gw_firemaker_oracle.c
class FiremakerOracle(object):
    def __init__(self, spec, n_envs, env_index_base=0, seed=0):
        L = lib()
        L.orf_create.restype = C.c_void_p
        L.orf_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64]
        L.orf_destroy.argtypes = [C.c_void_p]
        L.orf_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 11
        L.orf_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 10
        L.orf_observe.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        self.spec, self.n = spec, int(n_envs)
        self._h = L.orf_create(C.byref(spec.config), self.n, int(env_index_base), int(seed))
        n = self.n
        self.board = np.zeros((n, 17, 17), np.uint8)
        self.cube = np.zeros((n, 9, 17, 17), np.uint8)
        self.crop_w = np.zeros((n, 2, 5, 5), np.uint8)
        self.crop_s = np.zeros((n, 33, 33), np.uint8)
        self.lcrop_w = np.zeros((n, 2, 9, 5, 5), np.uint8)
        self.lcrop_s = np.zeros((n, 9, 33, 33), np.uint8)
        self.reward_w = np.zeros((n, 2, 2), np.float32)
        self.reward_s = np.zeros((n, 3), np.float32)
        self.terminated = np.zeros((n, 3), np.uint8)
        self.step_type = np.zeros((n, 3), np.uint8)

    def close(self):
        if self._h:
            lib().orf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _outs(self):
        return [_p(x) for x in (self.board, self.cube, self.crop_w, self.crop_s, self.lcrop_w, self.lcrop_s, self.reward_w,
                                self.reward_s, self.terminated, self.step_type)]

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().orf_reset(self._h, _p(m), *self._outs())

    def step(self, actions, order=None, draws=None):
        a = np.ascontiguousarray(actions, np.int32)
        assert a.shape == (self.n, 3)
        o = None if order is None else np.ascontiguousarray(order, np.int32)
        d = None if draws is None else np.ascontiguousarray(draws, np.float64)
        stride = 0 if d is None else d.shape[1]
        lib().orf_step(self._h, _p(a), _p(o), _p(d), stride, *self._outs())

    def observe(self):
        n = self.n
        out = dict(metrics=np.zeros((n, 16), np.float64), cumulative=np.zeros((n, 7), np.float32), frame=np.zeros(n, np.int32),
                   pos=np.zeros((n, 3, 2), np.int16), ext_fires=np.zeros(n, np.int32))
        lib().orf_observe(self._h, _p(out["metrics"]), _p(out["cumulative"]), _p(out["frame"]), _p(out["pos"]), _p(out["ext_fires"]))
        out["directions"] = np.zeros((n, 3, 2), np.int8)          # action direction, observation direction per agent
        lib().orf_directions.argtypes = [C.c_void_p, C.c_void_p]
        lib().orf_directions(self._h, _p(out["directions"]))
        return out


class IslandMaOracle(object):
    """N independent island_navigation_ex_ma games (2 agents), oracle/gw_island_ma_oracle.c."""

    def __init__(self, spec, n_envs, env_index_base=0, seed=0):
        L = lib()
        L.ori_create.restype = C.c_void_p
        L.ori_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64]
        L.ori_destroy.argtypes = [C.c_void_p]
        L.ori_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.ori_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_void_p] * 7
        L.ori_observe.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ori_set_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.spec, self.n = spec, int(n_envs)
        self._h = L.ori_create(C.byref(spec.config), self.n, int(env_index_base), int(seed))
        if not self._h:
            raise ValueError("oracle rejected the configuration")
        self.maps = None
        n, H, W, Lr, R = self.n, spec.height, spec.width, spec.n_layers, spec.n_rewards
        self.board = np.zeros((n, H, W), np.uint8)
        self.cube = np.zeros((n, Lr, H, W), np.uint8)
        self.crop = np.zeros((n, 2, 5, 5), np.uint8)
        self.lcrop = np.zeros((n, 2, Lr, 5, 5), np.uint8)
        self.reward = np.zeros((n, 2, R), np.float32)
        self.terminated = np.zeros((n, 2), np.uint8)
        self.step_type = np.zeros((n, 2), np.uint8)

    def close(self):
        if self._h:
            lib().ori_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _outs(self):
        return [_p(x) for x in (self.board, self.cube, self.crop, self.lcrop, self.reward, self.terminated, self.step_type)]

    def set_maps(self, maps, mode=0):
        """maps: uint8 [n, H, W] (kept and, in the shuffle modes, rewritten in place) or None; mode = GwImaMapMode"""
        if maps is not None:
            assert maps.dtype == np.uint8 and maps.flags["C_CONTIGUOUS"] and maps.shape == (self.n, self.spec.height, self.spec.width)
        self.maps = maps
        lib().ori_set_maps(self._h, _p(maps), int(mode))

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().ori_reset(self._h, _p(m), *self._outs())

    def step(self, actions, order=None):
        a = np.ascontiguousarray(actions, np.int32)
        assert a.shape == (self.n, 2)
        o = None if order is None else np.ascontiguousarray(order, np.int32)
        lib().ori_step(self._h, _p(a), _p(o), *self._outs())

    def observe(self):
        n, R = self.n, self.spec.n_rewards
        out = dict(metrics=np.zeros((n, 16), np.float64), cumulative=np.zeros((n, 2, R), np.float32), frame=np.zeros(n, np.int32),
                   pos=np.zeros((n, 2, 2), np.int16), directions=np.zeros((n, 2, 2), np.int8))
        lib().ori_observe(self._h, _p(out["metrics"]), _p(out["cumulative"]), _p(out["frame"]), _p(out["pos"]), _p(out["directions"]))
        return out
