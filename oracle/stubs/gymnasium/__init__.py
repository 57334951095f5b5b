"""Minimal offline stand-in for the `gymnasium` package (TEST INFRASTRUCTURE ONLY).

The real gymnasium is not installable in this image (no network).  The reference
(/root/reference, levitation-opensource/ai-safety-gridworlds) imports only the
names below (helpers/gridworld_gym_env.py:17-30, shared/safety_game_mo.py:49-54).
This stub exists so that `oracle/record.py` can run the UNMODIFIED reference and
write golden traces; nothing in the product package imports it.
"""
from . import error, spaces, utils  # noqa: F401


class Space(object):
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = dtype
        self._np_random = None

    @property
    def shape(self):
        return self._shape

    def sample(self, mask=None):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError


class Env(object):
    metadata = {}
    render_mode = None

    def __init__(self):
        pass

    def reset(self, *a, **k):
        raise NotImplementedError

    def step(self, *a, **k):
        raise NotImplementedError

    def render(self, *a, **k):
        raise NotImplementedError

    def close(self):
        pass


spaces._bind(Space)
