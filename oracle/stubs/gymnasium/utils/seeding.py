"""np_random(seed) -> (Generator(PCG64(SeedSequence(seed))), seed), as gymnasium does.

The generator subclass adds `rand(*size)`, which the reference's FireDrape calls
(firemaker_ex_ma.py:615,621) and which only exists on gym 0.22-0.25's
RandomNumberGenerator compatibility shim, where it is `self.random(size)`.
"""
import numpy as np


class RandomNumberGenerator(np.random.Generator):
    def rand(self, *size):
        return self.random(size if size else None)


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, (int, np.integer)) and seed >= 0):
        raise ValueError("seed must be a non-negative integer or None, got %r" % (seed,))
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    rng = RandomNumberGenerator(np.random.PCG64(seed_seq))
    return rng, np_seed
