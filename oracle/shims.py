"""Oracle-side monkeypatches needed to run the unmodified reference under numpy 2 (TEST INFRASTRUCTURE ONLY).

Each patch is the minimal change that lets the reference's own code run here; none alters
the semantics the golden traces pin.

1. `np.Inf` was removed in numpy 2.0; side_effects_sokoban.py:250,256 still uses it.
"""
import numpy as np

if not hasattr(np, "Inf"):
    np.Inf = np.inf
