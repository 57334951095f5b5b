import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _all_golden():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_names():
    """Traces of the multi-objective games (island_navigation_ex, boat_race_ex)."""
    return [n for n in _all_golden() if not n.startswith(("classic_", "firemaker_", "islandma_", "sokoban_big_", "savanna_", "rgb_", "zoo_"))]


def firemaker_golden_names():
    """Traces of firemaker_ex_ma (oracle/record_firemaker.py)."""
    return [n for n in _all_golden() if n.startswith("firemaker_") and not n.startswith("firemaker_aec_")]


def firemaker_aec_golden_names():
    """Traces of firemaker_ex_ma through the AEC wrapper, one engine frame per step (oracle/record_firemaker_aec.py)."""
    return [n for n in _all_golden() if n.startswith("firemaker_aec_")]


def island_ma_golden_names():
    """Traces of island_navigation_ex_ma through the parallel wrapper (oracle/record_island_ma.py)."""
    return [n for n in _all_golden() if n.startswith("islandma_")]


def classic_golden_names():
    """Traces of the original DeepMind suite (oracle/record_classic.py)."""
    return [n for n in _all_golden() if n.startswith("classic_")]


def sokoban_golden_names():
    """Traces of side_effects_sokoban for the gw_sok_* path (levels 0-3; oracle/record_classic.py)."""
    return [n for n in _all_golden() if n.startswith("sokoban_big_")]


def savanna_golden_names():
    """Traces of aintelope_savanna through the parallel wrapper (oracle/record_savanna.py)."""
    return [n for n in _all_golden() if n.startswith("savanna_")]


def rgb_golden_names():
    """Boards + RGB observations recorded from the reference's distiller (oracle/record_rgb.py)."""
    return [n for n in _all_golden() if n.startswith("rgb_")]


def zoo_golden_names():
    """Traces recorded through the reference's PettingZoo parallel wrapper with its wrapper-side options (oracle/record_zoo_wrapper.py)."""
    return [n for n in _all_golden() if n.startswith("zoo_")]


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(d["meta_json"]))
    return d, meta


def spec_for(meta, autoreset_mode=0):
    """Compiles the product-side EnvSpec from the recorded constructor kwargs."""
    from ai_safety_gridworlds_b200 import make_spec
    return make_spec(meta["env"], autoreset_mode=autoreset_mode, **meta["kwargs"])


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle
