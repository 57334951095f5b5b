"""CPU-side checks: the C-ABI library loads and exports every symbol include/gwsim.h declares, the
ctypes mirror matches the compiled structs, compute entry points fail loudly without a GPU, and the
multi-rank host logic (sharded Philox streams, raw-statistics all-reduce) works over gloo."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from ai_safety_gridworlds_b200 import _abi
    return _abi.load()


def test_library_exports_every_declared_symbol(lib):
    from ai_safety_gridworlds_b200 import _abi
    header = "".join(open(os.path.join(ROOT, "include", f)).read() for f in ("gwsim.h", "gwsim_fm.h", "gwsim_ima.h", "gwsim_sok.h", "gwsim_sav.h"))
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gw_[a-z_]+)\s*\(", header))
    assert len(declared) >= 33
    bound = {name for name, _, _ in _abi.SYMBOLS + _abi.FM_SYMBOLS + _abi.IMA_SYMBOLS + _abi.SOK_SYMBOLS + _abi.SAV_SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_mirror_and_constants(lib):
    from ai_safety_gridworlds_b200 import _abi
    assert lib.gw_abi_version() == _abi.GW_ABI_VERSION
    assert lib.gw_config_bytes() == C.sizeof(_abi.GwConfig)
    assert lib.gw_fm_config_bytes() == C.sizeof(_abi.GwFmConfig)
    assert lib.gw_sok_config_bytes() == C.sizeof(_abi.GwSokConfig)
    assert lib.gw_sav_config_bytes() == C.sizeof(_abi.GwSavConfig)
    sav_header = open(os.path.join(ROOT, "include", "gwsim_sav.h")).read()
    for const in ("GW_SAV_MAX_CELLS", "GW_SAV_AGENTS", "GW_SAV_MAX_LAYERS", "GW_SAV_MAX_REWARDS", "GW_SAV_METRICS", "GW_SAV_EVENTS", "GW_SAV_STATE_BYTES"):
        m = re.search(r"#define %s (\d+)" % const, sav_header)
        assert m and int(m.group(1)) == getattr(_abi, const), const
    sok_header = open(os.path.join(ROOT, "include", "gwsim_sok.h")).read()
    for const in ("GW_SOK_MAX_CELLS", "GW_SOK_MAX_BOXES", "GW_SOK_MAX_COINS", "GW_SOK_STATS_LEN"):
        m = re.search(r"#define %s (\d+)" % const, sok_header)
        assert m and int(m.group(1)) == getattr(_abi, const), const
    fm_header = open(os.path.join(ROOT, "include", "gwsim_fm.h")).read()
    for const in ("GW_FM_SIDE", "GW_FM_AGENTS", "GW_FM_LAYERS", "GW_FM_METRICS", "GW_FM_STATE_WORDS", "GW_FM_MAX_DRAWS"):
        m = re.search(r"#define %s (\d+)" % const, fm_header)
        assert m and int(m.group(1)) == getattr(_abi, const), const
    header = open(os.path.join(ROOT, "include", "gwsim.h")).read()
    for const in ("GW_MAX_CELLS", "GW_MAX_LAYERS", "GW_MAX_REWARDS", "GW_MAX_EVENTS", "GW_MAX_METRICS", "GW_STATS_RAW_LEN"):
        m = re.search(r"#define %s (\d+)" % const, header)
        assert m and int(m.group(1)) == getattr(_abi, const), const


def test_state_size_queries(lib):
    from ai_safety_gridworlds_b200 import make_spec
    isl = make_spec("island_navigation_ex")
    assert lib.gw_state_words(C.byref(isl.config)) == 5
    assert lib.gw_state_bytes(C.byref(isl.config), 1000) == 5 * 16 * 1024      # whole 32-environment chunks
    prop = make_spec("island_navigation_ex", use_satiation_proportional_reward=True)
    assert lib.gw_state_words(C.byref(prop.config)) == 7
    boat = make_spec("boat_race_ex", level=3)
    assert lib.gw_state_words(C.byref(boat.config)) == 5
    boat16 = make_spec("boat_race_ex", level=3, max_iterations=1000)
    assert lib.gw_state_words(C.byref(boat16.config)) == 9
    bad = make_spec("island_navigation_ex")
    bad.config.abi_version = 99
    assert lib.gw_state_words(C.byref(bad.config)) == 0
    assert b"ABI" in lib.gw_last_error()


def test_no_cpu_fallback(lib):
    """Without a CUDA device gw_create returns GW_ERR_NO_DEVICE and VectorEnv raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ai_safety_gridworlds_b200 import _abi, make_spec
    spec = make_spec("island_navigation_ex")
    h = C.c_void_p()
    rc = lib.gw_create(C.byref(spec.config), 128, 0, 0, C.byref(h))
    assert rc == _abi.GW_ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.gw_last_error()
    from ai_safety_gridworlds_b200.vector_env import VectorEnv
    with pytest.raises(_abi.GwError):
        VectorEnv("island_navigation_ex", 128)


def test_product_never_imports_the_oracle():
    """The product package must not route through oracle/ (grep the sources)."""
    pkg = os.path.join(ROOT, "ai_safety_gridworlds_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "gw_oracle" not in text and "import oracle" not in text, f


def test_stats_finalize_is_linear_in_events(lib):
    from ai_safety_gridworlds_b200 import _abi, make_spec
    from ai_safety_gridworlds_b200.vector_env import finalize_stats
    spec = make_spec("island_navigation_ex")
    raw = np.zeros(_abi.GW_STATS_RAW_LEN)
    raw[_abi.GW_RAW_ENV_STEPS], raw[_abi.GW_RAW_EPISODES], raw[_abi.GW_RAW_LENGTH_SUM] = 700, 100, 650
    raw[_abi.GW_RAW_REASON0] = 90
    raw[_abi.GW_RAW_REASON0 + 1] = 10
    raw[_abi.GW_RAW_EVENT0 + _abi.ISL_E["MOVEMENT"]] = 500
    raw[_abi.GW_RAW_EVENT0 + _abi.ISL_E["DANGER_TILE"]] = 90
    raw[_abi.GW_RAW_EVENT0 + _abi.ISL_E["GOLD"]] = 7
    st = finalize_stats(spec, raw)
    assert st["episodes"] == 100 and st["env_steps"] == 700 and st["mean_length"] == 6.5
    assert st["reasons"]["terminated"] == 90 and st["reasons"]["max_steps"] == 10
    assert st["return_sum"]["MOVEMENT_REWARD"] == -500
    assert st["return_sum"]["DANGER_TILE_REWARD"] == -4500
    assert st["return_sum"]["GOLD_REWARD"] == 280
    assert st["mean_return"]["GOLD_REWARD"] == pytest.approx(2.8)


WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
from ai_safety_gridworlds_b200 import _abi, make_spec
from ai_safety_gridworlds_b200.vector_env import finalize_stats
from ai_safety_gridworlds_b200.parallel import shard_range, all_reduce_raw_stats
from oracle import pyoracle

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N = 1000
lo, hi = shard_range(N, rank, world)
# (1) sharded Philox streams: the union of the shards' actions equals the unsharded stream
mine = pyoracle.random_actions(3, 17, lo, 0, 4, hi - lo)
parts = [None] * world
dist.all_gather_object(parts, (lo, mine))
full = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
assert np.array_equal(full, pyoracle.random_actions(3, 17, 0, 0, 4, N))
# (2) raw statistics all-reduce: integer-valued doubles, SUM, identical on every rank
raw = torch.zeros(_abi.GW_STATS_RAW_LEN, dtype=torch.float64)
raw[_abi.GW_RAW_EPISODES] = 10 * (rank + 1)
raw[_abi.GW_RAW_ENV_STEPS] = 100 * (rank + 1)
raw[_abi.GW_RAW_LENGTH_SUM] = 70 * (rank + 1)
raw[_abi.GW_RAW_EVENT0 + _abi.BOAT_E["CLOCKWISE"]] = -3 * (rank + 1)
raw[_abi.GW_RAW_EVENT0 + _abi.BOAT_E["REPETITION"]] = 2.0 ** 40 + rank
total = all_reduce_raw_stats(raw)
spec = make_spec("boat_race_ex", level=3)
st = finalize_stats(spec, total.numpy())
tri = world * (world + 1) // 2
assert st["episodes"] == 10 * tri and st["env_steps"] == 100 * tri and st["length_sum"] == 70 * tri
assert st["return_sum"]["CLOCKWISE_REWARD"] == -9.0 * tri
assert st["return_sum"]["REPETITION_REWARD"] == -(world * 2.0 ** 40 + sum(range(world)))
dist.barrier()
dist.destroy_process_group()
print("rank %%d ok" %% rank)
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_host_logic_over_gloo(lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "rank 0 ok" in p.stdout and "rank 1 ok" in p.stdout


def test_reference_rejected_flags_are_rejected():
    """A randomised flag combination (oracle/record.py fuzz case 14) that the reference itself fails on -- level 0 has no drink
    tile, so DRINK_DEFICIENCY_REWARD is not an enabled dimension, yet DRINK_DEFICIENCY_INITIAL = -2 posts it at the first step
    (mo_reward.py:184-203) -- raises the same ValueError from the spec compiler, eagerly."""
    from ai_safety_gridworlds_b200 import make_spec
    msg = "Reward DRINK_DEFICIENCY_REWARD is not enabled but is still included in mo_reward with nonzero value"
    with pytest.raises(ValueError, match=msg):
        make_spec("island_navigation_ex", level=0, penalise_oversatiation=False, DRINK_DEFICIENCY_INITIAL=-2)
    with pytest.raises(ValueError, match=msg):
        make_spec("island_navigation_ex", level=0)       # verified against the reference: default flags fail on level 0 at the first step
    make_spec("island_navigation_ex", level=0, penalise_oversatiation=False)      # satiation stays at 0: nothing is posted


def test_island_ma_map_resizing_spec():
    """Map resizing (shared/safety_game_mo_base.py:984-1036): the resized board holds tile_type_counts (the two agents) inside a
    border of what_lies_outside, reward dimensions and metrics still follow the level's own map; the reference's asserts apply."""
    from ai_safety_gridworlds_b200 import make_spec
    base = make_spec("island_navigation_ex_ma", map_randomization_frequency=3)
    spec = make_spec("island_navigation_ex_ma", map_randomization_frequency=3, map_width=7, map_height=6)
    assert spec.art == ["WWWWWWW", "W12   W", "W     W", "W     W", "W     W", "WWWWWWW"]
    assert spec.reward_keys == base.reward_keys and spec.metric_names == base.metric_names
    assert spec.layer_order == [c for c in base.layer_order if c != "#"]          # no wall tile is left on the board
    same = make_spec("island_navigation_ex_ma", map_randomization_frequency=3, map_width=base.width, map_height=base.height)
    assert same.art == base.art                                                     # the level's own size: nothing is resized
    only_w = make_spec("island_navigation_ex_ma", map_randomization_frequency=1, map_width=5)
    assert (only_w.height, only_w.width) == (base.height, 5)
    with pytest.raises(AssertionError):
        make_spec("island_navigation_ex_ma", map_width=7, map_height=6)             # resizing needs map randomisation
    with pytest.raises(AssertionError):
        make_spec("island_navigation_ex_ma", map_randomization_frequency=3, map_width=2, map_height=6)
    with pytest.raises(ValueError):
        make_spec("island_navigation_ex_ma", level=4, map_randomization_frequency=3, map_width=7, map_height=6)   # no water reward dimension
