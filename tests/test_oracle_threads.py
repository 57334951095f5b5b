"""The oracles' multi-threaded batch loops (or_set_threads / or_parallel_for, plain pthreads) give the very same
results as the scalar loop, and the checksum the BASELINE-size parity tests rely on is the documented sum.
CPU only."""
import ctypes as C

import numpy as np

import scale_util


def _run(oracle_lib, make, steps, act_shape, hi, threads):
    L = scale_util.bind(oracle_lib)
    L.or_set_threads(threads)
    try:
        orc = make()
        orc.reset()
        rng = np.random.default_rng(5)
        outs = []
        for t in range(steps):
            a = rng.integers(0, hi + 1, size=act_shape).astype(np.int32)
            orc.step(a)
            outs.append([np.array(getattr(orc, k)) for k in ("board", "terminated", "step_type")] +
                        [np.array(getattr(orc, k)) for k in ("reward", "reward_w", "reward_s", "lcrop", "lcrop_s", "cube") if hasattr(orc, k)
                         and getattr(orc, k) is not None])
        orc.close()
        return outs
    finally:
        L.or_set_threads(1)


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        for u, v in zip(x, y):
            np.testing.assert_array_equal(u, v)


def test_threaded_oracles_equal_the_scalar_loops(oracle_lib):
    from ai_safety_gridworlds_b200 import make_spec
    n = 777
    fm = make_spec("firemaker_ex_ma", autoreset_mode=1, max_iterations=60, amount_agents=3)
    ima = make_spec("island_navigation_ex_ma", autoreset_mode=1)
    sav = make_spec("aintelope_savanna", autoreset_mode=1)
    sok = make_spec("side_effects_sokoban", level=1, autoreset_mode=1)
    cls = [make_spec(nm, autoreset_mode=1) for nm in ("safe_interruptibility", "side_effects_sokoban", "absent_supervisor", "conveyor_belt", "whisky_gold")]
    cases = [
        (lambda: oracle_lib.FiremakerOracle(fm, n, seed=3), 30, (n, 3), 4),
        (lambda: oracle_lib.IslandMaOracle(ima, n, seed=3), 60, (n, 2), 4),
        (lambda: oracle_lib.SavannaOracle(sav, n, seed=3), 40, (n, 2), 4),
        (lambda: oracle_lib.SokobanOracle(sok, n), 60, (n,), 4),
        (lambda: oracle_lib.ClassicOracle(cls, [155, 156, 155, 156, 155], seed=3), 120, (n,), 4),
    ]
    for make, steps, shape, hi in cases:
        _same(_run(oracle_lib, make, steps, shape, hi, 1), _run(oracle_lib, make, steps, shape, hi, 5))


def test_threaded_random_actions_and_checksum(oracle_lib):
    L = scale_util.bind(oracle_lib)
    a1 = oracle_lib.random_actions(9, 4, 100, 0, 4, 10001)
    L.or_set_threads(7)
    try:
        a7 = oracle_lib.random_actions(9, 4, 100, 0, 4, 10001)
        x = (np.arange(100003, dtype=np.uint64) * np.uint64(2654435761)) ^ np.uint64(0xABCDEF)
        h7 = scale_util.host_checksum(L, x)
    finally:
        L.or_set_threads(1)
    np.testing.assert_array_equal(a1, a7)
    h1 = scale_util.host_checksum(L, x)
    j = np.arange(x.size, dtype=np.uint64)
    w = (j * np.uint64(scale_util.K1) + np.uint64(scale_util.K2)) | np.uint64(1)
    want = int((x * w).sum(dtype=np.uint64))
    assert h1 == h7 == want
    # a buffer whose size is not a multiple of 8 bytes is zero-padded
    y = np.arange(13, dtype=np.uint8)
    pad = np.zeros(16, np.uint8)
    pad[:13] = y
    assert scale_util.host_checksum(L, y) == scale_util.host_checksum(L, pad)


def test_device_checksum_formula_matches_host_on_cpu_tensors(oracle_lib):
    """The torch expression of tests/scale_util.device_checksum (run here on a CPU tensor) is the same sum."""
    import torch
    L = scale_util.bind(oracle_lib)
    rng = np.random.default_rng(1)
    for shape, dt in (((1000, 48), np.uint8), ((333, 10), np.float32), ((77,), np.int8), ((64, 9, 7, 7), np.uint8)):
        x = rng.integers(0, 255, size=shape).astype(dt)
        assert scale_util.device_checksum(torch.from_numpy(x)) == scale_util.host_checksum(L, x)
