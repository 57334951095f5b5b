"""BASELINE-size parity: the CUDA path against the CPU oracle (not against itself) at the sizes
BASELINE.json quotes -- config 3 (1,048,576 island_navigation_ex environments), config 2 (65,536
boat_race_ex), config 4 (262,144 firemaker_ex_ma games, Philox order and fire draws) and config 5
(1,048,576 environments of the original suite in one mixed batch).

Every output tensor of every step is compared through a position-weighted 64-bit checksum computed
on the device (torch int64) and on the host (oracle/gw_oracle.c or_checksum64) -- see
tests/scale_util.py -- and the state-derived quantities (frames, positions, metrics, cumulative
rewards, resource availabilities and regrowth fractions) are compared element-wise at checkpoints.
Integer and byte tensors are bit-exact.  Float32 reward rows are expected bit-equal too; if a row
differs the fallback is BASELINE.json's tolerance (1e-6 relative) and the test reports how often it
was needed.  The oracle steps on all host cores (plain pthreads; environments are independent).

This is the test SURVEY section 7 asks for about CUDA pow vs glibc pow under int() truncation
(island_navigation_ex.py:647-656): 2e8 environment steps of regrowth are compared here.
"""
import numpy as np
import pytest

import scale_util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.cpu().numpy()


def _mo_case(oracle_lib, name, kwargs, n, steps, mode, lo, hi, seed, checkpoints):
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.vector_env import VectorEnv
    chk = scale_util.Checker(oracle_lib)
    threads = scale_util.host_threads()
    spec = make_spec(name, autoreset_mode=mode, **kwargs)
    env = VectorEnv(spec, n, autoreset_mode=mode, want_value_board=False)
    orc = oracle_lib.Oracle(spec, n, want_value_board=False)
    orc.reset()
    chk.same("board", env.board, orc.board, "reset")
    chk.same("cube", env.cube, orc.cube, "reset")
    episodes = 0
    for t in range(steps):
        a = env.random_actions(seed=seed, step=t, lo=lo, hi=hi)
        a_ref = oracle_lib.random_actions(seed, t, 0, lo, hi, n)
        env.step(a)
        orc.step(a_ref, n_threads=threads)
        ctx = "%s n=%d mode=%d t=%d" % (name, n, mode, t)
        chk.same("actions", a, a_ref, ctx)
        chk.same("board", env.board, orc.board, ctx)
        chk.same("cube", env.cube, orc.cube, ctx)
        chk.same("reward", env.reward, orc.reward, ctx, rtol=1e-6)
        chk.same("terminated", env.terminated, orc.terminated, ctx)
        chk.same("step_type", env.step_type, orc.step_type, ctx)
        chk.same("reason", env.reason, orc.reason, ctx)
        episodes += int(orc.terminated.sum())
        if t in checkpoints or t == steps - 1:
            ex, ox = env.observe(), orc.observe()
            np.testing.assert_array_equal(_np(ex["frame"]), ox["frame"], err_msg=ctx)
            np.testing.assert_array_equal(_np(ex["pos"]), ox["pos"], err_msg=ctx)
            np.testing.assert_array_equal(_np(ex["safety"]), ox["safety"], err_msg=ctx)
            np.testing.assert_allclose(_np(ex["cumulative"]), ox["cumulative"], rtol=1e-6, atol=1e-5, err_msg=ctx)
            if spec.metric_names:
                np.testing.assert_allclose(_np(ex["metrics"]), ox["metrics"], rtol=1e-12, atol=1e-12, err_msg=ctx)
            if name == "island_navigation_ex":
                gd, gf = env.peek_fractions()
                od, of_ = orc.fractions()
                np.testing.assert_allclose(_np(gd), od, rtol=1e-12, atol=1e-13, err_msg=ctx)
                np.testing.assert_allclose(_np(gf), of_, rtol=1e-12, atol=1e-13, err_msg=ctx)
    st = env.stats()
    assert st["episodes"] == episodes and episodes > 0
    if mode == 1:
        assert st["env_steps"] == n * steps
    print("scale parity %s: %d envs x %d steps, %d episodes, %.1f GB compared by checksum, %d float fallbacks"
          % (name, n, steps, episodes, chk.compared_bytes / 1e9, chk.float_fallbacks))
    assert chk.float_fallbacks == 0, "float32 reward rows were not bit-equal in %d steps (within 1e-6 though)" % chk.float_fallbacks
    env.close()
    orc.close()
    chk.close()


def test_config3_island_navigation_ex_one_million_envs_vs_oracle(oracle_lib):
    """BASELINE config 3 / the headline workload: 1,048,576 environments x 200 steps (two full
    100-step episodes per environment plus the early endings), auto-reset in the ending step."""
    _mo_case(oracle_lib, "island_navigation_ex", {}, 1 << 20, 200, 1, 0, 4, seed=101, checkpoints={0, 49, 99, 100, 149})


def test_config3_next_call_autoreset_vs_oracle(oracle_lib):
    """The reference's own reset semantics (the call after LAST rebuilds the game), 262,144 environments x 120 steps."""
    _mo_case(oracle_lib, "island_navigation_ex", {}, 1 << 18, 120, 0, 0, 4, seed=102, checkpoints={60, 100, 101})


def test_config2_boat_race_ex_65536_envs_vs_oracle(oracle_lib):
    """BASELINE config 2 as written: boat_race_ex level 3 with the repetition, iteration and human-tile
    penalties, 65,536 environments x 300 steps."""
    _mo_case(oracle_lib, "boat_race_ex", {"level": 3}, 1 << 16, 300, 1, 0, 4, seed=103, checkpoints={99, 100, 199})


def test_config4_firemaker_262144_envs_vs_oracle(oracle_lib):
    """BASELINE config 4: 262,144 three-agent games x 36 parallel steps, shuffle orders and fire draws from the
    shared Philox streams (a 90-frame cut-off, so every game ends and restarts once inside the run)."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    chk = scale_util.Checker(oracle_lib)
    N, T = 1 << 18, 36
    spec = make_spec("firemaker_ex_ma", autoreset_mode=1, max_iterations=90, amount_agents=3)
    env = FiremakerVectorEnv(N, seed=77, autoreset_mode=1, spec=spec)
    orc = oracle_lib.FiremakerOracle(spec, N, seed=77)
    orc.reset()
    dev = env.device
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    fires = 0
    for t in range(T):
        a = torch.randint(0, 5, (N, 3), dtype=torch.int32, device=dev, generator=g)
        env.step(a)
        orc.step(_np(a))
        ctx = "firemaker n=%d t=%d" % (N, t)
        chk.same("board", env.board, orc.board, ctx)
        chk.same("cube", env.cube, orc.cube, ctx)
        chk.same("crop_workers", env.crop_workers, orc.crop_w, ctx)
        chk.same("crop_supervisor", env.crop_supervisor, orc.crop_s, ctx)
        chk.same("lcrop_workers", env.lcrop_workers, orc.lcrop_w, ctx)
        chk.same("lcrop_supervisor", env.lcrop_supervisor, orc.lcrop_s, ctx)
        chk.same("reward_workers", env.reward_workers, orc.reward_w, ctx)
        chk.same("reward_supervisor", env.reward_supervisor, orc.reward_s, ctx)
        chk.same("step_type", env.step_type, orc.step_type, ctx)
        chk.same("terminated", env.terminated, orc.terminated, ctx)
        if t % 12 == 11 or t == T - 1:
            ex, ox = env.observe(), orc.observe()
            for key in ("metrics", "cumulative", "frame", "pos", "ext_fires"):
                np.testing.assert_array_equal(_np(ex[key]), ox[key], err_msg=ctx + " " + key)
            fires = max(fires, int((orc.board == ord("F")).sum()))
    assert fires > 0
    st = env.stats()
    assert st["env_steps"] == N * T and st["episodes"] == N
    print("scale parity firemaker_ex_ma: %d games x %d parallel steps, %.1f GB compared by checksum" % (N, T, chk.compared_bytes / 1e9))
    env.close()
    orc.close()
    chk.close()


CONFIG5 = ["safe_interruptibility", "side_effects_sokoban", "absent_supervisor", "conveyor_belt", "whisky_gold"]


def test_config5_original_suite_one_million_envs_vs_oracle(oracle_lib):
    """BASELINE config 5: equal fifths of the five games in one mixed batch of 1,048,576 environments x 120 steps
    (every game's default max_iterations = 100 is crossed), Philox actions over each game's own range
    plus NOOP, Philox per-episode draws."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.classic_env import ClassicVectorEnv
    chk = scale_util.Checker(oracle_lib)
    N, T = 1 << 20, 120
    specs = [make_spec(nm, autoreset_mode=1) for nm in CONFIG5]
    k = len(specs)
    counts = [N // k] * (k - 1) + [N - (k - 1) * (N // k)]
    env = ClassicVectorEnv(specs, counts, seed=9, autoreset_mode=1, want_value_board=False)
    orc = oracle_lib.ClassicOracle(specs, counts, seed=9)
    orc.reset()
    chk.same("board", env.board, orc.board, "reset")
    episodes = 0
    for t in range(T):
        a = env.random_actions(13, t, lo=0, hi=4)
        a_ref = oracle_lib.random_actions(13, t, 0, 0, 4, N)
        env.step(a)
        orc.step(a_ref)
        ctx = "classic mixed n=%d t=%d" % (N, t)
        chk.same("actions", a, a_ref, ctx)
        chk.same("board", env.board, orc.board, ctx)
        chk.same("reward", env.reward, orc.reward, ctx)
        chk.same("terminated", env.terminated, orc.terminated, ctx)
        chk.same("step_type", env.step_type, orc.step_type, ctx)
        chk.same("reason", env.reason, orc.reason, ctx)
        chk.same("actual", env.actual, orc.actual, ctx)
        episodes += int(orc.terminated.sum())
        if t % 40 == 39:
            ex, ox = env.observe(), orc.observe()
            np.testing.assert_array_equal(_np(ex["cumulative"][:, 0]), ox["ret"].astype(np.float32), err_msg=ctx)
            np.testing.assert_array_equal(_np(ex["cumulative"][:, 1]), ox["hidden"].astype(np.float32), err_msg=ctx)
            np.testing.assert_array_equal(_np(ex["frame"]), ox["frame"], err_msg=ctx)
            np.testing.assert_array_equal(_np(ex["pos"]), ox["pos"], err_msg=ctx)
            np.testing.assert_array_equal(_np(ex["coin"]), ox["coin"], err_msg=ctx)
    st = env.stats()
    assert st["env_steps"] == N * T and st["episodes"] == episodes and episodes > N
    print("scale parity classic mixed: %d envs x %d steps, %d episodes, %.1f GB compared by checksum" % (N, T, episodes, chk.compared_bytes / 1e9))
    env.close()
    orc.close()
    chk.close()
