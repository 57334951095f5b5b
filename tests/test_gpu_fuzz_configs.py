"""Configuration fuzz: the CUDA path against the CPU oracle over RANDOM flag settings (levels, the four boolean game flags,
extraction / deficiency rates, thresholds, regrowth exponents and limits, reward magnitudes, frame cut-offs, both auto-reset
semantics) of island_navigation_ex and boat_race_ex -- the flag space the reference's experiments move in
(island_navigation_ex.py:227-337, boat_race_ex.py:263-327).  The oracle is pinned to the reference by the recorded traces
(flag variants included); this test carries that pin across the flag space.  Byte / integer outputs bit-exact, float32 reward
rows to BASELINE.json's 1e-6, the hidden fp64 resource state to 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.cpu().numpy()


def _island_config(rng):
    kw = dict(level=int(rng.integers(0, 10)), max_iterations=int(rng.integers(8, 130)),
              sustainability_challenge=bool(rng.integers(0, 2)), thirst_hunger_death=bool(rng.integers(0, 2)),
              penalise_oversatiation=bool(rng.integers(0, 2)), use_satiation_proportional_reward=bool(rng.integers(0, 2)))
    if rng.random() < 0.7:
        kw.update(DRINK_EXTRACTION_RATE=float(rng.choice([1.0, 2.5, 6.25, 10.0])), FOOD_EXTRACTION_RATE=float(rng.choice([0.75, 3.0, 10.0])),
                  DRINK_DEFICIENCY_RATE=float(rng.choice([-1.0, -0.3, -0.125, -2.0])), FOOD_DEFICIENCY_RATE=float(rng.choice([-1.0, -0.5, -0.25])),
                  DRINK_DEFICIENCY_LIMIT=float(rng.choice([-20.0, -6.0, -3.0])), FOOD_DEFICIENCY_LIMIT=float(rng.choice([-20.0, -5.0])),
                  DRINK_OVERSATIATION_LIMIT=float(rng.choice([4.0, 1.5, -1.0])), FOOD_OVERSATIATION_LIMIT=float(rng.choice([4.0, 2.0])))
    if rng.random() < 0.7:
        kw.update(DRINK_REGROWTH_EXPONENT=float(rng.choice([1.1, 1.3, 1.05, 2.0])), DRINK_GROWTH_LIMIT=float(rng.choice([20.0, 8.0, 35.0])),
                  FOOD_GROWTH_LIMIT=float(rng.choice([20.0, 6.0, 50.0])), DRINK_AVAILABILITY_INITIAL=float(rng.choice([20.0, 3.0, 1.0])),
                  FOOD_AVAILABILITY_INITIAL=float(rng.choice([20.0, 5.0, 0.5])))
    if rng.random() < 0.5:
        kw.update(MOVEMENT_REWARD="{'MOVEMENT_REWARD': %g}" % rng.choice([-1, -0.25, 0]), GOLD_REWARD="{'GOLD_REWARD': %g}" % rng.choice([40, 7.5]),
                  DRINK_REWARD="{'DRINK_REWARD': %g}" % rng.choice([20, 0.1]), DANGER_TILE_REWARD="{'DANGER_TILE_REWARD': %g}" % rng.choice([-50, -3.75]))
    return "island_navigation_ex", kw


def _boat_config(rng):
    kw = dict(level=int(rng.integers(0, 4)), max_iterations=int(rng.integers(8, 330)))
    if rng.random() < 0.6:
        kw.update(iterations_penalty=bool(rng.integers(0, 2)), repetition_penalty=bool(rng.integers(0, 2)))
    return "boat_race_ex", kw


@pytest.mark.parametrize("seed", range(48))
def test_random_configurations_match_the_oracle(seed, oracle_lib):
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.vector_env import VectorEnv
    rng = np.random.default_rng(1000 + seed)
    name, kw = _island_config(rng) if seed % 3 else _boat_config(rng)
    mode = int(rng.integers(0, 2))
    try:
        spec = make_spec(name, autoreset_mode=mode, **kw)
    except (NotImplementedError, ValueError) as e:             # a combination the spec compiler refuses (it names the reason)
        pytest.skip("%s %r: %s" % (name, kw, e))
    N, T = 257 + 32 * int(rng.integers(0, 4)), 90
    hi = 9 if rng.random() < 0.25 else 4                       # QUIT (9) and the unused action numbers now and then
    env = VectorEnv(spec, N, env_index_base=int(rng.integers(0, 1 << 30)), autoreset_mode=mode, want_value_board=True)
    orc = oracle_lib.Oracle(spec, N)
    orc.reset()
    np.testing.assert_array_equal(_np(env.board), orc.board)
    np.testing.assert_array_equal(_np(env.cube), orc.cube)
    ended = 0
    for t in range(T):
        a = rng.integers(0, hi + 1, size=N).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        ctx = "seed=%d %s %r mode=%d t=%d" % (seed, name, kw, mode, t)
        np.testing.assert_array_equal(_np(env.board), orc.board, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.cube), orc.cube, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.value_board), orc.value_board, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.step_type), orc.step_type, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.reason), orc.reason, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.terminated), orc.terminated, err_msg=ctx)
        np.testing.assert_allclose(_np(env.reward), orc.reward, rtol=1e-6, atol=0, err_msg=ctx)
        ended += int(orc.terminated.sum())
        if t % 10 == 9 or t == T - 1:
            ex, ox = env.observe(), orc.observe()
            np.testing.assert_array_equal(_np(ex["frame"]), ox["frame"], err_msg=ctx)
            np.testing.assert_array_equal(_np(ex["pos"]), ox["pos"], err_msg=ctx)
            np.testing.assert_allclose(_np(ex["cumulative"]), ox["cumulative"], rtol=1e-6, atol=1e-5, err_msg=ctx)
            if spec.metric_names:
                np.testing.assert_allclose(_np(ex["metrics"]), ox["metrics"], rtol=1e-12, atol=1e-12, err_msg=ctx)
            if name == "island_navigation_ex":
                gd, gf = env.peek_fractions()
                od, of_ = orc.fractions()
                np.testing.assert_allclose(_np(gd), od, rtol=1e-12, atol=1e-13, err_msg=ctx)
                np.testing.assert_allclose(_np(gf), of_, rtol=1e-12, atol=1e-13, err_msg=ctx)
    assert ended > 0 or spec.config.max_iterations > T, "no episode ended: the case does not exercise the reset path"
    env.close()
    orc.close()


def _same(ctx, pairs):
    for name, got, want in pairs:
        np.testing.assert_array_equal(_np(got), want, err_msg=ctx + " " + name)


@pytest.mark.parametrize("seed", range(12))
def test_random_firemaker_configurations_match_the_oracle(seed, oracle_lib):
    """firemaker_ex_ma: 2 / 3 agents, direction modes 0-2, stop-button duration, fire probabilities, reward magnitudes, frame cut-off,
    shuffled or fixed order, both auto-reset semantics; Philox orders and fire draws shared with the oracle."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    rng = np.random.default_rng(2000 + seed)
    dmode = int(rng.integers(0, 3))
    kw = dict(amount_agents=int(rng.integers(2, 4)), observation_direction_mode=dmode, action_direction_mode=dmode,
              max_iterations=int(rng.integers(9, 150)), randomize_agent_actions_order=bool(rng.integers(0, 2)),
              STOP_BUTTON_PRESS_EFFECT_DURATION=int(rng.choice([1, 3, 6])), FIRE_CONTINUATION_PROBABILITY=float(rng.choice([0.95, 0.7])),
              FIRE_SPREAD_PROBABILITY_AT_DISTANCE_ONE=float(rng.choice([0.01, 0.02])),
              AGENT_MOVEMENT_REWARD="{'ENERGY': %g}" % rng.choice([-1, -0.5]), SUPERVISOR_EXTERNAL_FIRE_REWARD="{'EXTERNAL_FIRE': %g}" % rng.choice([-10, -2.5]))
    mode = int(rng.integers(0, 2))
    spec = make_spec("firemaker_ex_ma", autoreset_mode=mode, **kw)
    N = 130 + int(rng.integers(0, 40))
    base = int(rng.integers(0, 1 << 20))
    env = FiremakerVectorEnv(N, env_index_base=base, seed=seed, autoreset_mode=mode, spec=spec)
    orc = oracle_lib.FiremakerOracle(spec, N, env_index_base=base, seed=seed)
    orc.reset()
    for t in range(70):
        a = rng.integers(0, spec.action_range[1] + 1, size=(N, 3)).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        ctx = "seed=%d %r mode=%d t=%d" % (seed, kw, mode, t)
        _same(ctx, (("board", env.board, orc.board), ("cube", env.cube, orc.cube), ("crop_w", env.crop_workers, orc.crop_w),
                    ("crop_s", env.crop_supervisor, orc.crop_s), ("lcrop_w", env.lcrop_workers, orc.lcrop_w),
                    ("lcrop_s", env.lcrop_supervisor, orc.lcrop_s), ("reward_w", env.reward_workers, orc.reward_w),
                    ("reward_s", env.reward_supervisor, orc.reward_s), ("step_type", env.step_type, orc.step_type),
                    ("terminated", env.terminated, orc.terminated)))
        if t % 10 == 9:
            ex, ox = env.observe(), orc.observe()
            _same(ctx, [(k, ex[k], ox[k]) for k in ("metrics", "cumulative", "frame", "pos", "ext_fires", "directions")])
    env.close()
    orc.close()


@pytest.mark.parametrize("seed", range(12))
def test_random_island_ma_configurations_match_the_oracle(seed, oracle_lib):
    """island_navigation_ex_ma: levels, the four game flags, direction modes 0-2, shared or per-environment (device-shuffled) maps."""
    from ai_safety_gridworlds_b200 import IslandMaVectorEnv, make_spec
    rng = np.random.default_rng(3000 + seed)
    dmode = int(rng.integers(0, 3))
    kw = dict(level=int(rng.integers(0, 11)), max_iterations=int(rng.integers(10, 100)), sustainability_challenge=bool(rng.integers(0, 2)),
              thirst_hunger_death=bool(rng.integers(0, 2)), penalise_oversatiation=bool(rng.integers(0, 2)),
              use_satiation_proportional_reward=bool(rng.integers(0, 2)), observation_direction_mode=dmode, action_direction_mode=dmode,
              randomize_agent_actions_order=bool(rng.integers(0, 2)), map_randomization_frequency=int(rng.choice([0, 3])))
    mode = int(rng.integers(0, 2))
    try:
        spec = make_spec("island_navigation_ex_ma", autoreset_mode=mode, **kw)
    except (NotImplementedError, ValueError) as e:
        pytest.skip("%r: %s" % (kw, e))
    N = 200 + int(rng.integers(0, 70))
    base = int(rng.integers(0, 1 << 20))
    env = IslandMaVectorEnv(N, device="cuda:0", seed=seed, autoreset_mode=mode, spec=spec, env_index_base=base)
    orc = oracle_lib.IslandMaOracle(spec, N, env_index_base=base, seed=seed)
    omaps = None
    if kw["map_randomization_frequency"]:
        art = np.array([[ord(ch) for ch in row] for row in spec.art], np.uint8)
        omaps = np.repeat(art[None], N, axis=0).copy()
        orc.set_maps(omaps, 1)                             # GW_IMA_MAPS_SHUFFLE_EVERY_GAME, as the vector env configures the library
    orc.reset()
    if omaps is not None:
        np.testing.assert_array_equal(_np(env.maps), omaps)
    for t in range(70):
        a = rng.integers(0, spec.action_range[1] + 1, size=(N, 2)).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        ctx = "seed=%d %r mode=%d t=%d" % (seed, kw, mode, t)
        _same(ctx, (("board", env.board, orc.board), ("cube", env.cube, orc.cube), ("crop", env.crop, orc.crop), ("lcrop", env.lcrop, orc.lcrop),
                    ("step_type", env.step_type, orc.step_type), ("terminated", env.terminated, orc.terminated)))
        np.testing.assert_allclose(_np(env.reward), orc.reward, rtol=1e-6, atol=0, err_msg=ctx)
        if omaps is not None:
            np.testing.assert_array_equal(_np(env.maps), omaps, err_msg=ctx)
        if t % 10 == 9:
            gx, ox = env.observe(), orc.observe()
            _same(ctx, [(k, gx[k], ox[k]) for k in ("frame", "pos", "directions")])
    env.close()
    orc.close()


@pytest.mark.parametrize("seed", range(12))
def test_random_savanna_configurations_match_the_oracle(seed, oracle_lib):
    """aintelope_savanna: agents, predators, tile counts, direction modes, view radius, homeostasis flags, resized maps; layouts drawn on
    the device for every game."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
    rng = np.random.default_rng(4000 + seed)
    dmode = int(rng.integers(0, 3))
    kw = dict(amount_agents=int(rng.integers(1, 3)), amount_predators=int(rng.choice([0, 0, 2, 4])), amount_food_patches=int(rng.integers(1, 4)),
              amount_drink_holes=int(rng.integers(0, 3)), amount_small_food_patches=int(rng.integers(0, 2)), amount_small_drink_holes=int(rng.integers(0, 2)),
              amount_gold_deposits=int(rng.integers(0, 3)), amount_silver_deposits=int(rng.integers(0, 2)), amount_water_tiles=int(rng.choice([0, 2, 4])),
              observation_direction_mode=dmode, action_direction_mode=dmode, observation_radius=[int(rng.choice([3, 5, 10]))] * 4,
              penalise_oversatiation=bool(rng.integers(0, 2)), thirst_hunger_death=bool(rng.integers(0, 2)),
              use_satiation_proportional_reward=bool(rng.integers(0, 2)), max_iterations=int(rng.integers(15, 60)),
              remove_unused_tile_types_from_layers=bool(rng.integers(0, 2)))
    if rng.random() < 0.4:
        kw.update(map_width=int(rng.integers(7, 12)), map_height=int(rng.integers(7, 11)))
    mode = int(rng.integers(0, 2))
    try:
        spec = make_spec("aintelope_savanna", autoreset_mode=mode, **kw)
    except (NotImplementedError, ValueError, AssertionError) as e:
        pytest.skip("%r: %s" % (kw, e))
    N = 150 + int(rng.integers(0, 60))
    base = int(rng.integers(0, 1 << 20))
    env = SavannaVectorEnv(N, spec=spec, env_index_base=base, seed=seed, autoreset_mode=mode)
    orc = oracle_lib.SavannaOracle(spec.with_autoreset(mode), N, env_index_base=base, seed=seed)
    orc.set_maps(orc.maps, 1)
    orc.reset()
    for t in range(70):
        if t > 0:
            a = rng.integers(0, spec.action_range[1] + 1, size=(N, 2)).astype(np.int32)
            env.step(torch.from_numpy(a).to(env.device))
            orc.step(a)
        ctx = "seed=%d %r mode=%d t=%d" % (seed, kw, mode, t)
        np.testing.assert_array_equal(_np(env.maps).reshape(N, -1), orc.maps, err_msg=ctx)
        _same(ctx, (("board", env.board, orc.board), ("cube", env.cube, orc.cube), ("crop", env.crop, orc.crop), ("lcrop", env.lcrop, orc.lcrop),
                    ("step_type", env.step_type, orc.step_type), ("terminated", env.terminated, orc.terminated)))
        np.testing.assert_allclose(_np(env.reward), orc.reward, rtol=1e-6, atol=1e-6, err_msg=ctx)
        if t % 10 == 0:
            ex, ox = env.observe(all_slots=True), orc.observe()
            np.testing.assert_allclose(_np(ex["metrics"]), ox["metrics"], rtol=1e-12, atol=1e-12, err_msg=ctx)
            _same(ctx, [(k, ex[k], ox[k]) for k in ("frame", "pos", "directions")])
    env.close()
    orc.close()


@pytest.mark.parametrize("n", [1, 2, 3, 5, 17, 33, 63])
def test_tiny_and_odd_batches_match_the_oracle(n, oracle_lib):
    """Batches smaller than a warp pass and sizes that are no multiple of anything (the savanna kernel splits the batch into one
    contiguous range per warp in units of 4 environments, the lane-per-environment kernels run a ragged last chunk)."""
    from ai_safety_gridworlds_b200 import IslandMaVectorEnv, make_spec
    from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    rng = np.random.default_rng(n)
    spec = make_spec("aintelope_savanna", autoreset_mode=1, amount_agents=2, amount_predators=2, max_iterations=12)
    env = SavannaVectorEnv(n, spec=spec, env_index_base=5, seed=2, autoreset_mode=1)
    orc = oracle_lib.SavannaOracle(spec.with_autoreset(1), n, env_index_base=5, seed=2)
    orc.set_maps(orc.maps, 1)
    orc.reset()
    for t in range(30):
        a = rng.integers(0, 5, size=(n, 2)).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        ctx = "savanna n=%d t=%d" % (n, t)
        _same(ctx, (("maps", env.maps.reshape(n, -1), orc.maps), ("board", env.board, orc.board), ("cube", env.cube, orc.cube), ("crop", env.crop, orc.crop),
                    ("lcrop", env.lcrop, orc.lcrop), ("step_type", env.step_type, orc.step_type)))
        np.testing.assert_allclose(_np(env.reward), orc.reward, rtol=1e-6, atol=1e-6, err_msg=ctx)
    env.close(); orc.close()
    spec = make_spec("island_navigation_ex_ma", autoreset_mode=1, map_randomization_frequency=3, max_iterations=12)
    env = IslandMaVectorEnv(n, device="cuda:0", seed=4, autoreset_mode=1, spec=spec, env_index_base=9)
    orc = oracle_lib.IslandMaOracle(spec, n, env_index_base=9, seed=4)
    art = np.array([[ord(ch) for ch in row] for row in spec.art], np.uint8)
    omaps = np.repeat(art[None], n, axis=0).copy()
    orc.set_maps(omaps, 1)
    orc.reset()
    for t in range(30):
        a = rng.integers(0, 5, size=(n, 2)).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        ctx = "island_ma n=%d t=%d" % (n, t)
        _same(ctx, (("maps", env.maps, omaps), ("board", env.board, orc.board), ("cube", env.cube, orc.cube), ("crop", env.crop, orc.crop),
                    ("lcrop", env.lcrop, orc.lcrop), ("step_type", env.step_type, orc.step_type)))
    env.close(); orc.close()
    spec = make_spec("firemaker_ex_ma", autoreset_mode=1, amount_agents=3, max_iterations=30, observation_direction_mode=1, action_direction_mode=1)
    env = FiremakerVectorEnv(n, env_index_base=3, seed=6, autoreset_mode=1, spec=spec)
    orc = oracle_lib.FiremakerOracle(spec, n, env_index_base=3, seed=6)
    orc.reset()
    for t in range(30):
        a = rng.integers(0, 5, size=(n, 3)).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        ctx = "firemaker n=%d t=%d" % (n, t)
        _same(ctx, (("board", env.board, orc.board), ("cube", env.cube, orc.cube), ("crop_s", env.crop_supervisor, orc.crop_s),
                    ("lcrop_s", env.lcrop_supervisor, orc.lcrop_s), ("lcrop_w", env.lcrop_workers, orc.lcrop_w), ("step_type", env.step_type, orc.step_type)))
    env.close(); orc.close()
