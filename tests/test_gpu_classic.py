"""GPU parity for the original DeepMind suite (BASELINE config 5): golden traces recorded from the
reference, the CPU oracle on a seeded mixed batch, and the rollout statistics.  Everything here is
integer or byte valued: bit-exact -- except the tomato games' float rewards (multiples of REWARD_FACTOR = 0.02):
the per-step reward equals float32(reference reward) exactly, the cumulative sums compare to 1e-6 relative."""
import numpy as np
import pytest

from conftest import classic_golden_names, load_golden, spec_for

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("name", classic_golden_names())
def test_cuda_replays_classic_reference_trace(name):
    from ai_safety_gridworlds_b200.classic_env import ClassicVectorEnv, crop_board
    d, meta = load_golden(name)
    spec = spec_for(meta)
    env = ClassicVectorEnv([spec], [3], autoreset_mode=0)
    coins = torch.full((3,), 255, dtype=torch.uint8, device=env.device)
    env.set_coin_override(coins)
    tomato = meta["env"].startswith("tomato")
    dried = torch.full((3,), 0xFFFF, dtype=torch.uint16, device=env.device)
    if tomato:
        env.set_dried_override(dried)
    H, W = spec.height, spec.width
    T = len(d["actions"])
    perf = float("nan")
    for t in range(T + 1):
        if (t == 0 or d["step_type"][t] == 0) and d["coin"][t] >= 0:
            coins.fill_(int(d["coin"][t]))               # replay the reference's MT19937 draw for this episode
        if tomato:
            dried.fill_(int(d["dried"][t]))              # and its per-frame draws (tomato games)
        if t == 0:
            env.reset()
        else:
            env.step(torch.full((3,), int(d["actions"][t - 1]), dtype=torch.int32, device=env.device))
        ctx = "%s t=%d" % (name, t)
        for k in range(3):
            np.testing.assert_array_equal(_np(crop_board(env.board[k], spec)), d["board"][t], err_msg=ctx)
            np.testing.assert_array_equal(_np(crop_board(env.value_board[k], spec)), d["obs"][t], err_msg=ctx)
        if W > 8:
            assert not bool(env.board.flatten(1)[:, H * W:].any()), ctx
        else:
            assert not bool(env.board[:, H:, :].any()) and not bool(env.board[:, :, W:].any()), ctx
        assert int(env.step_type[0]) == d["step_type"][t], ctx
        assert int(env.reason[1]) == d["reason"][t], ctx
        assert float(env.reward[2, 0]) == (np.float32(d["reward"][t]) if tomato else d["reward"][t]), ctx
        assert int(env.actual[0]) == d["actual"][t], ctx
        ex = env.observe()
        if tomato:
            assert float(ex["cumulative"][0, 0]) == pytest.approx(d["ret"][t], rel=1e-6, abs=1e-12), ctx
            assert float(ex["cumulative"][1, 1]) == pytest.approx(d["hidden"][t], rel=1e-6, abs=1e-12), ctx
        else:
            assert float(ex["cumulative"][0, 0]) == d["ret"][t], ctx
            assert float(ex["cumulative"][1, 1]) == d["hidden"][t], ctx
        np.testing.assert_array_equal(_np(ex["pos"][2]), d["pos"][t], err_msg=ctx)
        if d["coin"][t] >= 0:
            assert int(ex["coin"][0]) == d["coin"][t], ctx
        if "cube" in d:                                  # MO re-wrappings: obs['layers'] (un-occluded, gaps only where blank)
            lay = env.observe(layers=True)["layers"]
            for k in range(3):
                np.testing.assert_array_equal(_np(crop_board(lay[k], spec))[:len(spec.layer_order)], d["cube"][t], err_msg=ctx)
            assert not bool(lay[:, len(spec.layer_order):].any()), ctx
            assert float(ex["cumulative"][0, 0]) / (int(ex["frame"][0]) + 1) == d["average"][t], ctx
        if "policy" in d:                                # friend_foe: the PolicyEstimators, bit for bit ((0, 0) = never updated)
            pol = _np(env.policies())
            pol = np.where((pol == 0).all(axis=-1, keepdims=True), 0.5, pol)
            for k in range(3):
                np.testing.assert_array_equal(pol[k], d["policy"][t], err_msg=ctx)
    st = env.stats()
    ended = d["step_type"] == 2
    perf_col = d["ret"] if meta["env"] in ("whisky_gold", "distributional_shift", "friend_foe") or meta["env"].endswith("_ex") else d["hidden"]
    assert st["episodes"] == 3 * int(ended.sum())
    if tomato:
        assert st["return_sum"] == pytest.approx(3 * float(d["ret"][ended].sum()), rel=1e-6)
        assert st["hidden_sum"] == pytest.approx(3 * float(d["hidden"][ended].sum()), rel=1e-6)
        assert st["performance_sum"] == pytest.approx(3 * float(perf_col[ended].sum()), rel=1e-6)
    else:
        assert st["return_sum"] == 3 * float(d["ret"][ended].sum())
        assert st["hidden_sum"] == 3 * float(d["hidden"][ended].sum())
        assert st["performance_sum"] == 3 * float(perf_col[ended].sum())
    env.close()


MIX = [("safe_interruptibility", {}), ("side_effects_sokoban", {}), ("absent_supervisor", {}), ("conveyor_belt", {}),
       ("whisky_gold", {}), ("conveyor_belt", {"variant": "sushi_goal", "noops": True}), ("safe_interruptibility", {"level": 0, "max_iterations": 30})]


MIX2 = [("boat_race", {}), ("island_navigation", {}), ("whisky_gold", {}), ("boat_race", {"noops": True, "max_iterations": 40}),
        ("island_navigation", {"noops": False, "max_iterations": 25}), ("side_effects_sokoban", {"noops": True}), ("absent_supervisor", {})]


MIX3 = [("distributional_shift", {"is_testing": True}), ("rocks_diamonds", {}), ("tomato_watering", {}), ("tomato_crmdp", {}),
        ("rocks_diamonds", {"level": 1}), ("friend_foe", {}), ("friend_foe", {"bandit_type": "adversary", "extra_step": True})]
# the MO re-wrappings next to their originals (the same maps, different action decoding and reward routing)
MIX4 = [("conveyor_belt_ex", {}), ("conveyor_belt", {}), ("safe_interruptibility_ex", {}), ("safe_interruptibility", {}),
        ("conveyor_belt_ex", {"variant": "sushi_goal", "noops": True}), ("safe_interruptibility_ex", {"level": 2, "max_iterations": 30}),
        ("conveyor_belt_ex", {"variant": "sushi_goal2"})]
MIXES = {"config5": MIX, "with_boat_race_and_island_navigation": MIX2, "row3_games": MIX3, "mo_rewrappings": MIX4}


@pytest.mark.parametrize("mix", list(MIXES))
@pytest.mark.parametrize("mode", [0, 1])
def test_mixed_batch_matches_oracle(mode, mix, oracle_lib):
    """Seven types in one batch with type boundaries inside warps, 300 steps of Philox actions over
    NOOP..RIGHT plus QUIT, Philox per-episode draws: every tensor against the scalar oracle."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.classic_env import ClassicVectorEnv
    specs = [make_spec(n, autoreset_mode=mode, **kw) for n, kw in MIXES[mix]]
    unit = np.repeat([0.02 if n.startswith("tomato") else 1.0 for n, _ in MIXES[mix]], [157, 211, 96, 333, 64, 129, 77])
    counts = [157, 211, 96, 333, 64, 129, 77]
    env = ClassicVectorEnv(specs, counts, env_index_base=777, seed=99, autoreset_mode=mode)
    orc = oracle_lib.ClassicOracle(specs, counts, env_index_base=777, seed=99)
    orc.reset()
    N = sum(counts)
    np.testing.assert_array_equal(_np(env.board), orc.board)
    np.testing.assert_array_equal(_np(env.value_board), orc.value_board)
    ep = ret = hid = steps = 0
    for t in range(300):
        a = env.random_actions(5, t, lo=0, hi=8 if mix == "mo_rewrappings" and t % 5 == 4 else 4)    # 5-8: the MO turning actions
        if t % 37 == 36:
            a = torch.where(torch.arange(N, device=env.device) % 11 == 0, torch.full_like(a, 9), a)      # some QUITs
        a_np = _np(a)
        pre = orc.observe()
        was_last = orc.step_type == 2
        env.step(a)
        orc.step(a_np)
        ctx = "mode=%d t=%d" % (mode, t)
        np.testing.assert_array_equal(_np(env.board), orc.board, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.value_board), orc.value_board, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.reward), orc.reward, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.step_type), orc.step_type, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.reason), orc.reason, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.terminated), orc.terminated, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.actual), orc.actual, err_msg=ctx)
        ex, ox = env.observe(), orc.observe()
        # the episode sums are integers (tomato games: tomato counts, worth 0.02 each)
        np.testing.assert_array_equal(_np(ex["cumulative"][:, 0]), (ox["ret"] * unit).astype(np.float32), err_msg=ctx)
        np.testing.assert_array_equal(_np(ex["cumulative"][:, 1]), (ox["hidden"] * unit).astype(np.float32), err_msg=ctx)
        np.testing.assert_array_equal(_np(ex["frame"]), ox["frame"], err_msg=ctx)
        np.testing.assert_array_equal(_np(ex["pos"]), ox["pos"], err_msg=ctx)
        np.testing.assert_array_equal(_np(ex["coin"]), ox["coin"], err_msg=ctx)
        if mix == "mo_rewrappings" and t % 10 == 9:
            np.testing.assert_array_equal(_np(env.observe(layers=True)["layers"]), orc.layers(), err_msg=ctx)
        if mix == "row3_games" and t % 25 == 24:         # friend_foe's estimators (types 5 and 6 of the mix)
            lo = sum(counts[:5])
            pol = _np(env.policies())[lo:]
            pol = np.where((pol == 0).all(axis=-1, keepdims=True), 0.5, pol)
            np.testing.assert_array_equal(pol, orc.policies()[lo:], err_msg=ctx)
        ended = orc.terminated.astype(bool)
        steps += int((~was_last).sum()) if mode == 0 else N
        ep += int(ended.sum())
        if mode == 0:
            ret += float((ox["ret"][ended] * unit[ended]).sum()); hid += float((ox["hidden"][ended] * unit[ended]).sum())
        else:
            ret += float((pre["ret"][ended] * unit[ended] + orc.reward[ended, 0]).sum())
            hid += float((pre["hidden"][ended] * unit[ended] + orc.reward[ended, 1]).sum())
    st = env.stats()
    assert st["env_steps"] == steps and st["episodes"] == ep and ep > 100
    if mix == "row3_games":
        assert st["return_sum"] == pytest.approx(ret, rel=1e-6) and st["hidden_sum"] == pytest.approx(hid, rel=1e-6)
    else:
        assert st["return_sum"] == ret and st["hidden_sum"] == hid
    env.close()
    orc.close()


def test_classic_rejects_cube_and_wrong_constructor():
    import ctypes as C
    from ai_safety_gridworlds_b200 import _abi, make_spec
    from ai_safety_gridworlds_b200.classic_env import ClassicVectorEnv
    lib = _abi.load()
    spec = make_spec("whisky_gold")
    h = C.c_void_p()
    assert lib.gw_create(C.byref(spec.config), 64, 0, 0, C.byref(h)) == _abi.GW_ERR_INVALID
    assert b"gw_create_mixed" in lib.gw_last_error()
    env = ClassicVectorEnv(["whisky_gold"], [64])
    cube = torch.zeros(64 * 64, dtype=torch.uint8, device=env.device)
    obs = _abi.GwObs(C.c_void_p(env.board.data_ptr()), C.c_void_p(cube.data_ptr()), None)
    a = env.random_actions(0, 0)
    rc = lib.gw_step(env._h, C.c_void_p(a.data_ptr()), C.c_void_p(env.state.data_ptr()), C.byref(obs), None, None)
    assert rc == _abi.GW_ERR_INVALID and b"cube" in lib.gw_last_error()
    assert type(make_spec("side_effects_sokoban", level=1)).__name__ == "SokSpec"       # levels 1-3: the gw_sok_* path
    with pytest.raises(IndexError):
        make_spec("side_effects_sokoban", level=4)                                         # GAME_ART[level]
    with pytest.raises(ValueError):
        ClassicVectorEnv([make_spec("side_effects_sokoban", level=1)], [8])                # not a mixed-batch type
    env.close()
