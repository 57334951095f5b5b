"""island_navigation_ex_ma on the B200: the CUDA path replays every reference trace (shuffle order replayed) and matches
the C oracle on Philox-shuffled batches, bit-exactly for boards / cubes / rotated views / step types / integer metrics
and to 1e-6 relative for float rewards."""
import numpy as np
import pytest

from conftest import island_ma_golden_names, load_golden
from test_oracle_island_ma_golden import check_against_trace, ima_spec

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("name", island_ma_golden_names())
def test_cuda_replays_island_ma_reference_trace(name):
    from ai_safety_gridworlds_b200 import IslandMaVectorEnv
    d, meta = load_golden(name)
    spec = ima_spec(meta)
    spec.flags["map_randomization_frequency"] = 0            # the reference's layouts are replayed through set_maps below
    env = IslandMaVectorEnv(1, device="cuda:0", autoreset_mode=0, spec=spec)
    T = len(d["actions"])
    maps = None
    if meta["kwargs"].get("map_randomization_frequency") or name == "islandma_default_s1":   # (one fixed-map trace also goes through the per-environment-map kernel)
        maps = torch.from_numpy(np.ascontiguousarray(d["maps"][:1])).to(env.device)
        env.set_maps(maps, 0)
        env.reset()
    for t in range(T + 1):
        if maps is not None:
            maps.copy_(torch.from_numpy(np.ascontiguousarray(d["maps"][t:t + 1])))
            if t == 0:
                env.reset()
        if t > 0:
            a = torch.from_numpy(np.maximum(d["actions"][t - 1], 0)[None].astype(np.int32)).to(env.device)
            o = torch.from_numpy(d["order"][t - 1][None].astype(np.int32)).to(env.device)
            env.step(a, o)
        ox = {k: v[0].cpu().numpy() for k, v in env.observe().items()}
        full = np.zeros(16)
        full[[spec.config.metric_slots[i] for i in range(spec.config.n_metrics)]] = ox["metrics"]
        ox["metrics"] = full
        view = dict(board=env.board[0].cpu().numpy(), cube=env.cube[0].cpu().numpy(), crop=env.crop[0].cpu().numpy(),
                    lcrop=env.lcrop[0].cpu().numpy(), reward=env.reward[0].cpu().numpy(), step_type=env.step_type[0].cpu().numpy(),
                    terminated=env.terminated[0].cpu().numpy())
        check_against_trace(view, ox, spec, d, t, "%s t=%d" % (name, t))
    env.close()


@pytest.mark.parametrize("kwargs,mode,n", [
    ({}, 1, 4096 + 13),
    (dict(sustainability_challenge=True, thirst_hunger_death=True, penalise_oversatiation=True), 1, 2048),
    (dict(penalise_oversatiation=True, use_satiation_proportional_reward=True, level=10), 0, 1000),
    (dict(level=4, max_iterations=30, observation_direction_mode=0), 1, 777),
    (dict(level=0, action_direction_mode=0, observation_direction_mode=0, randomize_agent_actions_order=False), 1, 96),
    (dict(level=6, action_direction_mode=2, observation_direction_mode=2, penalise_oversatiation=True), 1, 1500 + 3),   # TURN_* actions 5..8
])
def test_island_ma_matches_oracle_with_philox_order(kwargs, mode, n, oracle_lib):
    from ai_safety_gridworlds_b200 import IslandMaVectorEnv, make_spec
    spec = make_spec("island_navigation_ex_ma", autoreset_mode=mode, **kwargs)
    env = IslandMaVectorEnv(n, device="cuda:0", seed=5, autoreset_mode=mode, spec=spec, env_index_base=1000)
    orc = oracle_lib.IslandMaOracle(spec, n, env_index_base=1000, seed=5)
    orc.reset()
    rng = np.random.default_rng(1)
    ended = 0
    R = spec.n_rewards
    run = np.zeros((n, 2, R))                     # episode return so far, rebuilt from the oracle's reward rows
    want = dict(env_steps=0, episodes=0, length_sum=0, agent_finishes=0, ret=np.zeros((2, R)))
    for t in range(80):
        a = rng.integers(0, 9 if kwargs.get("action_direction_mode") == 2 else 5, size=(n, 2)).astype(np.int32)
        # a call that only restarts a finished game (GW_AUTORESET_NEXT_STEP) plays no step
        want["env_steps"] += int((~(orc.step_type >= 2).all(axis=1)).sum()) if (t and mode == 0) else n
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        run += orc.reward
        over = orc.terminated.all(axis=1)
        want["episodes"] += int(over.sum())
        want["agent_finishes"] += int((orc.step_type == 2).sum())
        want["ret"] += run[over].sum(axis=0)
        run[over] = 0
        if mode == 0:
            want["length_sum"] += int(orc.observe()["frame"][over].sum())
        assert np.array_equal(env.board.cpu().numpy(), orc.board), t
        assert np.array_equal(env.cube.cpu().numpy(), orc.cube), t
        assert np.array_equal(env.crop.cpu().numpy(), orc.crop), t
        assert np.array_equal(env.lcrop.cpu().numpy(), orc.lcrop), t
        assert np.array_equal(env.step_type.cpu().numpy(), orc.step_type), t
        assert np.array_equal(env.terminated.cpu().numpy(), orc.terminated), t
        np.testing.assert_allclose(env.reward.cpu().numpy(), orc.reward, rtol=1e-6, atol=0, err_msg=str(t))
        ended += int(orc.terminated.all(axis=1).sum())
        if t % 20 == 19:
            gx, ox = env.observe(), orc.observe()
            slots = [spec.config.metric_slots[i] for i in range(spec.config.n_metrics)]
            np.testing.assert_allclose(gx["metrics"].cpu().numpy(), ox["metrics"][:, slots], rtol=1e-9)
            np.testing.assert_allclose(gx["cumulative"].cpu().numpy(), ox["cumulative"], rtol=1e-5, atol=1e-4)
            assert np.array_equal(gx["frame"].cpu().numpy(), ox["frame"]) and np.array_equal(gx["pos"].cpu().numpy(), ox["pos"])
            assert np.array_equal(gx["directions"].cpu().numpy(), ox["directions"])
    assert ended > 0
    st = env.stats()                                             # device-side rollout statistics: exact integer sums
    assert st["env_steps"] == want["env_steps"] and st["episodes"] == want["episodes"] == ended
    assert st["agent_finishes"] == want["agent_finishes"]
    if mode == 0:
        assert st["length_sum"] == want["length_sum"]
    got = np.array([[st["return_sum"][a][k] for k in spec.reward_keys] for a in ("1", "2")])
    np.testing.assert_allclose(got, want["ret"], rtol=1e-6, atol=1e-3)
    env.clear_stats()
    assert env.stats()["episodes"] == 0
    env.close(); orc.close()


def test_island_ma_full_size_sharding_invariance():
    """BASELINE-size batch (1,048,576 environments): two half-size shards with env_index_base offsets reproduce the
    unsharded batch exactly -- the Philox shuffle is keyed by the global environment index -- and their raw statistics
    vectors add up to the unsharded one bit for bit (what the end-of-rollout NCCL all-reduce relies on)."""
    from ai_safety_gridworlds_b200 import IslandMaVectorEnv
    N = 1 << 20
    dev = torch.device("cuda", 0)
    full = IslandMaVectorEnv(N, device=dev, seed=9, autoreset_mode=1)
    lo = IslandMaVectorEnv(N // 2, device=dev, seed=9, autoreset_mode=1, env_index_base=0)
    hi = IslandMaVectorEnv(N // 2, device=dev, seed=9, autoreset_mode=1, env_index_base=N // 2)
    g = torch.Generator(device=dev); g.manual_seed(4)
    for t in range(40):
        a = torch.randint(0, 5, (N, 2), dtype=torch.int32, device=dev, generator=g)
        full.step(a); lo.step(a[: N // 2].contiguous()); hi.step(a[N // 2:].contiguous())
    assert torch.equal(full.board[: N // 2], lo.board) and torch.equal(full.board[N // 2:], hi.board)
    assert torch.equal(full.lcrop[N // 2:], hi.lcrop) and torch.equal(full.reward[: N // 2], lo.reward)
    raw = lo.stats_raw_device().clone() + hi.stats_raw_device()
    assert torch.equal(raw, full.stats_raw_device())
    st = full.stats()
    assert st["env_steps"] == 40 * N and st["episodes"] > N
    # every environment shows exactly one '1' and one '2'; the cube's agent layers agree with the board
    assert int((full.board == ord("1")).sum()) == N and int((full.board == ord("2")).sum()) == N
    l1 = full.spec.layer_order.index("1")
    assert torch.equal(full.cube[:, l1].bool(), full.board == ord("1"))
    for e in (full, lo, hi):
        e.close()


def test_island_ma_masked_reset_and_single_agent_frames(oracle_lib):
    from ai_safety_gridworlds_b200 import IslandMaVectorEnv, make_spec
    n = 333
    spec = make_spec("island_navigation_ex_ma", autoreset_mode=0)
    env = IslandMaVectorEnv(n, device="cuda:0", seed=2, autoreset_mode=0, spec=spec)
    orc = oracle_lib.IslandMaOracle(spec, n, seed=2)
    orc.reset()
    rng = np.random.default_rng(3)
    for t in range(40):
        a = rng.integers(0, 5, size=(n, 2)).astype(np.int32)
        order = np.tile(np.array([[t % 2, -1]], np.int32), (n, 1))          # AEC-style: one agent's frame per call
        env.step(torch.from_numpy(a).to(env.device), torch.from_numpy(order).to(env.device))
        orc.step(a, order)
        if t % 7 == 6:
            mask = (rng.random(n) < 0.3).astype(np.uint8)
            env.reset(torch.from_numpy(mask).to(env.device))
            orc.reset(mask)
        assert np.array_equal(env.board.cpu().numpy(), orc.board), t
        assert np.array_equal(env.lcrop.cpu().numpy(), orc.lcrop), t
        assert np.array_equal(env.step_type.cpu().numpy(), orc.step_type), t
        np.testing.assert_allclose(env.reward.cpu().numpy(), orc.reward, rtol=1e-6, err_msg=str(t))
    env.close(); orc.close()


@pytest.mark.parametrize("mode,freq,resize", [(1, 3, None), (0, 3, None), (1, 1, None), (1, 3, (6, 9)), (0, 1, (8, 8))])
def test_island_ma_map_randomisation_matches_oracle(mode, freq, resize, oracle_lib):
    """Per-environment layouts drawn on the device (Philox Fisher-Yates of the interior at every new game / every explicit
    reset): CUDA and oracle produce the same layouts and the same games; every layout is a permutation of the level's
    interior with the edges preserved."""
    from ai_safety_gridworlds_b200 import IslandMaVectorEnv, make_spec, _abi
    n = 1500 + 7
    kw = {} if resize is None else dict(map_height=resize[0], map_width=resize[1])      # map resizing: the agents alone inside a water border
    spec = make_spec("island_navigation_ex_ma", autoreset_mode=mode, map_randomization_frequency=freq, max_iterations=40, **kw)
    if resize is not None:
        assert (spec.height, spec.width) == resize and "".join(spec.art).count("W") == 2 * (resize[0] + resize[1]) - 4
    env = IslandMaVectorEnv(n, device="cuda:0", seed=8, autoreset_mode=mode, spec=spec, env_index_base=64)
    orc = oracle_lib.IslandMaOracle(spec, n, env_index_base=64, seed=8)
    art = np.array([[ord(ch) for ch in row] for row in spec.art], np.uint8)
    omaps = np.repeat(art[None], n, axis=0).copy()
    orc.set_maps(omaps, _abi.GW_IMA_MAPS_SHUFFLE_EVERY_GAME if freq == 3 else _abi.GW_IMA_MAPS_SHUFFLE_ON_RESET)
    orc.reset()
    assert np.array_equal(env.maps.cpu().numpy(), omaps)
    first = omaps.copy()
    # the environments got different layouts (a resized map only holds the two agents: 28 * 27 / 36 * 35 placements)
    assert len({m.tobytes() for m in first}) > (n // 2 if resize is None else n // 4)
    assert np.array_equal(np.sort(first[:, 1:-1, 1:-1].reshape(n, -1), axis=1), np.sort(np.tile(art[1:-1, 1:-1].reshape(1, -1), (n, 1)), axis=1))
    assert (first[:, 0] == art[0]).all() and (first[:, -1] == art[-1]).all() and (first[:, :, 0] == art[:, 0]).all() and (first[:, :, -1] == art[:, -1]).all()
    rng = np.random.default_rng(5)
    for t in range(60):
        a = rng.integers(0, 5, size=(n, 2)).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        assert np.array_equal(env.maps.cpu().numpy(), omaps), t
        assert np.array_equal(env.board.cpu().numpy(), orc.board), t
        assert np.array_equal(env.cube.cpu().numpy(), orc.cube), t
        assert np.array_equal(env.crop.cpu().numpy(), orc.crop), t
        assert np.array_equal(env.lcrop.cpu().numpy(), orc.lcrop), t
        assert np.array_equal(env.step_type.cpu().numpy(), orc.step_type), t
        np.testing.assert_allclose(env.reward.cpu().numpy(), orc.reward, rtol=1e-6, atol=0, err_msg=str(t))
    changed = (omaps != first).any(axis=(1, 2)).mean()
    assert (changed > 0.9) if freq == 3 else (changed == 0.0)              # frequency 3 redraws at every game, 1 keeps the layout
    if freq == 1:
        env.reset(); orc.reset()
        assert np.array_equal(env.maps.cpu().numpy(), omaps) and (omaps != first).any(axis=(1, 2)).mean() > 0.9
    env.close(); orc.close()
