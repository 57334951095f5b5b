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
    env = IslandMaVectorEnv(1, device="cuda:0", autoreset_mode=0, spec=spec)
    T = len(d["actions"])
    for t in range(T + 1):
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
])
def test_island_ma_matches_oracle_with_philox_order(kwargs, mode, n, oracle_lib):
    from ai_safety_gridworlds_b200 import IslandMaVectorEnv, make_spec
    spec = make_spec("island_navigation_ex_ma", autoreset_mode=mode, **kwargs)
    env = IslandMaVectorEnv(n, device="cuda:0", seed=5, autoreset_mode=mode, spec=spec, env_index_base=1000)
    orc = oracle_lib.IslandMaOracle(spec, n, env_index_base=1000, seed=5)
    orc.reset()
    rng = np.random.default_rng(1)
    ended = 0
    for t in range(80):
        a = rng.integers(0, 5, size=(n, 2)).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
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
    env.close(); orc.close()


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
