"""GPU parity for firemaker_ex_ma (BASELINE config 4): reference traces (shuffle order and FireDrape
draws replayed) and the CPU oracle on Philox-driven batches.  Everything is byte / integer valued."""
import numpy as np
import pytest

from conftest import firemaker_golden_names, load_golden
from test_oracle_firemaker_golden import check_against_trace, fm_spec, replay_inputs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.cpu().numpy()


def _view(env, k):
    return dict(board=_np(env.board[k]), cube=_np(env.cube[k]), crop_w=_np(env.crop_workers[k]), crop_s=_np(env.crop_supervisor[k]),
                lcrop_w=_np(env.lcrop_workers[k]), lcrop_s=_np(env.lcrop_supervisor[k]), reward_w=_np(env.reward_workers[k]),
                reward_s=_np(env.reward_supervisor[k]), step_type=_np(env.step_type[k]), terminated=_np(env.terminated[k]))


@pytest.mark.parametrize("name", firemaker_golden_names())
def test_cuda_replays_firemaker_reference_trace(name):
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    d, meta = load_golden(name)
    spec = fm_spec(meta)
    env = FiremakerVectorEnv(2, autoreset_mode=0, spec=spec)
    T = len(d["actions"])
    for t in range(T + 1):
        if t == 0:
            env.reset()
        else:
            a, o, dr = replay_inputs(d, t)
            dev = env.device
            env.step(torch.from_numpy(np.repeat(a, 2, 0)).to(dev), torch.from_numpy(np.repeat(o, 2, 0)).to(dev),
                     torch.from_numpy(np.repeat(dr, 2, 0)).to(dev))
        ex = env.observe()
        for k in (0, 1):
            ox = {key: _np(v[k]) for key, v in ex.items()}
            check_against_trace(_view(env, k), ox, d, meta, t, "%s t=%d env=%d" % (name, t, k))
    env.close()


@pytest.mark.parametrize("agents", [3, 2])
@pytest.mark.parametrize("mode", [0, 1])
def test_firemaker_matches_oracle_with_philox_draws(mode, agents, oracle_lib):
    """100 environments, 90 parallel steps (frames beyond the 200-frame cut-off, so episodes end and
    restart), shuffle orders and fire draws from the shared Philox streams."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    spec = make_spec("firemaker_ex_ma", autoreset_mode=mode, max_iterations=200, amount_agents=agents)
    N = 100
    env = FiremakerVectorEnv(N, env_index_base=31, seed=5, autoreset_mode=mode, spec=spec)
    orc = oracle_lib.FiremakerOracle(spec, N, env_index_base=31, seed=5)
    orc.reset()
    rng = np.random.default_rng(0)
    for t in range(90):
        a = rng.integers(0, 5, size=(N, 3)).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        ctx = "mode=%d t=%d" % (mode, t)
        np.testing.assert_array_equal(_np(env.board), orc.board, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.cube), orc.cube, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.crop_workers), orc.crop_w, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.crop_supervisor), orc.crop_s, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.lcrop_workers), orc.lcrop_w, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.lcrop_supervisor), orc.lcrop_s, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.reward_workers), orc.reward_w, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.reward_supervisor), orc.reward_s, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.step_type), orc.step_type, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.terminated), orc.terminated, err_msg=ctx)
        ex, ox = env.observe(), orc.observe()
        for key in ("metrics", "cumulative", "frame", "pos", "ext_fires"):
            np.testing.assert_array_equal(_np(ex[key]), ox[key], err_msg=ctx + " " + key)
    assert int(_np(env.board == ord("F")).sum()) > 0
    env.close()
    orc.close()


@pytest.mark.parametrize("dmode,agents,autoreset", [(1, 3, 1), (2, 3, 0), (1, 2, 0), (2, 2, 1)])
def test_firemaker_direction_modes_match_oracle_with_philox_draws(dmode, agents, autoreset, oracle_lib):
    """Direction modes 1 (relative to the last move) and 2 (TURN_* actions 5..8): gw_fm_kernel<true> against the oracle on 150
    environments -- the three rotated views and their layers, the sprites' directions, everything else as in mode 0."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    spec = make_spec("firemaker_ex_ma", autoreset_mode=autoreset, max_iterations=150, amount_agents=agents,
                     observation_direction_mode=dmode, action_direction_mode=dmode)
    assert spec.action_range == (0, 8 if dmode == 2 else 4)
    N = 150
    env = FiremakerVectorEnv(N, env_index_base=7, seed=11, autoreset_mode=autoreset, spec=spec)
    orc = oracle_lib.FiremakerOracle(spec, N, env_index_base=7, seed=11)
    orc.reset()
    rng = np.random.default_rng(3)
    turned = set()
    for t in range(80):
        a = rng.integers(0, spec.action_range[1] + 1, size=(N, 3)).astype(np.int32)
        env.step(torch.from_numpy(a).to(env.device))
        orc.step(a)
        ctx = "dmode=%d agents=%d t=%d" % (dmode, agents, t)
        for name, got, want in (("board", env.board, orc.board), ("cube", env.cube, orc.cube), ("crop_w", env.crop_workers, orc.crop_w),
                                ("crop_s", env.crop_supervisor, orc.crop_s), ("lcrop_w", env.lcrop_workers, orc.lcrop_w),
                                ("lcrop_s", env.lcrop_supervisor, orc.lcrop_s), ("reward_w", env.reward_workers, orc.reward_w),
                                ("reward_s", env.reward_supervisor, orc.reward_s), ("step_type", env.step_type, orc.step_type),
                                ("terminated", env.terminated, orc.terminated)):
            np.testing.assert_array_equal(_np(got), want, err_msg=ctx + " " + name)
        ex, ox = env.observe(), orc.observe()
        for key in ("metrics", "cumulative", "frame", "pos", "ext_fires", "directions"):
            np.testing.assert_array_equal(_np(ex[key]), ox[key], err_msg=ctx + " " + key)
        turned |= set(np.unique(ox["directions"][:, 2, 1]).tolist())
    assert turned == {0, 1, 2, 3}                     # the supervisor's 33 x 33 view was emitted in all four rotations
    env.close()
    orc.close()


def test_firemaker_rejects_unsupported_configurations():
    from ai_safety_gridworlds_b200 import make_spec
    for kw in ({"amount_agents": 4}, {"amount_agents": 1}, {"observation_direction_mode": 1}, {"observation_direction_mode": 3, "action_direction_mode": 3},
               {"agent_observation_radius": [1, 1, 1, 1]},
               {"FIRE_SPREAD_EXCLUSIVE_MAX_DISTANCE": 4.0}, {"AGENT_MOVEMENT_REWARD": "{'OTHER': -1}"}):
        with pytest.raises(NotImplementedError):
            make_spec("firemaker_ex_ma", **kw)


def test_firemaker_full_size_sharding_invariance_and_statistics():
    """BASELINE config 4 size (262,144 environments): two half-size shards with env_index_base offsets reproduce the
    unsharded batch exactly (Philox order and fire draws are keyed by the global environment index) and their raw
    statistics add up bit for bit; the statistics agree with what the step outputs say."""
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    N = 1 << 18
    dev = torch.device("cuda", 0)
    kw = dict(device=dev, seed=21, autoreset_mode=1, max_iterations=45, want_cube=False, want_layer_crops=False)
    full = FiremakerVectorEnv(N, **kw)
    lo = FiremakerVectorEnv(N // 2, env_index_base=0, **kw)
    hi = FiremakerVectorEnv(N // 2, env_index_base=N // 2, **kw)
    g = torch.Generator(device=dev); g.manual_seed(6)
    episodes = 0
    ret = torch.zeros((7,), dtype=torch.float64, device=dev)
    run = torch.zeros((N, 7), dtype=torch.float64, device=dev)
    for t in range(32):
        a = torch.randint(0, 5, (N, 3), dtype=torch.int32, device=dev, generator=g)
        full.step(a); lo.step(a[: N // 2].contiguous()); hi.step(a[N // 2:].contiguous())
        run += torch.cat([full.reward_workers.reshape(N, 4), full.reward_supervisor], dim=1).double()
        over = full.terminated.bool().all(dim=1)
        episodes += int(over.sum())
        ret += run[over].sum(dim=0)
        run[over] = 0
    assert torch.equal(full.board[: N // 2], lo.board) and torch.equal(full.board[N // 2:], hi.board)
    assert torch.equal(full.crop_supervisor[N // 2:], hi.crop_supervisor)
    raw = lo.stats_raw_device().clone() + hi.stats_raw_device()
    assert torch.equal(raw, full.stats_raw_device())
    st = full.stats()
    assert st["env_steps"] == 32 * N and st["episodes"] == episodes == 2 * N     # 45 frames = 15 parallel steps per game
    assert st["length_sum"] == 45 * episodes and st["agent_finishes"] == 3 * episodes
    got = [st["return_sum"][a][k] for a in ("1", "2", "S") for k in full.spec.reward_keys[a]]
    np.testing.assert_allclose(got, ret.cpu().numpy(), rtol=1e-12)
    assert int((full.board == ord("S")).sum()) == N
    for e in (full, lo, hi):
        e.close()
