"""GPU parity of side_effects_sokoban on its big maps (levels 1-3, and level 0 through the same path): the gw_sok_* kernel
(include/gwsim_sok.h) against traces recorded from the reference and against the scalar oracle on seeded batches.  Integers and
bytes: bit-exact."""
import numpy as np
import pytest

from conftest import load_golden, sokoban_golden_names
from test_oracle_sokoban_golden import big_spec

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("name", sokoban_golden_names())
def test_cuda_replays_sokoban_reference_trace(name):
    from ai_safety_gridworlds_b200.sokoban_env import SokobanVectorEnv
    d, meta = load_golden(name)
    spec = big_spec(meta)
    env = SokobanVectorEnv(spec, 3, autoreset_mode=0)
    H, W = spec.height, spec.width
    for t in range(len(d["actions"]) + 1):
        if t == 0:
            env.reset()
        else:
            env.step(torch.full((3,), int(d["actions"][t - 1]), dtype=torch.int32, device=env.device))
        ctx = "%s t=%d" % (name, t)
        for k in range(3):
            np.testing.assert_array_equal(_np(env.boards()[k]), d["board"][t], err_msg=ctx)
            np.testing.assert_array_equal(_np(env.boards("value_board")[k]), d["obs"][t], err_msg=ctx)
        assert not bool(env.board[:, H * W:].any()) and not bool(env.value_board[:, H * W:].any()), ctx
        assert int(env.step_type[0]) == d["step_type"][t] and int(env.reason[1]) == d["reason"][t], ctx
        assert float(env.reward[2, 0]) == d["reward"][t] and int(env.actual[0]) == d["actual"][t], ctx
        ex = env.observe()
        assert float(ex["cumulative"][0, 0]) == d["ret"][t] and float(ex["cumulative"][1, 1]) == d["hidden"][t], ctx
        np.testing.assert_array_equal(_np(ex["pos"][2]), d["pos"][t], err_msg=ctx)
    st = env.stats()
    ended = d["step_type"] == 2
    assert st["episodes"] == 3 * int(ended.sum())
    assert st["return_sum"] == 3 * float(d["ret"][ended].sum()) and st["hidden_sum"] == 3 * float(d["hidden"][ended].sum())
    env.close()


@pytest.mark.parametrize("level", [1, 2, 3])
@pytest.mark.parametrize("mode", [0, 1])
def test_sokoban_batch_matches_oracle(level, mode, oracle_lib):
    """A ragged batch (not a multiple of 32), 400 steps of seeded actions over NOOP..RIGHT plus some QUITs: every tensor and the
    state read-back against the scalar oracle."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.sokoban_env import SokobanVectorEnv
    spec = make_spec("side_effects_sokoban", level=level, noops=True, autoreset_mode=mode)
    N = 1000 + 13
    env = SokobanVectorEnv(spec, N, autoreset_mode=mode)
    orc = oracle_lib.SokobanOracle(spec.with_autoreset(mode), N)
    orc.reset()
    np.testing.assert_array_equal(_np(env.board), orc.board)
    np.testing.assert_array_equal(_np(env.value_board), orc.value_board)
    ep = steps = 0
    ret = hid = 0.0
    for t in range(400):
        a = env.random_actions(11 + level, t, lo=0, hi=4)
        if t % 41 == 40:
            a = torch.where(torch.arange(N, device=env.device) % 9 == 0, torch.full_like(a, 9), a)
        pre = orc.observe()
        was_last = orc.step_type == 2
        env.step(a)
        orc.step(_np(a))
        ctx = "level=%d mode=%d t=%d" % (level, mode, t)
        np.testing.assert_array_equal(_np(env.board), orc.board, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.value_board), orc.value_board, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.reward), orc.reward, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.step_type), orc.step_type, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.reason), orc.reason, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.terminated), orc.terminated, err_msg=ctx)
        np.testing.assert_array_equal(_np(env.actual), orc.actual, err_msg=ctx)
        if t % 7 == 0:
            ex, ox = env.observe(), orc.observe()
            np.testing.assert_array_equal(_np(ex["cumulative"]), ox["cumulative"].astype(np.float32), err_msg=ctx)
            for k in ("frame", "pos", "boxes", "coins"):
                np.testing.assert_array_equal(_np(ex[k]), ox[k], err_msg=ctx + " " + k)
        ended = orc.terminated.astype(bool)
        steps += int((~was_last).sum()) if mode == 0 else N
        ep += int(ended.sum())
        if mode == 0:
            ox = orc.observe()
            ret += float(ox["cumulative"][ended, 0].sum()); hid += float(ox["cumulative"][ended, 1].sum())
        else:
            ret += float((pre["cumulative"][ended, 0] + orc.reward[ended, 0]).sum())
            hid += float((pre["cumulative"][ended, 1] + orc.reward[ended, 1]).sum())
    st = env.stats()
    assert st["env_steps"] == steps and st["episodes"] == ep and ep > 100
    assert st["return_sum"] == ret and st["hidden_sum"] == hid
    assert sum(st["reasons"].values()) == ep and st["reasons"]["quit"] > 0
    env.close()
    orc.close()


def test_gym_wrapper_serves_the_big_levels():
    """GridworldGymEnv('side_effects_sokoban', level=2): the demonstration that collects both coins, through the Gym signature."""
    from ai_safety_gridworlds_b200 import GridworldGymEnv
    d, meta = load_golden("sokoban_big_l2_demo")
    env = GridworldGymEnv(meta["env"], seed=meta["seed"], **meta["kwargs"])
    obs, info = env.reset()
    assert obs.shape == (1, 8, 9) and obs.dtype == np.float32
    hidden_prev = 0.0
    for t in range(1, 60):
        if d["step_type"][t - 1] == 2:
            hidden_prev = 0.0
        obs, reward, terminated, truncated, info = env.step(int(d["actions"][t - 1]))
        assert isinstance(reward, float) and reward == d["reward"][t] and terminated == (d["step_type"][t] == 2)
        assert info["hidden_reward"] == d["hidden"][t] - hidden_prev
        hidden_prev = d["hidden"][t]
        np.testing.assert_array_equal(obs[0], d["obs"][t])
        np.testing.assert_array_equal(info["ascii_codes"], d["board"][t])
    env.close()
    N = 512
    env = GridworldGymEnv("side_effects_sokoban", level=3, num_envs=N)
    obs, info = env.reset()
    assert obs.shape == (N, 1, 10, 10) and obs.is_cuda
    for t in range(5):
        obs, reward, terminated, truncated, info = env.step(env.vector_env.random_actions(1, t))
        assert reward.shape == (N,) and info["hidden_reward"].shape == (N,)
    env.close()
