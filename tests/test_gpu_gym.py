"""The Gym-signature wrapper over the CUDA backend, replaying traces that were recorded through the
reference's own GridworldGymEnv (oracle/record.py): same call sequence, same return tuple."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

NAMES = [n for n in golden_names() if n in ("island_ex_default_s0", "island_ex_fractional_s19", "island_ex_level5_s8",
                                             "boat_ex_level3_s0", "boat_ex_level2_noops_off_s9")]


@pytest.mark.parametrize("name", NAMES)
def test_single_env_drop_in_matches_reference_tuple(name):
    from ai_safety_gridworlds_b200 import GridworldGymEnv
    d, meta = load_golden(name)
    env = GridworldGymEnv(meta["env"], seed=meta["seed"], **meta["kwargs"])
    assert env.action_space.min_action == meta["action_min"] and env.action_space.max_action == meta["action_max"]
    obs, info = env.reset()
    H, W = d["board"].shape[1:]
    assert obs.shape == (1, H, W) and obs.dtype == np.float32 and env.observation_space.shape == (1, H, W)
    assert info["info_observation_layers_order"] == meta["layer_order"]
    for t in range(len(d["actions"]) + 1):
        if t > 0:
            obs, reward, terminated, truncated, info = env.step(int(d["actions"][t - 1]))
            assert truncated is False and isinstance(terminated, bool)
            assert terminated == (d["step_type"][t] == 2)
            assert reward.dtype == np.float64 and reward.shape == d["reward"][t].shape
            np.testing.assert_allclose(reward, d["reward"][t], rtol=1e-6, atol=0)
        np.testing.assert_array_equal(obs[0], d["obs"][t])
        np.testing.assert_array_equal(info["ascii_codes"], d["board"][t])
        np.testing.assert_array_equal(info["info_observation_layers_cube"], d["cube"][t].astype(bool))
        np.testing.assert_allclose(info["cumulative_reward"], d["cumulative"][t], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(info["average_reward"], d["average"][t], rtol=1e-6, atol=1e-6)
        reason = info["extra_observations"]["termination_reason"]
        assert (-1 if reason is None else reason) == d["reason"][t]
        disc = info["discount"]
        assert (np.isnan(d["discount"][t]) and disc is None) or disc == d["discount"][t]
        if meta["metric_names"]:
            assert list(info["metrics_dict"].keys()) == meta["metric_names"]
            np.testing.assert_allclose(list(info["metrics_dict"].values()), d["metrics"][t], rtol=1e-12, atol=1e-12)
        # coordinates: the agent layer holds exactly the recorded agent position
        assert info["info_observation_coordinates"]["A"] == [tuple(int(v) for v in d["pos"][t])]
    env.close()


def test_batched_form_and_transitions():
    from ai_safety_gridworlds_b200 import GridworldGymEnv
    N = 4096
    env = GridworldGymEnv("island_navigation_ex", num_envs=N, use_transitions=True, layers_order_in_cube=["A", "W", "#"])
    obs, info = env.reset()
    assert obs.shape == (N, 2, 6, 8) and obs.is_cuda and torch.equal(obs[:, 0], obs[:, 1])
    assert info["info_observation_layers_cube"].shape == (N, 3, 6, 8) and info["info_observation_layers_cube"].dtype == torch.bool
    prev = obs[:, 1].clone()
    total_done = 0
    for t in range(20):
        a = env.vector_env.random_actions(3, t)
        obs, reward, terminated, truncated, info = env.step(a)
        assert torch.equal(obs[:, 0], prev)                    # use_transitions: [board(t-1), board(t)]
        prev = obs[:, 1].clone()
        assert reward.shape == (N, 10) and reward.dtype == torch.float64
        assert terminated.dtype == torch.bool and not bool(truncated.any())
        total_done += int(terminated.sum())
        # an environment that terminated already shows the first frame of its next episode
        started = info["frame"][terminated]
        assert bool((started == 0).all())
    assert total_done > 0
    with pytest.raises(NotImplementedError):
        GridworldGymEnv("no_such_environment")
    env.close()
