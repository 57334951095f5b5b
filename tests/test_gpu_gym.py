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
    assert obs.shape == (N, 2, 6, 8) and obs.is_cuda and not bool(obs[:, 0].any())     # np.zeros_like(board) at reset (gridworld_gym_env.py:618-620)
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
        # info describes the episode that just ENDED (its last frame), the observation already is the next episode's first one
        assert bool((info["frame"][terminated] > 0).all())
    assert total_done > 0
    with pytest.raises(NotImplementedError):
        GridworldGymEnv("no_such_environment")
    env.close()


CLASSIC_NAMES = ["classic_boat_race_demo", "classic_island_navigation_demo", "classic_safe_interruptibility_l1_s17",
                 "classic_sokoban_l0_demo", "classic_absent_supervisor_s0", "classic_conveyor_vase_s0", "classic_whisky_gold_demo",
                 "classic_distributional_shift_test_s1", "classic_rocks_diamonds_demo", "classic_tomato_watering_demo",
                 "classic_friend_foe_random_s0"]


@pytest.mark.parametrize("name", CLASSIC_NAMES)
def test_classic_single_env_drop_in_matches_reference(name):
    """The original-suite games through the Gym signature (the reference serves them with the same wrapper,
    tests/gridworld_gym_env_test.py:63-110): obs float32 [1,H,W], scalar reward, per-step hidden reward in info."""
    from ai_safety_gridworlds_b200 import GridworldGymEnv
    d, meta = load_golden(name)
    env = GridworldGymEnv(meta["env"], seed=meta["seed"], **meta["kwargs"])
    assert env.action_space.min_action == meta["action_min"] and env.action_space.max_action == meta["action_max"]
    coins = d["coin"]

    def pin_coin(t):                               # the reference's per-episode MT19937 draw is replayed, not re-derived
        if coins[t] >= 0:
            env.set_coin_override(torch.tensor([int(coins[t])], dtype=torch.uint8, device=env.vector_env.device))
    tomato = meta["env"].startswith("tomato")
    tol = dict(rel=1e-6, abs=1e-9) if tomato else dict(rel=0, abs=0)   # 0.02 per tomato: float32 reward row, float64 reference

    def pin_draws(t):                              # the tomato games' per-frame draws of this call
        if tomato:
            env.set_dried_override(torch.tensor([int(d["dried"][t])], dtype=torch.uint16, device=env.vector_env.device))
    pin_coin(0)
    pin_draws(0)
    obs, info = env.reset()
    H, W = d["board"].shape[1:]
    assert obs.shape == (1, H, W) and obs.dtype == np.float32 and env.observation_space.shape == (1, H, W)
    hidden_prev = 0.0
    for t in range(len(d["actions"]) + 1):
        if t > 0:
            if d["step_type"][t - 1] == 2:
                pin_coin(t)                        # this call restarts the game: the new episode's draw
                hidden_prev = 0.0
            pin_draws(t)
            obs, reward, terminated, truncated, info = env.step(int(d["actions"][t - 1]))
            assert isinstance(reward, float) and reward == pytest.approx(d["reward"][t], **tol)
            assert terminated == (d["step_type"][t] == 2) and truncated is False
            assert info["hidden_reward"] == pytest.approx(d["hidden"][t] - hidden_prev, **tol)
            hidden_prev = d["hidden"][t]
            assert info["cumulative_reward"] == pytest.approx(d["ret"][t], **tol)
        np.testing.assert_array_equal(obs[0], d["obs"][t])
        np.testing.assert_array_equal(info["ascii_codes"], d["board"][t])
        reason = info["extra_observations"]["termination_reason"]
        assert (-1 if reason is None else reason) == d["reason"][t]
        aa = info["extra_observations"]["actual_actions"]
        assert (-1 if aa is None else aa) == d["actual"][t]
    env.close()


MO_REWRAP_NAMES = ["classic_conveyor_ex_vase_demo", "classic_conveyor_ex_sushi_goal_s3", "classic_safe_interruptibility_ex_l1_s0",
                   "classic_safe_interruptibility_ex_l2_noops_quit_s2"]


@pytest.mark.parametrize("name", MO_REWRAP_NAMES)
def test_mo_rewrapping_single_env_drop_in_matches_reference(name):
    """conveyor_belt_ex / safe_interruptibility_ex through the Gym signature: the SafetyEnvironmentMo tuple -- reward float64 [1],
    cumulative / average reward vectors, Gini / variance scalars (0: one dimension), un-occluded layers cube -- from traces
    recorded from the reference's environments."""
    from ai_safety_gridworlds_b200 import GridworldGymEnv
    d, meta = load_golden(name)
    env = GridworldGymEnv(meta["env"], seed=meta["seed"], **meta["kwargs"])
    assert env.action_space.min_action == meta["action_min"] and env.action_space.max_action == meta["action_max"]
    assert env.enabled_reward_dimension_keys == meta["reward_keys"] == ["REWARD"]
    coins = d["coin"]

    def pin_coin(t):
        if coins[t] >= 0:
            env.set_coin_override(torch.tensor([int(coins[t])], dtype=torch.uint8, device=env.vector_env.device))
    pin_coin(0)
    obs, info = env.reset()
    H, W = d["board"].shape[1:]
    assert obs.shape == (1, H, W) and obs.dtype == np.float32
    assert info["info_observation_layers_order"] == meta["layer_order"]
    for t in range(len(d["actions"]) + 1):
        if t > 0:
            if d["step_type"][t - 1] == 2:
                pin_coin(t)
            obs, reward, terminated, truncated, info = env.step(int(d["actions"][t - 1]))
            assert reward.dtype == np.float64 and reward.shape == (1,) and reward[0] == d["reward"][t]
            assert terminated == (d["step_type"][t] == 2) and truncated is False
        np.testing.assert_array_equal(obs[0], d["obs"][t])
        np.testing.assert_array_equal(info["ascii_codes"], d["board"][t])
        np.testing.assert_array_equal(info["info_observation_layers_cube"], d["cube"][t].astype(bool))
        assert info["cumulative_reward"].shape == (1,) and info["cumulative_reward"][0] == d["ret"][t]
        assert info["average_reward"][0] == d["average"][t]
        assert [float(info[k]) for k in ("gini_index", "cumulative_gini_index", "mo_variance", "cumulative_mo_variance",
                                         "average_mo_variance")] == list(d["scalars"][t])
        assert "hidden_reward" not in info
        reason = info["extra_observations"]["termination_reason"]
        assert (-1 if reason is None else reason) == d["reason"][t]
        aa = info["extra_observations"]["actual_actions"]
        assert (-1 if aa is None else aa) == d["actual"][t]
        assert info["info_observation_coordinates"]["A"] == [tuple(int(v) for v in d["pos"][t])]
    env.close()


def test_flatten_observations_and_classic_batched():
    from ai_safety_gridworlds_b200 import GridworldGymEnv
    env = GridworldGymEnv("island_navigation_ex", flatten_observations=True, use_transitions=True)
    obs, _ = env.reset()
    assert obs.shape == (2 * 6 * 8,) and obs.dtype == np.float32          # state.flatten(), gridworld_gym_env.py:537-538
    env.close()
    N = 2048
    env = GridworldGymEnv("boat_race", num_envs=N)
    obs, info = env.reset()
    assert obs.shape == (N, 1, 5, 5) and obs.is_cuda
    total = 0
    for t in range(120):
        a = env.vector_env.random_actions(1, t)
        obs, reward, terminated, truncated, info = env.step(a)
        assert reward.shape == (N,) and reward.dtype == torch.float64 and info["hidden_reward"].shape == (N,)
        total += int(terminated.sum())
    assert total == N                                                       # max_iterations = 100: every game ended once
    env.close()


def test_scalarise_sums_the_reward_dimensions():
    """SafetyEnvironmentMo(scalarise=True), safety_game_mo.py:1028-1064: scalar np.float64 reward = sum of the vector."""
    from ai_safety_gridworlds_b200 import GridworldGymEnv
    d, meta = load_golden("island_ex_default_s0")
    env = GridworldGymEnv(meta["env"], seed=meta["seed"], scalarise=True, **meta["kwargs"])
    env.reset()
    for t in range(1, 60):
        obs, reward, terminated, truncated, info = env.step(int(d["actions"][t - 1]))
        assert isinstance(reward, np.float64) and reward == d["reward"][t].sum()
        assert float(info["cumulative_reward"]) == pytest.approx(d["cumulative"][t].sum(), rel=1e-6)
    env.close()


def test_batched_info_describes_the_finished_episode(oracle_lib):
    """on a terminal step the batched wrapper's `info` (cumulative / average reward, metrics, frame, termination
    reason, ascii board) must describe the finished episode like `reward` and `terminated` do, while `obs` is the first
    observation of the next episode and info['final_observation'] the last one of the finished episode.  Checked against the
    CPU oracle stepped with the reference's semantics plus an explicit masked reset."""
    from ai_safety_gridworlds_b200 import GridworldGymEnv, make_spec
    N = 3000
    for name, kw in (("island_navigation_ex", {}), ("boat_race_ex", {"level": 3, "max_iterations": 30})):
        env = GridworldGymEnv(name, num_envs=N, **kw)
        spec = make_spec(name, autoreset_mode=0, **kw)
        orc = oracle_lib.Oracle(spec, N)
        orc.reset()
        obs, info = env.reset()
        np.testing.assert_array_equal(obs[:, 0].cpu().numpy(), orc.value_board)
        finished = 0
        for t in range(130):
            a = env.vector_env.random_actions(21, t)
            obs, reward, terminated, truncated, info = env.step(a)
            orc.step(a.cpu().numpy())
            ox = orc.observe()
            ctx = "%s t=%d" % (name, t)
            term = orc.terminated.astype(bool)
            np.testing.assert_array_equal(terminated.cpu().numpy(), term, err_msg=ctx)
            np.testing.assert_allclose(reward.cpu().numpy(), orc.reward, rtol=1e-6, atol=0, err_msg=ctx)
            # the finished episodes' final values
            np.testing.assert_array_equal(info["frame"].cpu().numpy(), ox["frame"], err_msg=ctx)
            np.testing.assert_allclose(info["cumulative_reward"].cpu().numpy(), ox["cumulative"], rtol=1e-6, atol=1e-5, err_msg=ctx)
            np.testing.assert_array_equal(info["extra_observations"]["termination_reason"].cpu().numpy(), orc.reason, err_msg=ctx)
            np.testing.assert_array_equal(info["ascii_codes"].cpu().numpy(), orc.board, err_msg=ctx)
            np.testing.assert_array_equal(info["final_observation"][:, 0].cpu().numpy(), orc.value_board, err_msg=ctx)
            for j, mname in enumerate(spec.metric_names):
                np.testing.assert_allclose(info["metrics_dict"][mname].cpu().numpy(), ox["metrics"][:, j], rtol=1e-12, atol=1e-12, err_msg=ctx)
            assert bool((info["frame"][terminated] > 0).all())
            finished += int(term.sum())
            orc.reset(term.astype(np.uint8))                               # the wrapper restarted exactly those environments
            np.testing.assert_array_equal(obs[:, 0].cpu().numpy(), orc.value_board, err_msg=ctx)
        assert finished > N // 2
        env.close()
        orc.close()
