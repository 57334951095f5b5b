"""The PettingZoo-parallel-signature wrapper over the CUDA backend, replaying a trace that was recorded
through the reference's own GridworldZooParallelEnv (same call sequence, same return dicts); the
reference's shuffle orders and FireDrape draws are fed back through the wrapper's replay hooks."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("name", ["firemaker_s1", "firemaker_maxiter60_s2"])
def test_single_env_drop_in_replays_reference_trace(name):
    from ai_safety_gridworlds_b200 import GridworldZooParallelEnv
    d, meta = load_golden(name)
    env = GridworldZooParallelEnv("firemaker_ex_ma", amount_agents=3, seed=meta["seed"], **meta["kwargs"])
    obs, infos = env.reset(seed=meta["seed"])
    assert env.agents == ["agent_1", "agent_2", "agent_S"]
    assert obs["agent_1"].shape == (1, 5, 5) and obs["agent_S"].shape == (1, 33, 33) and obs["agent_1"].dtype.kind == "U"
    codes = lambda o: np.vectorize(ord)(o[0]).astype(np.uint8)
    np.testing.assert_array_equal(codes(obs["agent_1"]), d["crop1"][0])
    np.testing.assert_array_equal(codes(obs["agent_S"]), d["cropS"][0])
    np.testing.assert_array_equal(infos["agent_2"]["info_agent_observation_layers_cube"], d["lcrop2"][0].astype(bool))
    assert infos["agent_1"]["info_observation_layers_order"] == meta["layer_order"]
    names = ["agent_1", "agent_2", "agent_S"]
    for t in range(1, len(d["actions"]) + 1):
        if not env.agents:                       # the recorder called reset() when every agent was done
            obs, infos = env.reset()
            np.testing.assert_array_equal(infos["agent_1"]["ascii_codes"], d["board"][t])
            continue
        a = d["actions"][t - 1]
        lo, hi = int(d["draw_ofs"][t - 1]), int(d["draw_ofs"][t])
        obs, rewards, terms, truncs, infos = env.step({n: int(a[i]) for i, n in enumerate(names)},
                                                      replay_order=d["order"][t - 1], replay_draws=d["draws"][lo:hi])
        np.testing.assert_array_equal(codes(obs["agent_1"]), d["crop1"][t])
        np.testing.assert_array_equal(codes(obs["agent_2"]), d["crop2"][t])
        np.testing.assert_array_equal(codes(obs["agent_S"]), d["cropS"][t])
        np.testing.assert_array_equal(rewards["agent_1"], d["reward1"][t])
        np.testing.assert_array_equal(rewards["agent_S"], d["rewardS"][t])
        assert rewards["agent_1"].dtype == np.float64 and truncs["agent_S"] is False
        assert [terms[n] for n in names] == [bool(x) for x in d["done"][t]]
        np.testing.assert_array_equal(infos["agent_1"]["ascii_codes"], d["board"][t])
        np.testing.assert_array_equal(infos["agent_1"]["info_observation_layers_cube"], d["cube"][t].astype(bool))
        assert list(infos["agent_1"]["metrics_dict"].values()) == list(d["metrics"][t])
        np.testing.assert_array_equal(infos["agent_S"]["cumulative_reward"], d["cumS"][t])
    env.close()


def test_single_env_done_agents_and_float_observations():
    from ai_safety_gridworlds_b200 import GridworldZooParallelEnv
    env = GridworldZooParallelEnv("firemaker_ex_ma", seed=1, max_iterations=9, ascii_observation_format=False)
    obs, _ = env.reset()
    assert obs["agent_S"].dtype == np.float32 and obs["agent_S"].shape == (1, 33, 33)
    assert obs["agent_S"][0, 0, 0] == 1.0                       # beyond the board: what_lies_outside '#' -> value 1.0
    acts = {a: 1 for a in env.possible_agents}
    for k in range(3):
        obs, rewards, terms, truncs, infos = env.step(acts)
    assert all(terms.values()) and env.agents == []             # frame 9 >= max_iterations: every agent LAST
    with pytest.raises(ValueError):
        env.step(acts)
    env.reset()
    assert len(env.agents) == 3
    env.close()


def test_batched_form():
    from ai_safety_gridworlds_b200 import GridworldZooParallelEnv
    N = 2048
    env = GridworldZooParallelEnv("firemaker_ex_ma", num_envs=N, seed=7, max_iterations=60)
    obs, infos = env.reset()
    assert obs["agent_1"].shape == (N, 1, 5, 5) and obs["agent_S"].shape == (N, 1, 33, 33) and obs["agent_1"].dtype == torch.uint8
    done_seen = 0
    for t in range(25):
        acts = {a: torch.randint(0, 5, (N,), device=env.vector_env.device) for a in env.possible_agents}
        obs, rewards, terms, truncs, infos = env.step(acts)
        assert rewards["agent_1"].shape == (N, 2) and rewards["agent_S"].shape == (N, 3) and rewards["agent_S"].dtype == torch.float64
        assert terms["agent_2"].dtype == torch.bool and not bool(truncs["agent_2"].any())
        done_seen += int(terms["agent_1"].sum())
        # each agent sees itself at the centre of its own view
        assert bool((obs["agent_1"][:, 0, 2, 2] == ord("1")).all()) and bool((obs["agent_S"][:, 0, 16, 16] == ord("S")).all())
    assert done_seen == N                                        # 60 frames = 20 parallel steps: every game ended once and restarted
    assert len(env.agents) == 3
    env.close()
