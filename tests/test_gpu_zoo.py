"""The PettingZoo-parallel-signature wrapper over the CUDA backend, replaying a trace that was recorded
through the reference's own GridworldZooParallelEnv (same call sequence, same return dicts); the
reference's shuffle orders and FireDrape draws are fed back through the wrapper's replay hooks."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("name", ["firemaker_s1", "firemaker_maxiter60_s2", "firemaker_2agents_maxiter40_s6"])
def test_single_env_drop_in_replays_reference_trace(name):
    from ai_safety_gridworlds_b200 import GridworldZooParallelEnv
    d, meta = load_golden(name)
    kw = dict(meta["kwargs"])
    if meta["amount_agents"] == 3:
        kw["amount_agents"] = 3                   # the 2-agent trace uses the wrapper's default, like the reference's
    env = GridworldZooParallelEnv("firemaker_ex_ma", seed=meta["seed"], **kw)
    obs, infos = env.reset(seed=meta["seed"])
    chars = ["1", "2", "S"] if meta["amount_agents"] == 3 else ["1", "S"]
    names = ["agent_" + c for c in chars]
    cols = [["1", "2", "S"].index(c) for c in chars]
    assert env.agents == names and sorted(obs) == sorted(names)
    assert obs["agent_1"].shape == (1, 5, 5) and obs["agent_S"].shape == (1, 33, 33) and obs["agent_1"].dtype.kind == "U"
    codes = lambda o: np.vectorize(ord)(o[0]).astype(np.uint8)
    np.testing.assert_array_equal(codes(obs["agent_1"]), d["crop1"][0])
    np.testing.assert_array_equal(codes(obs["agent_S"]), d["cropS"][0])
    np.testing.assert_array_equal(infos["agent_S"]["info_agent_observation_layers_cube"], d["lcropS"][0].astype(bool))
    assert infos["agent_1"]["info_observation_layers_order"] == meta["layer_order"]
    for t in range(1, len(d["actions"]) + 1):
        if not env.agents:                       # the recorder called reset() when every agent was done
            obs, infos = env.reset()
            np.testing.assert_array_equal(infos["agent_1"]["ascii_codes"], d["board"][t])
            continue
        a = d["actions"][t - 1]
        lo, hi = int(d["draw_ofs"][t - 1]), int(d["draw_ofs"][t])
        obs, rewards, terms, truncs, infos = env.step({n: int(a[k]) for n, k in zip(names, cols)},
                                                      replay_order=d["order"][t - 1], replay_draws=d["draws"][lo:hi])
        for n, c in zip(names, chars):
            np.testing.assert_array_equal(codes(obs[n]), d["crop" + c][t])
            np.testing.assert_array_equal(rewards[n], d["reward" + c][t])
            np.testing.assert_array_equal(infos[n]["cumulative_reward"], d["cum" + c][t])
        assert rewards["agent_1"].dtype == np.float64 and truncs["agent_S"] is False
        assert [terms[n] for n in names] == [bool(d["done"][t][k]) for k in cols]
        np.testing.assert_array_equal(infos["agent_1"]["ascii_codes"], d["board"][t])
        np.testing.assert_array_equal(infos["agent_1"]["info_observation_layers_cube"], d["cube"][t].astype(bool))
        assert list(infos["agent_1"]["metrics_dict"].keys()) == meta["metric_names"]
        assert list(infos["agent_1"]["metrics_dict"].values()) == list(d["metrics"][t])
    env.close()


def test_single_env_done_agents_and_float_observations():
    from ai_safety_gridworlds_b200 import GridworldZooParallelEnv
    env = GridworldZooParallelEnv("firemaker_ex_ma", seed=1, max_iterations=9, ascii_observation_format=False, amount_agents=3)
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
    env = GridworldZooParallelEnv("firemaker_ex_ma", num_envs=N, seed=7, max_iterations=60, amount_agents=3)
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


@pytest.mark.parametrize("name", ["islandma_default_s0", "islandma_homeostasis_s2", "islandma_level4_s5"])
def test_island_ma_single_env_drop_in_replays_reference_trace(name):
    """island_navigation_ex_ma through the parallel wrapper: agents finish one by one, leave `agents` and the returned
    dicts, and the recorder's reset() calls are replayed where every agent was done."""
    from ai_safety_gridworlds_b200 import GridworldZooParallelEnv
    d, meta = load_golden(name)
    env = GridworldZooParallelEnv("island_navigation_ex_ma", seed=meta["seed"], **meta["kwargs"])
    obs, infos = env.reset(seed=meta["seed"])
    names = ["agent_1", "agent_2"]
    assert env.agents == names and obs["agent_1"].shape == (1, 5, 5) and obs["agent_1"].dtype.kind == "U"
    codes = lambda o: np.vectorize(ord)(o[0]).astype(np.uint8)
    np.testing.assert_array_equal(codes(obs["agent_2"]), d["crop2"][0])
    assert infos["agent_1"]["info_observation_layers_order"] == meta["layer_order"]
    for t in range(1, len(d["actions"]) + 1):
        a = d["actions"][t - 1]
        if (a < 0).all():                        # the recorder called reset(): every agent was done
            assert env.agents == []
            with pytest.raises(ValueError):
                env.step({})
            obs, infos = env.reset()
            np.testing.assert_array_equal(infos["agent_1"]["ascii_codes"], d["board"][t])
            continue
        live = [n for i, n in enumerate(names) if a[i] >= 0]
        assert env.agents == live
        obs, rewards, terms, truncs, infos = env.step({n: int(a[names.index(n)]) for n in live}, replay_order=d["order"][t - 1])
        assert sorted(obs) == sorted(live) == sorted(rewards) == sorted(terms)
        for n in live:
            k = n[-1]
            i = names.index(n)
            np.testing.assert_array_equal(codes(obs[n]), d["crop" + k][t])
            np.testing.assert_allclose(rewards[n], d["reward" + k][t], rtol=1e-6)
            assert terms[n] == bool(d["done"][t][i]) and truncs[n] is False
            np.testing.assert_array_equal(infos[n]["info_agent_observation_layers_cube"], d["lcrop" + k][t].astype(bool))
            np.testing.assert_allclose(infos[n]["cumulative_reward"], d["cum"][t][i], rtol=1e-6)
            assert int(infos[n]["observation_direction"]) == d["odir"][t][i]
        n0 = live[0]
        np.testing.assert_array_equal(infos[n0]["ascii_codes"], d["board"][t])
        np.testing.assert_array_equal(infos[n0]["info_observation_layers_cube"], d["cube"][t].astype(bool))
        np.testing.assert_allclose(list(infos[n0]["metrics_dict"].values()), d["metrics"][t], rtol=1e-9)
    env.close()


def test_island_ma_batched_parallel_and_aec_forms():
    from ai_safety_gridworlds_b200 import GridworldZooAecEnv, GridworldZooParallelEnv
    N = 1024
    env = GridworldZooParallelEnv("island_navigation_ex_ma", num_envs=N, seed=3)
    obs, infos = env.reset()
    assert obs["agent_1"].shape == (N, 1, 5, 5) and obs["agent_1"].dtype == torch.uint8 and len(env.agents) == 2
    ended = 0
    for t in range(30):
        acts = {a: torch.randint(0, 5, (N,), device=env.vector_env.device) for a in env.possible_agents}
        obs, rewards, terms, truncs, infos = env.step(acts)
        assert rewards["agent_2"].shape == (N, 8) and rewards["agent_2"].dtype == torch.float64
        ended += int((terms["agent_1"] & terms["agent_2"]).sum())
    assert ended > 0                                             # games ended and restarted inside the step
    env.close()
    aec = GridworldZooAecEnv("island_navigation_ex_ma", num_envs=N, seed=3)
    aec.reset()
    for k, a in zip(range(8), aec.agent_iter()):
        assert a == ["agent_1", "agent_2"][k % 2]
        obs, cum, term, trunc, info = aec.last()
        assert obs.shape == (N, 1, 5, 5) and cum.shape == (N, 8)
        aec.step(torch.randint(0, 5, (N,), device=aec.vector_env.device))
    assert int(aec.get_step_no().max()) <= 8
    aec.close()


@pytest.mark.parametrize("name", ["savanna_maxiter40_s2", "savanna_two_agents_s7", "savanna_exp_food_sharing",
                                  "savanna_exp_food_drink_homeostasis_danger_gold_silver", "savanna_exp_food_sustainability"])
def test_savanna_single_env_drop_in_replays_reference_trace(name):
    """aintelope_savanna (and its experiment overlays, by their factory names) through the parallel wrapper.  The reference draws
    a new layout for every game from its Generator; the recorded layouts are replayed through the per-environment maps tensor."""
    from ai_safety_gridworlds_b200 import GridworldZooParallelEnv
    d, meta = load_golden(name)
    A = meta["amount_agents"]
    env = GridworldZooParallelEnv(meta["env"], seed=meta["seed"], **meta["kwargs"])
    ve = env.vector_env
    maps = torch.zeros((1,) + d["maps"].shape[1:], dtype=torch.uint8, device=ve.device)
    ve.set_maps(maps, 0)                                         # GW_IMA_MAPS_STATIC: the caller writes the layouts
    maps.copy_(torch.from_numpy(d["maps"][0]).to(ve.device))
    obs, infos = env.reset(seed=meta["seed"])
    names = ["agent_0", "agent_1"][:A]
    V = meta["view"]
    assert env.agents == names and obs["agent_0"].shape == (1, V, V) and obs["agent_0"].dtype.kind == "U"
    codes = lambda o: np.vectorize(ord)(o[0]).astype(np.uint8)
    np.testing.assert_array_equal(codes(obs["agent_0"]), d["crop"][0, 0])
    assert infos["agent_0"]["info_observation_layers_order"] == meta["layer_order"]
    assert list(infos["agent_0"]["metrics_dict"].keys()) == list(dict.fromkeys(meta["metric_names"]))
    for t in range(1, len(d["actions"]) + 1):
        a = d["actions"][t - 1]
        if (a < 0).all():                        # the recorder called reset(): every agent was done
            assert env.agents == []
            maps.copy_(torch.from_numpy(d["maps"][t]).to(ve.device))
            obs, infos = env.reset()
            np.testing.assert_array_equal(infos["agent_0"]["ascii_codes"], d["board"][t])
            continue
        live = [n for i, n in enumerate(names) if a[i] >= 0]
        assert env.agents == live
        order = list(d["order"][t - 1]) + [-1] * (2 - A)
        picks = d["draws"][t - 1] if meta["env"] == "food_sustainability" else None     # the drapes' Generator.choice picks of this step
        obs, rewards, terms, truncs, infos = env.step({n: int(a[names.index(n)]) for n in live}, replay_order=order, replay_draws=picks)
        assert sorted(obs) == sorted(live) == sorted(rewards) == sorted(terms)
        for n in live:
            i = names.index(n)
            np.testing.assert_array_equal(codes(obs[n]), d["crop"][t, i])
            np.testing.assert_allclose(rewards[n], d["reward"][t, i], rtol=1e-6, atol=1e-6)
            assert terms[n] == bool(d["done"][t][i]) and truncs[n] is False
            np.testing.assert_array_equal(infos[n]["info_agent_observation_layers_cube"], d["lcrop"][t, i].astype(bool))
            np.testing.assert_allclose(infos[n]["cumulative_reward"], d["cum"][t][i], rtol=1e-6, atol=1e-5)
            assert int(infos[n]["observation_direction"]) == d["odir"][t][i]
        n0 = live[0]
        np.testing.assert_array_equal(infos[n0]["ascii_codes"], d["board"][t])
        np.testing.assert_array_equal(infos[n0]["info_observation_layers_cube"], d["cube"][t].astype(bool))
        got = infos[n0]["metrics_dict"]
        for label, want in zip(meta["metric_names"], d["metrics"][t]):
            if not np.isnan(want):
                assert got[label] == pytest.approx(want, rel=1e-12, abs=1e-12), (t, label)
    env.close()


def test_savanna_batched_parallel_and_aec_forms():
    from ai_safety_gridworlds_b200 import GridworldZooAecEnv, GridworldZooParallelEnv
    N = 512
    env = GridworldZooParallelEnv("aintelope_savanna", num_envs=N, seed=3, amount_agents=2, max_iterations=12)
    obs, infos = env.reset()
    assert obs["agent_0"].shape == (N, 1, 21, 21) and obs["agent_0"].dtype == torch.uint8 and len(env.agents) == 2
    first_maps = env.vector_env.maps.clone()
    assert int((first_maps != first_maps[0]).any(dim=(1, 2)).sum()) > N // 2          # every environment drew its own layout
    ended = 0
    for t in range(14):
        acts = {a: torch.randint(0, 5, (N,), device=env.vector_env.device) for a in env.possible_agents}
        obs, rewards, terms, truncs, infos = env.step(acts)
        assert rewards["agent_1"].shape == (N, 4) and rewards["agent_1"].dtype == torch.float64
        ended += int((terms["agent_0"] & terms["agent_1"]).sum())
    assert ended == 2 * N                                        # max_iterations = 12 frames = 6 parallel steps of two agents
    assert bool((env.vector_env.maps != first_maps).any())       # frequency 3: a new layout for every game
    env.close()
    aec = GridworldZooAecEnv("food_unbounded", num_envs=N, seed=3)
    aec.reset()
    for k, a in zip(range(6), aec.agent_iter()):
        assert a == "agent_0"
        obs, cum, term, trunc, info = aec.last()
        assert obs.shape == (N, 1, 21, 21) and cum.shape == (N, 1)
        aec.step(torch.randint(0, 5, (N,), device=aec.vector_env.device))
    aec.close()
    sus = GridworldZooParallelEnv("food_sustainability", num_envs=N, seed=3, max_iterations=40)     # tiles spawn and vanish during play
    sus.reset()
    counts = []
    for t in range(30):
        sus.step({"agent_0": torch.randint(0, 5, (N,), device=sus.vector_env.device)})
        counts.append((sus.vector_env.board == ord("F")).sum(dim=(1, 2)))
    counts = torch.stack(counts)
    assert int(counts.max()) > 2 and bool((counts[-1] != counts[0]).any())
    avail = sus.vector_env.observe()["metrics"][:, sus.vector_env.spec.metric_names.index("FoodAvailability")]
    assert torch.equal(torch.ceil(avail).long(), (sus.vector_env.live_maps == ord("F")).sum(dim=(1, 2)))     # ceil(availability) tiles are visible
    sus.close()
    with pytest.raises(NotImplementedError):
        GridworldZooParallelEnv("aintelope_savanna", sustainability_challenge=True, amount_drink_holes=1)   # a spawning drape must be the only one
