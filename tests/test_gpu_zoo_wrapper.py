"""The PettingZoo parallel wrapper's own options against traces recorded through the reference's wrapper
(oracle/record_zoo_wrapper.py): `layers_order_in_cube` / `layers_order_in_cube_per_agent` (unknown names = zero planes), the
global and per-agent (relative) coordinate infos, `observation_space(s)`, `use_transitions` / `flatten_observations`, the float
board format, and `test_death` with the reference's own generator draws replayed."""
import json

import numpy as np
import pytest

from conftest import load_golden, zoo_golden_names

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
NAMES = ["agent_1", "agent_2"]


def _codes(o):
    a = np.asarray(o)
    if a.dtype.kind == "U":
        return np.vectorize(lambda ch: ord(ch) if ch else 0)(a).astype(np.uint8)
    return a.astype(np.float32)


def _pairs(d):
    return {k: [tuple(p) for p in v] for k, v in d.items()}


@pytest.mark.parametrize("name", zoo_golden_names())
def test_wrapper_options_replay_the_reference(name):
    from ai_safety_gridworlds_b200.helpers.gridworld_zoo_parallel_env import GridworldZooParallelEnv
    d, meta = load_golden(name)
    steps = json.loads(str(d["steps_json"]))
    env = GridworldZooParallelEnv("island_navigation_ex_ma", seed=meta["seed"], **meta["wrapper"], **meta["env_kwargs"])
    for nm in NAMES:
        sp = env.observation_space(nm)
        assert list(sp.shape) == meta["spaces"][nm]["shape"] and str(sp.dtype) == meta["spaces"][nm]["dtype"], nm
        assert env.observation_spaces[nm] is sp
    for t, rec in enumerate(steps):
        ctx = "%s t=%d" % (name, t)
        rewards = terms = None
        if rec["actions"] is None or rec["actions"] == [-1, -1]:
            obs, infos = env.reset()
        else:
            acts = {nm: rec["actions"][i] for i, nm in enumerate(NAMES) if rec["actions"][i] >= 0}
            assert sorted(acts) == sorted(env.agents), ctx
            obs, rewards, terms, truncs, infos = env.step(acts, replay_order=rec["order"], replay_death_draws=rec["death_draws"])
        assert list(env.agents) == rec["agents_after"], ctx
        for nm in NAMES:
            pa = rec["per_agent"][nm]
            if pa["obs"] is None:
                assert nm not in obs, ctx
            else:
                got = _codes(obs[nm])
                assert list(got.shape) == pa["obs_shape"], ctx
                if not (meta["wrapper"].get("use_transitions") and meta["wrapper"].get("flatten_observations")):
                    # (the reference's space says (2, cells) for that combination while its observation is state.flatten(): (2 * cells,))
                    assert env.observation_space(nm).contains(obs[nm]), ctx
                if got.dtype == np.uint8:
                    np.testing.assert_array_equal(got, np.array(pa["obs"], np.uint8), err_msg=ctx)
                else:
                    np.testing.assert_allclose(got, np.array(pa["obs"], np.float32), rtol=0, atol=0, err_msg=ctx)
            if rewards is not None:
                assert (nm in rewards) == pa["has_reward"], ctx
                if pa["has_reward"] and pa["reward"] is not None:
                    np.testing.assert_allclose(np.asarray(rewards[nm], np.float64), pa["reward"], rtol=1e-6, atol=0, err_msg=ctx)
                assert (nm in terms) == pa["has_done"], ctx
                if pa["has_done"]:
                    assert bool(terms[nm]) == pa["done"], ctx
            if "layers_order" not in pa:
                assert nm not in infos, ctx
                continue
            info = infos[nm]
            assert list(info["info_observation_layers_order"]) == pa["layers_order"], ctx
            np.testing.assert_array_equal(np.asarray(info["info_observation_layers_cube"]).astype(np.uint8), np.array(pa["layers_cube"], np.uint8), err_msg=ctx)
            assert list(info["info_agent_observation_layers_order"]) == pa["agent_layers_order"], ctx
            np.testing.assert_array_equal(np.asarray(info["info_agent_observation_layers_cube"]).astype(np.uint8),
                                          np.array(pa["agent_layers_cube"], np.uint8), err_msg=ctx)
            assert _pairs(info["info_observation_coordinates"]) == _pairs(pa["coordinates"]), ctx
            want = pa["agent_coordinates"]
            got = info["info_agent_observation_coordinates"]
            assert (got == [] and want == []) or _pairs(got) == _pairs(want), ctx
            assert sorted(info["info_agent_observation_layers_dict"].keys()) == pa["agent_layers_dict_keys"], ctx
    env.close()


def test_wrapper_refuses_what_it_does_not_honour():
    from ai_safety_gridworlds_b200.helpers.gridworld_zoo_parallel_env import GridworldZooParallelEnv
    for kw in ({"occlusion_in_layers": True}, {"use_multi_discrete_action_space": True}, {"ascii_attributes_format": True},
               {"layers_order_in_cube_per_agent": {"agent_9": []}}):
        with pytest.raises((NotImplementedError, ValueError)):
            GridworldZooParallelEnv("island_navigation_ex_ma", **kw)
    with pytest.raises(NotImplementedError):
        GridworldZooParallelEnv("island_navigation_ex_ma", num_envs=8, object_coordinates_in_observation=True)


def test_batched_wrapper_spaces_layers_death_quit_and_rgb():
    from ai_safety_gridworlds_b200.helpers.gridworld_zoo_parallel_env import GridworldZooParallelEnv, _philox_uniform
    N = 513
    env = GridworldZooParallelEnv("firemaker_ex_ma", amount_agents=3, num_envs=N, seed=7, test_death=True, test_death_probability=0.1,
                                  layers_order_in_cube=["F", "zz", "#"], layers_order_in_cube_per_agent={"agent_S": ["S", "1"]},
                                  use_transitions=True)
    assert env.observation_space("agent_S").shape == (2, 33, 33) and env.observation_space("agent_1").shape == (2, 5, 5)
    obs, infos = env.reset()
    assert tuple(obs["agent_S"].shape) == (N, 2, 33, 33) and not bool(obs["agent_S"][:, 0].any())
    prev = {a: o[:, 1].clone() for a, o in obs.items()}
    dead = {a: torch.zeros(N, dtype=torch.bool, device=env.vector_env.device) for a in env.possible_agents}
    g = torch.Generator(device=env.vector_env.device); g.manual_seed(1)
    for t in range(25):
        acts = {a: torch.randint(0, 5, (N,), device=env.vector_env.device, generator=g) for a in env.possible_agents}
        obs, rewards, terms, truncs, infos = env.step(acts)
        really_all = torch.stack([infos[a]["step_type"] >= 2 for a in env.possible_agents], dim=1)
        for i, a in enumerate(env.possible_agents):
            assert torch.equal(obs[a][:, 0], prev[a])                          # use_transitions: (previous view, view)
            prev[a] = obs[a][:, 1].clone()
            u = _philox_uniform(7 ^ 0x7e57dea7, 0, N, (t + 1) * 3 + i, env.vector_env.device)
            really = really_all[:, i]
            was = dead[a].clone()
            dead[a] = was | ((~was) & (~really) & (u < 0.1))
            assert torch.equal(terms[a], really | dead[a]), (t, a)
            assert not bool(rewards[a][was].any())                                # a dead agent's reward row is zeroed
        over = really_all.all(dim=1)
        for a in env.possible_agents:
            dead[a] &= ~over                                                      # a finished game restarts with everybody alive
        cube = infos["agent_1"]["info_observation_layers_cube"]
        assert infos["agent_1"]["info_observation_layers_order"] == ["F", "zz", "#"] and tuple(cube.shape) == (N, 3, 17, 17)
        assert not bool(cube[:, 1].any())
        assert torch.equal(cube[:, 2], env.vector_env.cube[:, 1].bool()) and torch.equal(cube[:, 0], env.vector_env.cube[:, 6].bool())
        acube = infos["agent_S"]["info_agent_observation_layers_cube"]
        assert tuple(acube.shape) == (N, 2, 33, 33) and torch.equal(acube[:, 0], env.vector_env.lcrop_supervisor[:, 7].bool())
    assert int(sum(int(v.sum()) for v in dead.values())) > 0
    rgb = env.render("rgb_array")
    assert tuple(rgb.shape) == (N, 3, 17, 17) and rgb.dtype == torch.uint8
    with pytest.raises(NotImplementedError):                                      # QUIT in the batched form is refused, not mis-played
        env.step({a: torch.full((N,), 9, device=env.vector_env.device) for a in env.possible_agents})
    env.close()


def test_batched_infos_describe_the_finished_game(oracle_lib):
    """On the step that ends a game the batched wrapper's infos (cumulative rewards, frame, metrics, the boards and views they
    carry) describe the finished game, the returned observations belong to the next one -- against the CPU oracle stepped with the
    reference's semantics plus an explicit masked reset."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.helpers.gridworld_zoo_parallel_env import GridworldZooParallelEnv
    N = 300
    env = GridworldZooParallelEnv("firemaker_ex_ma", amount_agents=3, num_envs=N, seed=5, max_iterations=60)
    spec = make_spec("firemaker_ex_ma", autoreset_mode=0, amount_agents=3, max_iterations=60)
    orc = oracle_lib.FiremakerOracle(spec, N, seed=5)
    orc.reset()                                                           # the constructor reset the environments once: the same call count
    dev = env.vector_env.device
    rng = np.random.default_rng(2)
    ended = 0
    for t in range(50):
        a = rng.integers(0, 5, size=(N, 3)).astype(np.int32)
        obs, rewards, terms, truncs, infos = env.step({nm: torch.from_numpy(a[:, i].copy()).to(dev) for i, nm in enumerate(env.possible_agents)})
        orc.step(a)
        ox = orc.observe()
        ctx = "t=%d" % t
        over = (orc.step_type >= 2).all(axis=1)
        for i, nm in enumerate(env.possible_agents):
            np.testing.assert_array_equal(terms[nm].cpu().numpy(), orc.terminated[:, i].astype(bool), err_msg=ctx)
            want_r = orc.reward_s if i == 2 else orc.reward_w[:, i]
            np.testing.assert_array_equal(rewards[nm].cpu().numpy(), want_r.astype(np.float64), err_msg=ctx)
            info = infos[nm]
            np.testing.assert_array_equal(info["frame"].cpu().numpy(), ox["frame"], err_msg=ctx)
            np.testing.assert_array_equal(info["ascii_codes"].cpu().numpy(), orc.board, err_msg=ctx)
            want_c = orc.crop_s if i == 2 else orc.crop_w[:, i]
            np.testing.assert_array_equal(info["info_agent_observations"].cpu().numpy(), want_c, err_msg=ctx)
            cum = ox["cumulative"][:, 4:7] if i == 2 else ox["cumulative"][:, 2 * i:2 * i + 2]
            np.testing.assert_array_equal(info["cumulative_reward"].cpu().numpy(), cum.astype(np.float64), err_msg=ctx)
        assert bool((infos["agent_1"]["frame"][torch.from_numpy(over).to(dev)] == 60).all())
        ended += int(over.sum())
        orc.reset(over.astype(np.uint8))
        for i, nm in enumerate(env.possible_agents):
            want_c = orc.crop_s if i == 2 else orc.crop_w[:, i]
            np.testing.assert_array_equal(obs[nm][:, 0].cpu().numpy(), want_c, err_msg=ctx)
    assert ended == 2 * N                                                 # 60 frames = 20 parallel steps per game
    env.close()
    orc.close()
