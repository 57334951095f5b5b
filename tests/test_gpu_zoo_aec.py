"""The PettingZoo-AEC-signature wrapper over the CUDA backend, replaying traces recorded through the reference's
own GridworldZooAecEnv (same call sequence: agent_selection -> step(action), dead steps with None): observations,
`rewards`, the cumulative rewards `last()` returns, terminations, the live-agent list and the next selection
must all match after every step."""
import numpy as np
import pytest

from conftest import firemaker_aec_golden_names, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

NAMES = ["agent_1", "agent_2", "agent_S"]


def _vec(x, R):
    return np.zeros(R) if x is None or np.isscalar(x) else np.asarray(x, dtype=np.float64)


@pytest.mark.parametrize("name", firemaker_aec_golden_names())
def test_single_env_aec_drop_in_replays_reference_trace(name):
    from ai_safety_gridworlds_b200 import GridworldZooAecEnv
    d, meta = load_golden(name)
    env = GridworldZooAecEnv("firemaker_ex_ma", amount_agents=3, seed=meta["seed"], **meta["kwargs"])
    env.reset(seed=meta["seed"])
    codes = lambda o: np.vectorize(ord)(o[0]).astype(np.uint8)
    T = len(d["action"])
    it = env.agent_iter(max_iter=T)
    for t in range(T + 1):
        if t > 0:
            sel = next(it)
            assert sel == env.agent_selection == NAMES[int(d["agent"][t - 1])]
            obs, cum, term, trunc, info = env.last()
            assert bool(term) == bool(d["term"][t - 1][NAMES.index(sel)]) and trunc is False
            if d["action"][t - 1] < 0:
                with pytest.raises(ValueError):
                    env.step(1)                                   # a dead agent may only be stepped with None
                env.step(None)
            else:
                lo, hi = int(d["draw_ofs"][t - 1]), int(d["draw_ofs"][t])
                env.step(int(d["action"][t - 1]), replay_draws=d["draws"][lo:hi])
        ctx = "%s t=%d" % (name, t)
        np.testing.assert_array_equal(env.observe_info("agent_1")["ascii_codes"], d["board"][t], err_msg=ctx)
        np.testing.assert_array_equal(env.observe_info("agent_2")["info_observation_layers_cube"], d["cube"][t].astype(bool), err_msg=ctx)
        for i, (a, key) in enumerate(zip(NAMES, ("1", "2", "S"))):
            R = 3 if i == 2 else 2
            np.testing.assert_array_equal(codes(env.observe(a)), d["crop" + key][t], err_msg=ctx)
            np.testing.assert_array_equal(env.observe_info(a)["info_agent_observation_layers_cube"], d["lcrop" + key][t].astype(bool), err_msg=ctx)
            np.testing.assert_array_equal(_vec(env.rewards.get(a), R), d["reward" + key][t], err_msg=ctx + " rewards " + a)
            np.testing.assert_array_equal(_vec(env._cumulative_rewards.get(a), R), d["cum" + key][t], err_msg=ctx + " cumulative " + a)
            assert int(bool(env.terminations.get(a, True))) == d["term"][t][i], ctx
            assert int(a in env.agents) == d["alive"][t][i], ctx
        assert (NAMES.index(env.agent_selection) if env.agent_selection is not None else -1) == d["selection"][t], ctx
        assert env.get_step_no() == d["frame"][t], ctx
    env.close()


def test_batched_aec_matches_the_parallel_step_with_identity_order():
    """Three batched AEC steps ('1', '2', 'S') are the same three engine frames as one parallel step executed in
    identity order; with the same seed the Philox draws differ per call, so compare on a fire-free horizon."""
    from ai_safety_gridworlds_b200 import GridworldZooAecEnv, GridworldZooParallelEnv
    N = 512
    aec = GridworldZooAecEnv("firemaker_ex_ma", num_envs=N, seed=3, randomize_agent_actions_order=False, amount_agents=3)
    par = GridworldZooParallelEnv("firemaker_ex_ma", num_envs=N, seed=3, randomize_agent_actions_order=False, amount_agents=3)
    aec.reset(); par.reset()
    dev = aec.vector_env.device
    g = torch.Generator(device=dev); g.manual_seed(0)
    clean = torch.ones((N,), dtype=torch.bool, device=dev)      # environments in which no fire has started in either run
    for t in range(6):
        acts = {a: torch.randint(0, 5, (N,), device=dev, generator=g) for a in NAMES}
        total = {a: 0 for a in NAMES}
        for a in aec.agent_iter(max_iter=3):
            assert a == aec.agent_selection
            aec.step(acts[a])
            for b in NAMES:
                total[b] = total[b] + aec.rewards[b]
        obs, rewards, terms, truncs, infos = par.step(acts)
        # the two runs draw from different Philox calls: compare the environments that stayed fire-free in both
        clean &= ~((par.vector_env.board == ord("F")).flatten(1).any(1) | (aec.vector_env.board == ord("F")).flatten(1).any(1))
        assert torch.equal(aec.vector_env.board[clean], par.vector_env.board[clean])
        for b in NAMES:
            assert torch.equal(total[b][clean], rewards[b][clean]), (t, b)
            assert torch.equal(aec.observe(b)[clean], obs[b][clean])
    assert int(clean.sum()) > N // 4
    aec.close(); par.close()


def test_single_frame_order_matches_oracle(oracle_lib):
    """CUDA vs the C oracle on a Philox batch stepped one agent frame at a time."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    n = 777
    spec = make_spec("firemaker_ex_ma", autoreset_mode=1, max_iterations=50, amount_agents=3)
    env = FiremakerVectorEnv(n, device="cuda:0", seed=11, autoreset_mode=1, spec=spec)
    orc = oracle_lib.FiremakerOracle(spec, n, seed=11)
    orc.reset()
    rng = np.random.default_rng(2)
    for t in range(120):
        ag = t % 3
        a = rng.integers(0, 5, size=(n, 3)).astype(np.int32)
        order = np.tile(np.array([[ag, -1, -1]], np.int32), (n, 1))
        env.step(torch.from_numpy(a).to(env.device), torch.from_numpy(order).to(env.device))
        orc.step(a, order)
        assert np.array_equal(env.board.cpu().numpy(), orc.board), t
        assert np.array_equal(env.step_type.cpu().numpy(), orc.step_type), t
        assert np.array_equal(env.reward_workers.cpu().numpy(), orc.reward_w), t
        assert np.array_equal(env.reward_supervisor.cpu().numpy(), orc.reward_s), t
        assert np.array_equal(env.lcrop_supervisor.cpu().numpy(), orc.lcrop_s), t
    assert int((orc.board == ord("F")).sum()) > 0
    env.close(); orc.close()
