"""GPU parity of aintelope_savanna (include/gwsim_sav.h, csrc/gwsim_sav.cuh): the traces recorded from the reference replayed
(the sustainability challenge's among them: regrowth through the device's pow, the reference's tile picks) through the CUDA path with the same checks as the oracle's (tests/test_oracle_savanna_golden.py::replay), and seeded batches
with device-drawn layouts against the scalar oracle.  Bytes bit-exact, rewards to 1e-6."""
import numpy as np
import pytest

from conftest import load_golden, savanna_golden_names, spec_for
from test_oracle_savanna_golden import replay

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


class CudaAdapter(object):
    """SavannaVectorEnv behind the interface of oracle.pyoracle.SavannaOracle (numpy in / numpy views out)."""

    def __init__(self, spec, n):
        from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
        self.env = SavannaVectorEnv(n, spec=spec, autoreset_mode=0)
        self.dev = self.env.device

    def set_maps(self, maps, mode):
        s = self.env.spec
        self.env.set_maps(torch.from_numpy(np.ascontiguousarray(maps, np.uint8)).to(self.dev).reshape(-1, s.height, s.width).contiguous(), mode)

    def reset(self):
        self.env.reset()

    def step(self, actions, order, draws=None):
        a = torch.from_numpy(np.ascontiguousarray(actions, np.int32)).to(self.dev)
        o = None if order is None else torch.from_numpy(np.ascontiguousarray(order, np.int32)).to(self.dev)
        d = None if draws is None else torch.from_numpy(np.ascontiguousarray(draws, np.float64)).to(self.dev)
        self.env.step(a, o, d)

    def observe(self):
        ex = self.env.observe(all_slots=True)
        return {k: v.cpu().numpy() for k, v in ex.items()}

    def close(self):
        self.env.close()

    def __getattr__(self, name):
        if name in ("board", "cube", "crop", "lcrop", "reward", "terminated", "step_type"):
            return getattr(self.env, name).cpu().numpy()
        raise AttributeError(name)


@pytest.mark.parametrize("name", savanna_golden_names())
def test_cuda_replays_savanna_reference_trace(name):
    d, meta = load_golden(name)
    meta = dict(meta, name=name)
    replay(d, meta, spec_for(meta), CudaAdapter, n=3)


@pytest.mark.parametrize("kwargs", [dict(max_iterations=25), dict(sustainability_challenge=True, max_iterations=45, FOOD_GROWTH_LIMIT=9),
                                    dict(sustainability_challenge=True, amount_agents=2, map_width=7, map_height=7, max_iterations=35),
                                    dict(sustainability_challenge=True, amount_drink_holes=2, amount_predators=2, amount_water_tiles=2, max_iterations=30,
                                         penalise_oversatiation=True, use_food_availability_metric_instead_of_spawning_tiles=True,
                                         use_drink_availability_metric_instead_of_spawning_tiles=True),
                                    dict(amount_predators=5, amount_agents=2, amount_water_tiles=3, max_iterations=40),
                                    dict(observation_direction_mode=2, action_direction_mode=2, amount_agents=2, amount_predators=2, max_iterations=40,
                                         observation_radius=[5, 5, 5, 5]),
                                    dict(amount_predators=4, PREDATOR_MOVEMENT_PROBABILITY=0.9, map_width=8, map_height=7, max_iterations=30), dict(amount_agents=2, amount_drink_holes=2, penalise_oversatiation=True, max_iterations=30,
                                                                    amount_gold_deposits=2, amount_water_tiles=3),
                                    dict(map_width=9, map_height=8, amount_food_patches=3, observation_radius=[4, 4, 4, 4], max_iterations=20,
                                         amount_agents=2, thirst_hunger_death=True, penalise_oversatiation=True, FOOD_DEFICIENCY_LIMIT=-2)])
@pytest.mark.parametrize("mode", [0, 1])
def test_savanna_batch_matches_oracle(kwargs, mode, oracle_lib):
    """Device-drawn layouts (Philox Fisher-Yates per game), Philox-shuffled agent order, both auto-reset modes, a ragged batch:
    every tensor against the scalar oracle."""
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
    spec = make_spec("aintelope_savanna", autoreset_mode=mode, **kwargs)
    N = 300 + 7
    env = SavannaVectorEnv(N, spec=spec, env_index_base=1000, seed=5, autoreset_mode=mode)
    orc = oracle_lib.SavannaOracle(spec.with_autoreset(mode), N, env_index_base=1000, seed=5)
    orc.set_maps(orc.maps, 1)                              # GW_IMA_MAPS_SHUFFLE_EVERY_GAME
    orc.reset()
    rng = np.random.default_rng(3)
    games = 0
    for t in range(120):
        if t > 0:
            a = rng.integers(0, 9 if kwargs.get("action_direction_mode") == 2 else 5, size=(N, 2)).astype(np.int32)   # mode 2: TURN_* = 5..8
            env.step(torch.from_numpy(a).to(env.device))
            orc.step(a)
        ctx = "t=%d" % t
        np.testing.assert_array_equal(env.maps.cpu().numpy().reshape(N, -1), orc.maps, err_msg=ctx)
        np.testing.assert_array_equal(env.board.cpu().numpy(), orc.board, err_msg=ctx)
        np.testing.assert_array_equal(env.cube.cpu().numpy(), orc.cube, err_msg=ctx)
        np.testing.assert_array_equal(env.crop.cpu().numpy(), orc.crop, err_msg=ctx)
        np.testing.assert_array_equal(env.lcrop.cpu().numpy(), orc.lcrop, err_msg=ctx)
        np.testing.assert_array_equal(env.step_type.cpu().numpy(), orc.step_type, err_msg=ctx)
        np.testing.assert_array_equal(env.terminated.cpu().numpy(), orc.terminated, err_msg=ctx)
        np.testing.assert_allclose(env.reward.cpu().numpy(), orc.reward, rtol=1e-6, atol=1e-6, err_msg=ctx)
        if t % 10 == 0:
            ex, ox = env.observe(all_slots=True), orc.observe()
            np.testing.assert_allclose(ex["metrics"].cpu().numpy(), ox["metrics"], rtol=1e-12, atol=1e-12, err_msg=ctx)
            np.testing.assert_allclose(ex["cumulative"].cpu().numpy(), ox["cumulative"], rtol=1e-5, atol=1e-4, err_msg=ctx)
            for k in ("frame", "pos", "directions"):
                np.testing.assert_array_equal(ex[k].cpu().numpy(), ox[k], err_msg=ctx + " " + k)
        games += int((orc.step_type >= 2).all(axis=1).sum())
    st = env.stats()
    assert st["episodes"] == games and games > N
    env.close()
    orc.close()
