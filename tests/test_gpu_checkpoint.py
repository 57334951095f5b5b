"""Checkpoint / resume (the reference pickles its environments, safety_game_mo.py:406-419): a batch saved with
state_dict(), sent through torch.save / torch.load and restored into a freshly built batch continues bit for bit -- every
kernel family, Philox-driven games included (the call counter that keys the streams travels with the state)."""
import io

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _roundtrip(sd):
    buf = io.BytesIO()
    torch.save(sd, buf)
    buf.seek(0)
    return torch.load(buf, weights_only=False)


def _outputs(env, names):
    return {nm: getattr(env, nm).clone() for nm in names if getattr(env, nm, None) is not None}


def _run(make, action_fn, names, before=37, after=41):
    a = make()
    for t in range(before):
        a.step(action_fn(a, t))
    sd = _roundtrip(a.state_dict())
    b = make()
    b.load_state_dict(sd)
    for nm, v in _outputs(a, names).items():
        assert torch.equal(v, getattr(b, nm)), nm
    for t in range(before, before + after):
        act = action_fn(a, t)
        a.step(act)
        b.step(act.clone())
        for nm, v in _outputs(a, names).items():
            assert torch.equal(v, getattr(b, nm)), (nm, t)
    assert torch.equal(a.state, b.state)
    a.close()
    b.close()


def _rand(n, cols, hi=4):
    def fn(env, t):
        g = torch.Generator(device=env.device)
        g.manual_seed(1000 + t)
        shape = (n,) if cols == 0 else (n, cols)
        return torch.randint(0, hi + 1, shape, dtype=torch.int32, device=env.device, generator=g)
    return fn


def test_single_agent_and_classic_batches_resume():
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.vector_env import VectorEnv
    from ai_safety_gridworlds_b200.classic_env import ClassicVectorEnv
    from ai_safety_gridworlds_b200.sokoban_env import SokobanVectorEnv
    N = 777
    _run(lambda: VectorEnv("island_navigation_ex", N, autoreset_mode=1, env_index_base=5), _rand(N, 0),
         ("board", "cube", "value_board", "reward", "terminated", "step_type", "reason"))
    specs = [make_spec(nm, autoreset_mode=1) for nm in ("safe_interruptibility", "tomato_watering", "friend_foe", "whisky_gold")]
    _run(lambda: ClassicVectorEnv(specs, [200, 201, 202, 174], seed=4, autoreset_mode=1), _rand(N, 0),
         ("board", "value_board", "reward", "terminated", "step_type", "reason", "actual"), before=60, after=80)
    sok = make_spec("side_effects_sokoban", level=2, autoreset_mode=1)
    _run(lambda: SokobanVectorEnv(sok, N, autoreset_mode=1), _rand(N, 0), ("board", "reward", "terminated", "step_type"))


def test_multi_agent_batches_resume_with_their_philox_streams():
    from ai_safety_gridworlds_b200 import make_spec
    from ai_safety_gridworlds_b200.firemaker_env import FiremakerVectorEnv
    from ai_safety_gridworlds_b200.island_ma_env import IslandMaVectorEnv
    from ai_safety_gridworlds_b200.savanna_env import SavannaVectorEnv
    N = 300
    fm = make_spec("firemaker_ex_ma", autoreset_mode=1, amount_agents=3, max_iterations=90)
    _run(lambda: FiremakerVectorEnv(N, seed=9, env_index_base=11, autoreset_mode=1, spec=fm), _rand(N, 3),
         ("board", "cube", "crop_supervisor", "lcrop_workers", "reward_workers", "reward_supervisor", "terminated", "step_type"))
    ima = make_spec("island_navigation_ex_ma", autoreset_mode=1, map_randomization_frequency=3)
    _run(lambda: IslandMaVectorEnv(N, seed=2, autoreset_mode=1, spec=ima), _rand(N, 2),
         ("board", "cube", "crop", "lcrop", "reward", "terminated", "step_type", "maps"))
    sav = make_spec("food_sustainability", autoreset_mode=1)
    _run(lambda: SavannaVectorEnv(N, seed=3, autoreset_mode=1, spec=sav), _rand(N, 2),
         ("board", "cube", "crop", "reward", "terminated", "step_type", "maps", "availability", "live_maps"))


def test_checkpoint_refuses_a_different_batch():
    from ai_safety_gridworlds_b200.vector_env import VectorEnv
    a = VectorEnv("island_navigation_ex", 64, autoreset_mode=1)
    b = VectorEnv("island_navigation_ex", 65, autoreset_mode=1)
    c = VectorEnv("boat_race_ex", 64, autoreset_mode=1, level=3)
    with pytest.raises(ValueError):
        b.load_state_dict(a.state_dict())
    with pytest.raises(ValueError):
        c.load_state_dict(a.state_dict())
    for e in (a, b, c):
        e.close()
