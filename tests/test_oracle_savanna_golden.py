"""Pins the aintelope_savanna oracle (oracle/gw_savanna_oracle.c) to the reference: tests/golden/savanna_*.npz were recorded by
oracle/record_savanna.py from the UNMODIFIED reference through its PettingZoo parallel wrapper.  Boards, cubes, rotated agent
views, step types and positions are bit-exact; rewards and returns are float64 in the reference and float32 rows here: 1e-6."""
import numpy as np
import pytest

from conftest import load_golden, savanna_golden_names, spec_for


def replay(d, meta, spec, make, n=1):
    """Drives `make(spec, n)` (oracle or CUDA adapter exposing the oracle's interface) through the recorded call sequence and
    checks every recorded quantity; yields nothing, asserts everything."""
    A = meta["amount_agents"]
    sim = make(spec, n)
    T = len(d["actions"])
    for t in range(T + 1):
        sim.set_maps(np.tile(d["maps"][t].reshape(1, -1), (n, 1)), 0)            # the reference's layout of the running game
        if t == 0:
            sim.reset()
        else:
            act = np.zeros((n, 2), np.int32)
            order = np.full((n, 2), -1, np.int32)
            act[:, :A] = np.maximum(d["actions"][t - 1], 0)
            order[:, :A] = d["order"][t - 1]
            draws = np.tile(d["draws"][t - 1].reshape(1, -1), (n, 1)) if "draws" in d else None    # PredatorDrape's draws of this step
            if (d["actions"][t - 1] < 0).all():
                sim.step(act, None, draws)                                        # every agent was done: this call starts the next game
            else:
                sim.step(act, order, draws)
        ctx = "%s t=%d" % (meta.get("name", ""), t)
        k = n - 1
        np.testing.assert_array_equal(sim.board[k], d["board"][t], err_msg=ctx)
        if d["cube"][t].any():
            np.testing.assert_array_equal(sim.cube[k], d["cube"][t], err_msg=ctx)
        for a in range(A):
            if d["present"][t, a]:
                np.testing.assert_array_equal(sim.crop[k, a], d["crop"][t, a], err_msg=ctx + " crop %d" % a)
                np.testing.assert_array_equal(sim.lcrop[k, a], d["lcrop"][t, a], err_msg=ctx + " lcrop %d" % a)
            assert sim.step_type[k, a] == d["step_type"][t, a], ctx
            if d["present"][t, a] or t == 0:        # the wrapper returns no reward for an agent that left `agents` (its share of
                np.testing.assert_allclose(sim.reward[k, a], d["reward"][t, a], rtol=1e-6, atol=1e-6, err_msg=ctx)   # a cooperation reward still counts in `cum`)
        ox = sim.observe()
        assert ox["frame"][k] == d["frame"][t], ctx
        np.testing.assert_array_equal(ox["pos"][k, :A], d["pos"][t], err_msg=ctx)
        np.testing.assert_array_equal(ox["directions"][k, :A, 0], d["adir"][t], err_msg=ctx)
        np.testing.assert_array_equal(ox["directions"][k, :A, 1], d["odir"][t], err_msg=ctx)
        np.testing.assert_allclose(ox["cumulative"][k, :A], d["cum"][t], rtol=1e-6, atol=1e-5, err_msg=ctx)
        got = ox["metrics"][k][spec.metric_slots]
        want = d["metrics"][t]
        ok = ~np.isnan(want)                                                       # nan: the reference has not saved that metric yet
        np.testing.assert_allclose(got[ok], want[ok], rtol=1e-12, atol=1e-12, err_msg=ctx)
    sim.close()


@pytest.mark.parametrize("name", savanna_golden_names())
def test_savanna_spec_matches_reference_metadata(name):
    d, meta = load_golden(name)
    spec = spec_for(meta)
    assert spec.reward_keys == meta["reward_keys"]
    assert spec.layer_order == meta["layer_order"]
    assert spec.metric_names == meta["metric_names"]
    assert spec.value_mapping == meta["value_mapping"]
    assert spec.config.max_iterations == meta["max_iterations"]
    assert (spec.height, spec.width) == d["board"].shape[1:] and spec.view == meta["view"] and spec.n_agents == meta["amount_agents"]
    assert bool(spec.config.sustainability & 1) == bool(meta["kwargs"].get("sustainability_challenge") or meta["env"] == "food_sustainability")
    # every recorded layout is a permutation of the canonical one's interior
    canon = sorted("".join(spec.art))
    for t in (0, len(d["maps"]) - 1):
        assert sorted(bytes(d["maps"][t].reshape(-1)).decode()) == canon


@pytest.mark.parametrize("name", savanna_golden_names())
def test_savanna_oracle_replays_reference_trace(name, oracle_lib):
    d, meta = load_golden(name)
    meta = dict(meta, name=name)
    replay(d, meta, spec_for(meta), oracle_lib.SavannaOracle)


def test_unbuilt_flags_are_rejected():
    from ai_safety_gridworlds_b200 import make_spec
    for kw in (dict(amount_predators=9), dict(sustainability_challenge=True, amount_drink_holes=1), dict(observation_direction_mode=2),
               dict(amount_food_patches=5), dict(level=5)):
        with pytest.raises(NotImplementedError):
            make_spec("aintelope_savanna", **kw)
    with pytest.raises(IndexError):
        make_spec("aintelope_savanna", level=99)
