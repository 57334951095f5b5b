"""Pins the island_navigation_ex_ma CPU oracle to the reference: traces recorded by oracle/record_island_ma.py through
the reference's PettingZoo parallel wrapper, with the shuffle order replayed.  Boards, cubes, rotated agent views, step
types, positions, directions, integer metrics: bit-exact; float rewards / satiations <= 1e-9 relative."""
import numpy as np
import pytest

from conftest import island_ma_golden_names, load_golden


def ima_spec(meta, autoreset_mode=0):
    from ai_safety_gridworlds_b200 import make_spec
    return make_spec("island_navigation_ex_ma", autoreset_mode=autoreset_mode, **meta["kwargs"])


def check_against_trace(view, ox, spec, d, t, ctx):
    """`view`: board / cube / crop / lcrop / reward / step_type / terminated numpy arrays of environment 0."""
    np.testing.assert_array_equal(view["board"], d["board"][t], err_msg=ctx)
    present = d["present"][t]
    if present.any():                                              # the wrapper returns no global cube once every agent is done
        np.testing.assert_array_equal(view["cube"], d["cube"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["step_type"], d["step_type"][t], err_msg=ctx)
    for a, key in enumerate(("1", "2")):
        if present[a]:
            np.testing.assert_array_equal(view["crop"][a], d["crop" + key][t], err_msg=ctx + " crop" + key)
            np.testing.assert_array_equal(view["lcrop"][a], d["lcrop" + key][t], err_msg=ctx + " lcrop" + key)
            np.testing.assert_allclose(view["reward"][a], d["reward" + key][t], rtol=1e-6, atol=0, err_msg=ctx + " reward" + key)
            assert bool(view["terminated"][a]) == bool(d["done"][t][a]), ctx
    assert ox["frame"] == d["frame"][t], ctx
    np.testing.assert_array_equal(ox["pos"], d["pos"][t], err_msg=ctx)
    np.testing.assert_array_equal(ox["directions"][:, 0], d["adir"][t], err_msg=ctx)
    np.testing.assert_array_equal(ox["directions"][:, 1], d["odir"][t], err_msg=ctx)
    got = ox["metrics"][[spec.config.metric_slots[i] for i in range(spec.config.n_metrics)]]
    np.testing.assert_allclose(got, d["metrics"][t], rtol=1e-9, atol=0, err_msg=ctx + " metrics")
    np.testing.assert_allclose(ox["cumulative"], d["cum"][t], rtol=1e-6, atol=0, err_msg=ctx + " cumulative")


@pytest.mark.parametrize("name", island_ma_golden_names())
def test_island_ma_oracle_replays_reference_trace(name, oracle_lib):
    d, meta = load_golden(name)
    spec = ima_spec(meta)
    assert spec.layer_order == meta["layer_order"] and spec.metric_names == meta["metric_names"]
    assert spec.reward_keys == meta["reward_keys"]["1"] == meta["reward_keys"]["2"]
    assert spec.value_mapping == {k: v for k, v in meta["value_mapping"].items()} or set(meta["value_mapping"]) <= set(spec.value_mapping)
    assert spec.config.max_iterations == meta["max_iterations"]
    orc = oracle_lib.IslandMaOracle(spec, 1)
    T = len(d["actions"])
    maps = None
    if meta["kwargs"].get("map_randomization_frequency"):  # map randomisation: the reference's layouts are replayed, not re-derived
        assert len({m.tobytes() for m in d["maps"]}) > (1 if meta["kwargs"]["map_randomization_frequency"] == 3 else 0)
        maps = np.ascontiguousarray(d["maps"][:1]).copy()
        orc.set_maps(maps, 0)
    for t in range(T + 1):
        if maps is not None:
            maps[0] = d["maps"][t]                         # the layout of the game that runs after this call
        if t == 0:
            orc.reset()
        else:
            a = np.maximum(d["actions"][t - 1], 0)[None].astype(np.int32)          # -1 = absent agent (ignored) / recorder reset
            orc.step(a, d["order"][t - 1][None].astype(np.int32))
        ox = {k: v[0] for k, v in orc.observe().items()}
        view = dict(board=orc.board[0], cube=orc.cube[0], crop=orc.crop[0], lcrop=orc.lcrop[0], reward=orc.reward[0],
                    step_type=orc.step_type[0], terminated=orc.terminated[0])
        check_against_trace(view, ox, spec, d, t, "%s t=%d" % (name, t))
    orc.close()
