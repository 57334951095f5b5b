"""Pins the firemaker_ex_ma CPU oracle to the reference: traces recorded by oracle/record_firemaker.py
through the reference's PettingZoo parallel wrapper, with the shuffle order and every FireDrape draw
replayed.  Boards, cubes, crops, step types, metrics: bit-exact; rewards are integer valued."""
import json

import numpy as np
import pytest

from conftest import firemaker_golden_names, load_golden


def fm_spec(meta, autoreset_mode=0):
    from ai_safety_gridworlds_b200 import make_spec
    kw = dict(meta["kwargs"])
    kw.setdefault("amount_agents", meta.get("amount_agents", 3))
    return make_spec("firemaker_ex_ma", autoreset_mode=autoreset_mode, **kw)


def replay_inputs(d, t, stride=1800):
    """actions / order / draws of trace step t (1-based index into the per-step arrays)."""
    lo, hi = int(d["draw_ofs"][t - 1]), int(d["draw_ofs"][t])
    draws = np.full((1, stride), 2.0)
    draws[0, :hi - lo] = d["draws"][lo:hi]
    return d["actions"][t - 1][None].astype(np.int32), d["order"][t - 1][None].astype(np.int32), draws


def check_against_trace(view, ox, d, meta, t, ctx):
    """`view`: object with board/cube/crop_w/... numpy arrays for environment 0."""
    np.testing.assert_array_equal(view["board"], d["board"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["cube"], d["cube"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["step_type"], d["step_type"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["terminated"], d["done"][t] if d["step_type"][t].max() >= 2 else np.zeros(3), err_msg=ctx)
    if d["step_type"][t].max() < 2 or True:
        np.testing.assert_array_equal(view["crop_w"][0], d["crop1"][t], err_msg=ctx)
        np.testing.assert_array_equal(view["crop_w"][1], d["crop2"][t], err_msg=ctx)
        np.testing.assert_array_equal(view["crop_s"], d["cropS"][t], err_msg=ctx)
        np.testing.assert_array_equal(view["lcrop_w"][0], d["lcrop1"][t], err_msg=ctx)
        np.testing.assert_array_equal(view["lcrop_w"][1], d["lcrop2"][t], err_msg=ctx)
        np.testing.assert_array_equal(view["lcrop_s"], d["lcropS"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["reward_w"][0], d["reward1"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["reward_w"][1], d["reward2"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["reward_s"], d["rewardS"][t], err_msg=ctx)
    assert ox["frame"] == d["frame"][t], ctx
    np.testing.assert_array_equal(ox["pos"], d["pos"][t], err_msg=ctx)
    if "dirs" in d and "directions" in ox:                               # traces of the direction modes carry the sprites' directions
        np.testing.assert_array_equal(ox["directions"], d["dirs"][t], err_msg=ctx)
    assert ox["ext_fires"] == d["ext_fires"][t], ctx
    from ai_safety_gridworlds_b200.envs.firemaker_ex_ma import METRIC_NAMES
    cols = [METRIC_NAMES.index(m) for m in meta["metric_names"]]         # amount_agents = 2: no columns of worker '2'
    np.testing.assert_array_equal(ox["metrics"][cols], d["metrics"][t], err_msg=ctx)
    np.testing.assert_array_equal(ox["cumulative"], np.concatenate([d["cum1"][t], d["cum2"][t], d["cumS"][t]]), err_msg=ctx)


@pytest.mark.parametrize("name", firemaker_golden_names())
def test_firemaker_oracle_replays_reference_trace(name, oracle_lib):
    d, meta = load_golden(name)
    spec = fm_spec(meta)
    assert spec.layer_order == meta["layer_order"] and spec.metric_names == meta["metric_names"]
    assert spec.reward_keys == meta["reward_keys"] and spec.value_mapping == meta["value_mapping"]
    assert spec.config.max_iterations == meta["max_iterations"]
    orc = oracle_lib.FiremakerOracle(spec, 1)
    T = len(d["actions"])
    for t in range(T + 1):
        if t == 0:
            orc.reset()
        else:
            a, o, dr = replay_inputs(d, t)
            orc.step(a, o, dr)
        ox = {k: v[0] for k, v in orc.observe().items()}
        view = dict(board=orc.board[0], cube=orc.cube[0], crop_w=orc.crop_w[0], crop_s=orc.crop_s[0], lcrop_w=orc.lcrop_w[0],
                    lcrop_s=orc.lcrop_s[0], reward_w=orc.reward_w[0], reward_s=orc.reward_s[0], step_type=orc.step_type[0],
                    terminated=orc.terminated[0])
        check_against_trace(view, ox, d, meta, t, "%s t=%d" % (name, t))
    orc.close()
