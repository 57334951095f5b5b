"""Pins the single-agent sub-step of the firemaker_ex_ma CPU oracle (execution order {agent, -1, -1}) to the
reference: traces recorded by oracle/record_firemaker_aec.py through the reference's AEC wrapper, where every
step is one `EnvironmentMa.step({agent: action})` = one engine frame.  Bit-exact."""
import numpy as np
import pytest

from conftest import firemaker_aec_golden_names, load_golden


def aec_inputs(d, k, stride=1800):
    """actions / order / draws of AEC step k (0-based); None for a dead step (no engine call)."""
    if d["action"][k] < 0:
        return None
    ag = int(d["agent"][k])
    act = np.zeros((1, 3), np.int32)
    act[0, ag] = d["action"][k]
    order = np.array([[ag, -1, -1]], np.int32)
    lo, hi = int(d["draw_ofs"][k]), int(d["draw_ofs"][k + 1])
    draws = np.full((1, stride), 2.0)
    draws[0, :hi - lo] = d["draws"][lo:hi]
    return act, order, draws


def check_engine_state(view, ox, d, t, ctx):
    np.testing.assert_array_equal(view["board"], d["board"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["cube"], d["cube"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["step_type"], d["step_type"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["crop_w"][0], d["crop1"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["crop_w"][1], d["crop2"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["crop_s"], d["cropS"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["lcrop_w"][0], d["lcrop1"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["lcrop_w"][1], d["lcrop2"][t], err_msg=ctx)
    np.testing.assert_array_equal(view["lcrop_s"], d["lcropS"][t], err_msg=ctx)
    assert ox["frame"] == d["frame"][t], ctx
    np.testing.assert_array_equal(ox["pos"], d["pos"][t], err_msg=ctx)


@pytest.mark.parametrize("name", firemaker_aec_golden_names())
def test_firemaker_oracle_replays_aec_trace(name, oracle_lib):
    from ai_safety_gridworlds_b200 import make_spec
    d, meta = load_golden(name)
    spec = make_spec("firemaker_ex_ma", autoreset_mode=0, amount_agents=3, **meta["kwargs"])
    orc = oracle_lib.FiremakerOracle(spec, 1)
    orc.reset()
    T = len(d["action"])
    fires = 0
    for t in range(T + 1):
        if t > 0:
            inp = aec_inputs(d, t - 1)
            if inp is not None:
                orc.step(*inp)
                # env.rewards after the step = this frame's reward of every agent that is still live in the wrapper
                term_before = d["term"][t - 1]
                for i, key in enumerate(("reward1", "reward2", "rewardS")):
                    if not term_before[i]:
                        got = orc.reward_s[0] if i == 2 else orc.reward_w[0][i]
                        np.testing.assert_array_equal(got, d[key][t], err_msg="%s t=%d %s" % (name, t, key))
        ox = {k: v[0] for k, v in orc.observe().items()}
        view = dict(board=orc.board[0], cube=orc.cube[0], crop_w=orc.crop_w[0], crop_s=orc.crop_s[0], lcrop_w=orc.lcrop_w[0],
                    lcrop_s=orc.lcrop_s[0], step_type=orc.step_type[0])
        check_engine_state(view, ox, d, t, "%s t=%d" % (name, t))
        fires += int((orc.board[0] == ord("F")).sum())
    if "maxiter" not in name:
        assert fires > 0
    orc.close()
