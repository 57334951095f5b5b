"""Pins the CPU oracle (oracle/gw_oracle.c) to the reference.

Every trace under tests/golden/ was recorded by oracle/record.py from the UNMODIFIED reference
running in the build container.  The oracle replays the stored action arrays and must
reproduce the reference bit-exactly for integer / byte quantities and to 1e-9 relative for the
float64 reward dimensions (the oracle emits float32 rewards; comparison is against the float32
rounding of the reference value, which for these traces is exact unless flags are fractional).
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden, spec_for


@pytest.mark.parametrize("name", golden_names())
def test_spec_matches_reference_metadata(name):
    d, meta = load_golden(name)
    spec = spec_for(meta)
    assert spec.reward_keys == meta["reward_keys"]
    assert spec.layer_order == meta["layer_order"]
    assert spec.metric_names == meta["metric_names"]
    assert spec.art == meta["ascii_art"]
    assert spec.action_range == (meta["action_min"], meta["action_max"])
    assert spec.config.max_iterations == meta["max_iterations"]
    for ch, v in meta["value_mapping"].items():
        assert spec.value_mapping[ch] == v
    assert d["board"].shape[1:] == (spec.height, spec.width)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_replays_reference_trace(name, oracle_lib):
    d, meta = load_golden(name)
    spec = spec_for(meta)
    orc = oracle_lib.Oracle(spec, 1)
    orc.reset()
    T = len(d["actions"])
    integer_metric = [n.endswith("Visits") or n.endswith("Availability") for n in meta["metric_names"]]
    for t in range(T + 1):
        if t > 0:
            orc.step(np.array([d["actions"][t - 1]], np.int32))
        ctx = "%s t=%d" % (name, t)
        np.testing.assert_array_equal(orc.board[0], d["board"][t], err_msg=ctx)
        np.testing.assert_array_equal(orc.cube[0], d["cube"][t], err_msg=ctx)
        np.testing.assert_array_equal(orc.value_board[0], d["obs"][t], err_msg=ctx)
        assert orc.step_type[0] == d["step_type"][t], ctx
        assert orc.reason[0] == d["reason"][t], ctx
        assert orc.terminated[0] == (d["step_type"][t] == 2), ctx
        np.testing.assert_allclose(orc.reward[0], d["reward"][t], rtol=1e-6, atol=0, err_msg=ctx)
        ex = orc.observe()
        assert ex["frame"][0] == d["frame"][t], ctx
        np.testing.assert_array_equal(ex["pos"][0], d["pos"][t], err_msg=ctx)
        assert ex["safety"][0] == d["safety"][t], ctx
        np.testing.assert_allclose(ex["cumulative"][0], d["cumulative"][t], rtol=1e-6, atol=1e-6, err_msg=ctx)
        if meta["metric_names"]:
            got, want = ex["metrics"][0], d["metrics"][t]
            for j, is_int in enumerate(integer_metric):
                if is_int:
                    assert got[j] == want[j], (ctx, meta["metric_names"][j])
                else:
                    assert got[j] == pytest.approx(want[j], rel=1e-12, abs=1e-12), (ctx, meta["metric_names"][j])
    orc.close()


def test_discount_is_a_function_of_reason():
    """The engine does not emit `discount`; the wrapper derives it: 0.0 after a game-initiated
    termination or QUIT (pycolab/plot.py:176-199), 1.0 otherwise, None on FIRST."""
    for name in golden_names():
        d, _ = load_golden(name)
        st, reason, disc = d["step_type"], d["reason"], d["discount"]
        assert np.all(np.isnan(disc[st == 0]))
        ended = (st == 2) & ((reason == 0) | (reason == 3))
        assert np.all(disc[ended] == 0.0)
        other = (st != 0) & ~ended
        assert np.all(disc[other] == 1.0)
