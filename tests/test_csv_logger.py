"""CSV logging parity: the logger must reproduce, byte for byte, the log files the UNMODIFIED reference wrote
(tests/golden/logs/*.csv, recorded by oracle/record_csv_log.py) -- header, column order, episode / iteration counters and the
10-significant-digit number format.  The CPU cases feed the logger from the golden traces; the GPU cases run the CUDA path
through GridworldGymEnv(log_columns=...) with the call sequence the recorder used."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden, spec_for

LOGS = os.path.join(GOLDEN_DIR, "logs")
CASES = sorted(os.path.basename(p)[:-len(".meta.json")] for p in glob.glob(os.path.join(LOGS, "*.meta.json")))


def _meta(name):
    with open(os.path.join(LOGS, name + ".meta.json")) as f:
        return json.load(f)


def _golden_text(name):
    with open(os.path.join(LOGS, name + ".csv"), newline="") as f:
        return f.read()


def test_format_float_matches_the_reference_number_format():
    from ai_safety_gridworlds_b200.helpers.csv_logger import format_float, widen_float32
    assert str(format_float(-1.7999999999999998)) == "-1.8"
    assert str(format_float(30.0)) == "30" and str(format_float(0.6666666666666666)) == "0.6666666667"
    assert str(format_float(np.int64(7))) == "7" and format_float(None) == "None"
    assert str(format_float(3547.2400000000002)) == "3547.24" and str(format_float(1e-12)) == "1E-12"   # str() uses the thread context, like the csv writer
    assert widen_float32(np.float32(-1.8)) == -1.8 and widen_float32(np.array([0.1, 2.5], np.float32)).tolist() == [0.1, 2.5]


@pytest.mark.parametrize("name", [c for c in CASES if not _meta(c)["explicit_resets"]])
def test_logger_reproduces_reference_csv_from_the_golden_trace(name, tmp_path):
    from ai_safety_gridworlds_b200.helpers import csv_logger
    lm = _meta(name)
    d, meta = load_golden(lm["trace"])
    spec = spec_for(meta)
    unit = csv_logger.reward_unit_space(spec.config.reward_table, spec.n_rewards)
    log = csv_logger.CsvLogger(csv_logger.reference_class(meta["env"]), lm["log_columns"], spec.reward_keys, spec.metric_names,
                               csv_logger.tile_types_of(spec.art), log_dir=str(tmp_path), log_filename_comment=lm["log_filename_comment"],
                               unit_space=unit)
    log.on_reset(state_is_first=False, state_is_none=True)            # the recorder's first reset(): nothing is opened yet
    assert log.file is None
    log.on_reset(state_is_first=True, state_is_none=False)            # its second reset() opens the file
    for t in range(1, lm["steps"] + 1):
        if d["frame"][t] > 0:                                         # the restart call after a terminal step logs nothing
            log.write_row(int(d["frame"][t]), d["reward"][t], d["cumulative"][t], None, d["metrics"][t] if d["metrics"].size else [])
    log.close()
    with open(os.path.join(str(tmp_path), log.log_filename), newline="") as f:
        assert f.read() == _golden_text(name)
    # the arguments file: the reward dimensions with their unit ranges and the metric keys
    with open(os.path.join(LOGS, name + ".arguments.txt")) as f:
        want = f.read()
    with open(os.path.join(str(tmp_path), log.arguments_filename)) as f:
        got = f.read()
    assert got[got.index("\t'reward_dimensions'"):] == want[want.index("\t'reward_dimensions'"):]
    assert os.path.basename(log.log_filename).startswith(lm["reference_filename"][:lm["reference_filename"].index("-golden-") + 8])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gym_wrapper_writes_the_reference_csv(name, tmp_path):
    pytest.importorskip("torch")
    from ai_safety_gridworlds_b200 import GridworldGymEnv
    lm = _meta(name)
    d, meta = load_golden(lm["trace"])
    env = GridworldGymEnv(meta["env"], seed=meta["seed"], log_columns=lm["log_columns"], log_dir=str(tmp_path),
                          log_filename_comment=lm["log_filename_comment"], **meta["kwargs"])
    env.reset()
    assert not os.listdir(str(tmp_path))                               # safety_game_mo.py:577-583: opened by the second reset()
    env.reset()
    for k, a in enumerate(d["actions"][:lm["steps"]]):
        for _ in range(lm["explicit_resets"].count(k)):
            env.reset()
        env.step(int(a))
    env.close()
    files = sorted(os.listdir(str(tmp_path)))
    assert len(files) == 2 and sorted(os.path.splitext(f)[1] for f in files) == [".csv", ".txt"]
    with open(os.path.join(str(tmp_path), [f for f in files if f.endswith(".csv")][0]), newline="") as f:
        assert f.read() == _golden_text(name)
