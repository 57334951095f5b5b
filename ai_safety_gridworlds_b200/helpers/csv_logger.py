"""CSV logging with the reference's schema.

Mirrors SafetyEnvironmentMo's logging (environments/shared/safety_game_mo.py): the column types
(:83-106), the header (`_write_log_header`, :727-807), one row per played step (`_write_log_row`, :1110-1215),
the number format (`format_float`, :1218-1227: decimal context of 10 significant digits, ROUND_HALF_UP,
trailing zeros dropped), ';' as the delimiter, the file name `<class>-<comment>-<timestamp>.csv` and the
separate arguments file (:586-625).  The episode / trial counters follow the reference's rules: an explicit
reset() after a played game starts the next episode (:697-705), the restart made inside step() after a
terminal timestep does not.

Host-side by nature (file IO); the numbers it prints come from the CUDA path: reward rows, the float64
episode return (GwExtras.cumulative_f64) and the metrics of gw_observe; the Gini / variance columns are derived
from those with the reference's own numpy expressions so that all 10 printed digits agree.
"""
import csv
import datetime
import decimal
import gzip
import itertools
import numbers
import os

import numpy as np

LOG_TIMESTAMP = "timestamp"
LOG_ENVIRONMENT = "env"
LOG_TRIAL = "trial"                       # obsolete alias for env layout seed
LOG_ENV_LAYOUT_SEED = "env layout seed"
LOG_ENV_SEED = "env seed"
LOG_EPISODE = "episode"
LOG_ITERATION = "iteration"
LOG_ARGUMENTS = "arguments"
LOG_REWARD_UNITS = "reward_unit"
LOG_REWARD = "reward"
LOG_SCALAR_REWARD = "scalar_reward"
LOG_CUMULATIVE_REWARD = "cumulative_reward"
LOG_AVERAGE_REWARD = "average_reward"
LOG_GINI_INDEX = "gini_index"
LOG_CUMULATIVE_GINI_INDEX = "cumulative_gini_index"
LOG_MO_VARIANCE = "mo_variance"
LOG_CUMULATIVE_MO_VARIANCE = "cumulative_mo_variance"
LOG_AVERAGE_MO_VARIANCE = "average_mo_variance"
LOG_SCALAR_CUMULATIVE_REWARD = "scalar_cumulative_reward"
LOG_SCALAR_AVERAGE_REWARD = "scalar_average_reward"
LOG_METRICS = "metric"
LOG_QVALUES_PER_TILETYPE = "tiletype_qvalue"

# the reference's class paths, written into the `env` column and the file name (safety_game_mo.py:589,1124)
REFERENCE_CLASS = {
    "island_navigation_ex": "ai_safety_gridworlds.environments.island_navigation_ex.IslandNavigationEnvironmentEx",
    "boat_race_ex": "ai_safety_gridworlds.environments.boat_race_ex.BoatRaceEnvironmentEx",
    "conveyor_belt_ex": "ai_safety_gridworlds.environments.conveyor_belt_ex.ConveyorBeltEnvironmentEx",
    "safe_interruptibility_ex": "ai_safety_gridworlds.environments.safe_interruptibility_ex.SafeInterruptibilityEnvironmentEx",
}



def reference_class(env_name):
    """Module path + class name of the reference environment behind a factory name (helpers/factory.py:100-201); the
    experiment overlays all subclass IslandNavigationEnvironmentEx under the same name in their own module."""
    name = env_name.lower()
    if name in REFERENCE_CLASS:
        return REFERENCE_CLASS[name]
    from ..envs import experiments
    if name in experiments.OVERLAYS:
        return "ai_safety_gridworlds.experiments.%s.IslandNavigationEnvironmentExExperiment" % name
    raise NotImplementedError("CSV logging: %r is not a SafetyEnvironmentMo game" % (env_name,))


_CTX = decimal.Context(prec=10, rounding=decimal.ROUND_HALF_UP, capitals=0)       # safety_game_mo.py:399-401


def format_float(value):
    """safety_game_mo.py:1218-1227"""
    if isinstance(value, numbers.Number):
        d = _CTX.create_decimal_from_float(float(value))
        integral = d.to_integral()
        return integral if d == integral else d.normalize()
    return str(value)


def widen_float32(x):
    """float32 -> the float64 its shortest round-trip decimal denotes.  The reward rows leave the kernel as float32; the
    reference prints Python floats with 10 significant digits, so -1.8f (-1.7999999523...) must read -1.8 again."""
    x = np.asarray(x, np.float32)
    return np.array([float(str(v)) for v in x.ravel()], np.float64).reshape(x.shape)


def gini_coefficient(dims):
    """safety_game_mo.py:1645-1681"""
    if len(dims) == 0:
        return np.float64(0.0)
    x = np.array(dims) - min(dims)
    mad = np.abs(np.subtract.outer(x, x)).mean()
    return 0.5 * (mad / (np.mean(x) + np.finfo(float).eps))


def tile_types_of(art, impassable="#", agent_chr="A", gap_chr=" "):
    """environment_data[TILE_TYPES] (AgentSafetySpriteMo.__init__, safety_game_mo.py:1326-1336)"""
    return sorted((set(itertools.chain.from_iterable(art)) - set(impassable) - set(agent_chr)) | set(gap_chr))


def reward_unit_space(reward_table, n_rewards):
    """mo_reward.get_enabled_reward_unit_space (mo_reward.py:149-181) from the spec's event table: per dimension the smallest
    and the largest unit value an enabled reward holds (0 where a reward lacks the dimension)."""
    rows = np.array([[reward_table[e][d] for d in range(n_rewards)] for e in range(len(reward_table))], np.float64)
    return np.minimum(rows.min(axis=0), 0.0), np.maximum(rows.max(axis=0), 0.0)


class CsvLogger(object):
    def __init__(self, classname, log_columns, reward_keys, metrics_keys, tile_types, log_dir="logs", log_filename_comment="",
                 log_arguments=None, flags=None, unit_space=None, log_arguments_to_separate_file=True, gzip_log=False,
                 env_layout_seed=1, env_seed=None, episode_no=None):
        self.classname = classname
        self.log_columns = list(log_columns or [])
        self.reward_keys = list(reward_keys)
        self.metrics_keys = list(metrics_keys)
        self.tile_types = list(tile_types)
        self.log_dir = log_dir
        self.log_filename_comment = log_filename_comment
        self.log_arguments = dict(log_arguments or {})
        self.flags = dict(flags or {})
        self.unit_space = unit_space
        self.log_arguments_to_separate_file = log_arguments_to_separate_file
        self.gzip_log = gzip_log
        self.env_layout_seed = env_layout_seed
        self.env_seed = env_layout_seed if env_seed is None else env_seed
        self.episode_no = 1 if episode_no is None else int(episode_no)
        self.create_new_log_file = True
        self.file = None
        self.log_filename = None
        self.arguments_filename = None

    # ------------------------------------------------------------------ reset() bookkeeping (safety_game_mo.py:526-705)
    def on_reset(self, state_is_first, state_is_none, env_layout_seed=None, start_new_experiment=False):
        """Call at the top of every explicit reset().  `state_is_first`: the environment stands at a FIRST timestep (no step
        played since the last reset); `state_is_none`: it was never reset (the constructor's hidden reset leaves _state None)."""
        if start_new_experiment:
            self.create_new_log_file = True
        if self.create_new_log_file and self.file is not None:
            self.close()
        if state_is_first and self.create_new_log_file:
            self.create_new_log_file = False
            if self.log_columns:
                self._open()
        if start_new_experiment or env_layout_seed is not None:
            if start_new_experiment and env_layout_seed is None:
                env_layout_seed = 1
            if (start_new_experiment or self.env_layout_seed != env_layout_seed
                    or (env_layout_seed == 1 and self.episode_no == 1 and (state_is_none or state_is_first))):
                self.env_layout_seed = env_layout_seed
                self.episode_no = 1
        elif not state_is_none and not state_is_first:
            self.episode_no += 1                              # only if the previous game was played

    def _open(self):
        if self.log_dir and not os.path.exists(self.log_dir):
            os.makedirs(self.log_dir)
        stamp = datetime.datetime.strftime(datetime.datetime.now(), "%Y.%m.%d-%H.%M.%S")
        sep = "-" if self.log_filename_comment else ""
        self.log_filename = self.classname + sep + self.log_filename_comment + "-" + stamp + ".csv"
        self.arguments_filename = self.classname + sep + self.log_filename_comment + "-arguments-" + stamp + ".txt"
        if self.log_arguments_to_separate_file:
            with open(os.path.join(self.log_dir, self.arguments_filename), mode="wt", encoding="utf-8") as f:
                self.write_arguments(f)
        if self.gzip_log:
            self.file = gzip.open(os.path.join(self.log_dir, self.log_filename + ".gz"), mode="wt", newline="", encoding="utf-8")
        else:
            self.file = open(os.path.join(self.log_dir, self.log_filename), mode="wt", buffering=1024 * 1024, newline="", encoding="utf-8")
        self.write_header(self.file)

    def write_arguments(self, f):
        """safety_game_mo.py:601-625"""
        print("{", file=f)
        for key, arg in self.log_arguments.items():
            print("\t'" + str(key) + "': " + str(arg) + ",", file=f)
        print("\t'FLAGS': {", file=f)
        for key, value in self.flags.items():
            print("\t\t'" + str(key) + "': " + str(value) + ",", file=f)
        print("\t},", file=f)
        print("\t'reward_dimensions': {", file=f)
        for index, key in enumerate(self.reward_keys):
            lo, hi = (self.unit_space[0][index], self.unit_space[1][index]) if self.unit_space is not None else (None, None)
            print("\t\t'" + str(key) + "': [" + str(lo) + ", " + str(hi) + "],", file=f)
        print("\t},", file=f)
        print("\t'metrics_keys': [", file=f)
        for key in self.metrics_keys:
            print("\t\t'" + str(key) + "',", file=f)
        print("\t],", file=f)
        print("}", file=f)
        f.flush()

    def close(self):
        if self.file is not None:
            self.file.flush()
            self.file.close()
            self.file = None

    # ------------------------------------------------------------------ header / rows
    def header(self):
        data = []
        for col in self.log_columns:
            if col in (LOG_TIMESTAMP, LOG_ENVIRONMENT, LOG_ENV_SEED, LOG_ENV_LAYOUT_SEED, LOG_TRIAL, LOG_EPISODE, LOG_ITERATION,
                       LOG_ARGUMENTS, LOG_SCALAR_REWARD, LOG_SCALAR_CUMULATIVE_REWARD, LOG_SCALAR_AVERAGE_REWARD, LOG_GINI_INDEX,
                       LOG_CUMULATIVE_GINI_INDEX, LOG_MO_VARIANCE, LOG_CUMULATIVE_MO_VARIANCE, LOG_AVERAGE_MO_VARIANCE):
                data.append(col)
            elif col in (LOG_REWARD, LOG_CUMULATIVE_REWARD, LOG_AVERAGE_REWARD):
                data += [col + "_" + k for k in self.reward_keys]
            elif col == LOG_METRICS:
                data += [LOG_METRICS + "_" + k for k in self.metrics_keys]
            elif col == LOG_QVALUES_PER_TILETYPE:
                data += [LOG_QVALUES_PER_TILETYPE + "_" + t.strip() + "_" + k for t in self.tile_types for k in self.reward_keys]
        return data

    def write_header(self, f):
        csv.writer(f, quoting=csv.QUOTE_MINIMAL, delimiter=";").writerow(self.header())
        f.flush()

    def row(self, iteration, reward, cumulative, scalars_cumulative, metrics, q_value_per_tiletype=None):
        """One log row.  reward / cumulative: float64 vectors over the reward dimensions; scalars_cumulative:
        (cumulative_gini_index, cumulative_mo_variance, average_mo_variance) as computed on the device, or None to compute
        them here; metrics: values in metrics_keys order.  Everything else is derived the way _process_timestep derives it
        (safety_game_mo.py:1027-1084)."""
        reward = [float(x) for x in reward]
        cumulative = [float(x) for x in cumulative]
        average = [x / (iteration + 1) for x in cumulative]
        gini, var = gini_coefficient(reward) * 100, np.var(reward, ddof=0)
        if scalars_cumulative is None:
            scalars_cumulative = (gini_coefficient(cumulative) * 100, np.var(cumulative, ddof=0), np.var(average, ddof=0))
        cgini, cvar, avar = scalars_cumulative
        data = []
        for col in self.log_columns:
            if col == LOG_TIMESTAMP:
                data.append(datetime.datetime.strftime(datetime.datetime.now(), "%Y.%m.%d-%H.%M.%S"))
            elif col == LOG_ENVIRONMENT:
                data.append(self.classname)
            elif col == LOG_ENV_SEED:
                data.append(self.env_seed)
            elif col in (LOG_ENV_LAYOUT_SEED, LOG_TRIAL):
                data.append(self.env_layout_seed)
            elif col == LOG_EPISODE:
                data.append(self.episode_no)
            elif col == LOG_ITERATION:
                data.append(iteration)
            elif col == LOG_ARGUMENTS:
                data.append(str(self.log_arguments))
            elif col == LOG_REWARD:
                data += [format_float(v) for v in reward]
            elif col == LOG_SCALAR_REWARD:
                data.append(format_float(sum(reward)))
            elif col == LOG_CUMULATIVE_REWARD:
                data += [format_float(v) for v in cumulative]
            elif col == LOG_AVERAGE_REWARD:
                data += [format_float(v) for v in average]
            elif col == LOG_SCALAR_CUMULATIVE_REWARD:
                data.append(format_float(sum(cumulative)))
            elif col == LOG_SCALAR_AVERAGE_REWARD:
                data.append(format_float(sum(average)))
            elif col == LOG_GINI_INDEX:
                data.append(format_float(gini))
            elif col == LOG_CUMULATIVE_GINI_INDEX:
                data.append(format_float(cgini))
            elif col == LOG_MO_VARIANCE:
                data.append(format_float(var))
            elif col == LOG_CUMULATIVE_MO_VARIANCE:
                data.append(format_float(cvar))
            elif col == LOG_AVERAGE_MO_VARIANCE:
                data.append(format_float(avar))
            elif col == LOG_METRICS:
                data += [format_float(v) for v in metrics]
            elif col == LOG_QVALUES_PER_TILETYPE:
                q = q_value_per_tiletype or {}
                data += [format_float(v) for t in self.tile_types for v in q.get(t, np.zeros([len(reward)]))]
        return data

    def write_row(self, *args, **kwargs):
        if self.file is None:
            return
        csv.writer(self.file, quoting=csv.QUOTE_MINIMAL, delimiter=";").writerow(self.row(*args, **kwargs))
        self.file.flush()
