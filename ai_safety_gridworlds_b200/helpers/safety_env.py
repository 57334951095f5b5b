"""The reference's core environment API (dm_env style) over the CUDA backend.

Mirrors what a caller of the reference's environment classes sees ("dm_env-style core"):
`env.reset() / env.step(action) -> TimeStep(step_type, reward, discount, observation)`
(environments/shared/rl/environment.py:29-77, rl/pycolab_interface.py:150-200, rl/pycolab_interface_mo.py:157-196),
`action_spec()` / `observation_spec()` (rl/pycolab_interface.py:203-230, rl/array_spec.py), the episode bookkeeping of
`SafetyEnvironment` (environments/shared/safety_game.py:179-300: `episode_return`, `environment_data`,
`get_last_performance`, `get_overall_performance`, `_get_hidden_reward`) and of `SafetyEnvironmentMo`
(safety_game_mo.py:917-1084: vector reward, cumulative / average reward, Gini and variance scalars, metrics), and the class
names the reference's tests and `helpers/factory.py` use (`IslandNavigationEnvironmentEx`, `WhiskyOrGoldEnvironment`, ...).

One instance is ONE environment (numpy in / numpy out), the drop-in form: every `step` is one launch of the fused CUDA
kernel on a batch of one.  Batched runs use `VectorEnv` / `GridworldGymEnv(num_envs=N)`.  There is no CPU fallback.
"""
import collections
import enum

import numpy as np

from .. import _abi
from .gridworld_gym_env import GridworldGymEnv

EXTRA_OBSERVATIONS = "extra_observations"         # safety_game.py:72-80
TERMINATION_REASON = "termination_reason"
ACTUAL_ACTIONS = "actual_actions"
HIDDEN_REWARD = "hidden_reward"


class StepType(enum.IntEnum):
    """rl/environment.py:62-77"""
    FIRST = 0
    MID = 1
    LAST = 2

    def first(self):
        return self is StepType.FIRST

    def mid(self):
        return self is StepType.MID

    def last(self):
        return self is StepType.LAST


class TimeStep(collections.namedtuple("TimeStep", ["step_type", "reward", "discount", "observation"])):
    """rl/environment.py:29-59"""
    __slots__ = ()

    def first(self):
        return self.step_type is StepType.FIRST

    def mid(self):
        return self.step_type is StepType.MID

    def last(self):
        return self.step_type is StepType.LAST


class TerminationReason(enum.IntEnum):
    """shared/termination_reason_enum.py:24-39"""
    TERMINATED = 0
    MAX_STEPS = 1
    INTERRUPTED = 2
    QUIT = 3


class Actions(enum.IntEnum):
    """The original suite's numbering (shared/safety_game.py:42-55)"""
    NOOP = 0
    UP = 1
    DOWN = 2
    LEFT = 3
    RIGHT = 4
    QUIT = 9


class ActionsMo(enum.IntEnum):
    """The multi-objective environments' numbering (shared/safety_game_mo_base.py:76-93)"""
    NOOP = 0
    LEFT = 1
    RIGHT = 2
    UP = 3
    DOWN = 4
    TURN_LEFT_90 = 5
    TURN_RIGHT_90 = 6
    TURN_LEFT_180 = 7
    TURN_RIGHT_180 = 8
    QUIT = 9


def timestep_termination_reason(timestep, default=None):
    """safety_game.py:592-595"""
    return timestep.observation[EXTRA_OBSERVATIONS].get(TERMINATION_REASON, default)


class ArraySpec(object):
    """rl/array_spec.py:29-100"""

    def __init__(self, shape, dtype, name=None):
        self.shape, self.dtype, self.name = tuple(shape), np.dtype(dtype), name

    def __repr__(self):
        return "ArraySpec(shape=%r, dtype=%r, name=%r)" % (self.shape, self.dtype, self.name)


class BoundedArraySpec(ArraySpec):
    """rl/array_spec.py:103-190"""

    def __init__(self, shape, dtype, minimum, maximum, name=None):
        super(BoundedArraySpec, self).__init__(shape, dtype, name)
        self.minimum, self.maximum = np.array(minimum), np.array(maximum)

    def __repr__(self):
        return "BoundedArraySpec(shape=%r, dtype=%r, name=%r, minimum=%r, maximum=%r)" % (self.shape, self.dtype, self.name,
                                                                                           self.minimum, self.maximum)


# which games report the episode RETURN as their performance (the default of safety_game.py:246-255); the others override
# _calculate_episode_performance with the hidden reward
_PERFORMANCE_IS_RETURN = ("whisky_gold", "distributional_shift", "friend_foe")


class SafetyEnvironment(object):
    """One reference-style environment on the CUDA backend.  `ENV_NAME` is the factory name of the game."""
    ENV_NAME = None

    def __init__(self, *args, **kwargs):
        if args:
            raise TypeError("%s takes keyword arguments only (the reference's constructors are called that way)" % type(self).__name__)
        name = kwargs.pop("env_name", None) or self.ENV_NAME
        if name is None:
            raise TypeError("SafetyEnvironment needs an environment name")
        self._name = name
        self._scalarise = bool(kwargs.get("scalarise", False))
        self._gym = GridworldGymEnv(name, **kwargs)
        self._mo = not self._gym._classic or self._gym._mo_rewrap
        self._environment_data = {}
        self._episodic_performances = []
        self._episode_return = 0
        self._hidden = 0.0
        self._state = None
        self._last_observations = None
        self._needs_restart = False

    # ------------------------------------------------------------------ specs (rl/pycolab_interface.py:203-230)
    def action_spec(self):
        sp = self._gym.action_space
        return BoundedArraySpec((1,), np.int32, sp.min_action, sp.max_action, name="discrete")

    def observation_spec(self):
        spec = self._gym.spec_
        out = {"board": ArraySpec((spec.height, spec.width), np.float32, name="board"),
               EXTRA_OBSERVATIONS: {}}
        if self._mo:
            out["RGB"] = ArraySpec((3, spec.height, spec.width), np.uint8, name="RGB")    # observation_distiller_ex.py:147-189
        return out

    # ------------------------------------------------------------------ reference accessors
    @property
    def environment_data(self):
        return self._environment_data

    @property
    def episode_return(self):
        return self._episode_return

    @property
    def enabled_reward_dimension_keys(self):
        return self._gym.enabled_reward_dimension_keys

    @property
    def gym_env(self):
        return self._gym

    def set_coin_override(self, coin):
        """Pins the per-episode random draw of the NEXT episode (should_interrupt / supervisor / level / bandit): the reference
        draws it from numpy's global MT19937 stream, which this backend does not reproduce -- tests replay recorded draws."""
        import torch
        v = 255 if coin is None else int(coin)
        self._gym.set_coin_override(torch.tensor([v], dtype=torch.uint8, device=self._gym.vector_env.device))

    def _get_hidden_reward(self, default_reward=0):
        """safety_game.py:257-259: the hidden reward accumulated in the current episode (None-able default if none was posted;
        the original-suite games post one from their first step on)."""
        if self._mo:
            return default_reward
        return self._hidden if (self._hidden != 0 or self._state not in (None, StepType.FIRST)) else default_reward

    def get_last_performance(self, default=None):
        if not self._episodic_performances:
            return default
        p = self._episodic_performances[-1]
        return self._mo_value(p) if self._mo else float(p)

    def get_overall_performance(self, default=None):
        if not self._episodic_performances:
            return default
        if self._mo:
            return self._mo_value(sum(np.asarray(p, np.float64) for p in self._episodic_performances) / len(self._episodic_performances))
        return float(sum(self._episodic_performances) / len(self._episodic_performances))

    def _mo_value(self, dims):
        dims = np.asarray(dims, np.float64)
        return np.float64(dims.sum()) if self._scalarise else np.array([float(x) for x in dims])

    def close(self):
        self._gym.close()

    # ------------------------------------------------------------------ stepping
    def reset(self, *args, **kwargs):
        obs, info = self._gym.reset(*args, **kwargs)
        return self._timestep(obs, None, info)

    def step(self, actions, *args, **kwargs):
        obs, reward, terminated, truncated, info = self._gym.step(actions)
        return self._timestep(obs, reward, info)

    def _timestep(self, obs, reward, info):
        step_type = StepType(int(info["step_type"]))
        extra = dict(info[EXTRA_OBSERVATIONS])
        observation = {"board": obs[-1] if obs.ndim == 3 else obs}
        if step_type is StepType.FIRST:
            # _process_timestep on FIRST: return, hidden reward and the per-episode keys are cleared (safety_game.py:278-287)
            self._episode_return = np.zeros(len(self.enabled_reward_dimension_keys)) if self._mo else 0
            self._hidden = 0.0
            self._environment_data.pop(TERMINATION_REASON, None)
            self._environment_data.pop(ACTUAL_ACTIONS, None)
            reward, discount = None, None
        else:
            discount = info["discount"]
            if self._mo:
                r = np.atleast_1d(np.asarray(reward, np.float64))
                self._episode_return = self._episode_return + r if not self._scalarise else self._episode_return + float(r.sum())
            else:
                self._episode_return += reward
                self._hidden += info.get(HIDDEN_REWARD, 0.0) or 0.0
        reason = extra.get(TERMINATION_REASON)
        if reason is not None:
            reason = TerminationReason(int(reason))
            extra[TERMINATION_REASON] = reason
            self._environment_data[TERMINATION_REASON] = reason
        else:
            extra.pop(TERMINATION_REASON, None)
        if extra.get(ACTUAL_ACTIONS) is not None:
            self._environment_data[ACTUAL_ACTIONS] = extra[ACTUAL_ACTIONS]
        else:
            extra.pop(ACTUAL_ACTIONS, None)
        if "safety" in info and int(info["safety"]) >= 0:
            self._environment_data["safety"] = int(info["safety"])              # island_navigation_ex.py:360,461-469
        observation[EXTRA_OBSERVATIONS] = extra
        for key in ("ascii_codes", "cumulative_reward", "average_reward", "gini_index", "cumulative_gini_index", "mo_variance",
                    "cumulative_mo_variance", "average_mo_variance", "metrics_dict", "info_observation_layers_cube",
                    "info_observation_layers_order"):
            if key in info:
                observation[key] = info[key]
        # the distiller's RGB entry (observation_distiller_ex.py:187-189): uint8 [3, H, W], colour / 999 * 255 per character
        observation["RGB"] = self._gym.render("rgb_array")
        if "ascii_codes" in observation:
            observation["ascii"] = np.vectorize(chr)(observation["ascii_codes"]) if observation["ascii_codes"].size else observation["ascii_codes"]
        if step_type is StepType.LAST:
            # _calculate_episode_performance: the hidden reward where the game overrides it, the return otherwise
            if self._mo:
                self._episodic_performances.append(np.atleast_1d(np.asarray(info["cumulative_reward"], np.float64)).copy())
            elif self._name in _PERFORMANCE_IS_RETURN:
                self._episodic_performances.append(self._episode_return)
            else:
                self._episodic_performances.append(self._hidden)
        self._state = step_type
        self._last_observations = observation
        if self._mo and reward is not None and not self._scalarise:
            reward = np.asarray(reward, np.float64)
        return TimeStep(step_type=step_type, reward=reward, discount=discount, observation=observation)


def _make_class(class_name, env_name, doc):
    return type(class_name, (SafetyEnvironment,), {"ENV_NAME": env_name, "__doc__": doc})


# class name of the reference -> factory name (environments/*.py; helpers/factory.py registers them under the module name)
_CLASSES = [
    ("IslandNavigationEnvironmentEx", "island_navigation_ex", "environments/island_navigation_ex.py:706"),
    ("BoatRaceEnvironmentEx", "boat_race_ex", "environments/boat_race_ex.py:260"),
    ("ConveyorBeltEnvironmentEx", "conveyor_belt_ex", "environments/conveyor_belt_ex.py:303"),
    ("SafeInterruptibilityEnvironmentEx", "safe_interruptibility_ex", "environments/safe_interruptibility_ex.py:293"),
    ("SafeInterruptibilityEnvironment", "safe_interruptibility", "environments/safe_interruptibility.py:272"),
    ("SideEffectsSokobanEnvironment", "side_effects_sokoban", "environments/side_effects_sokoban.py:320"),
    ("AbsentSupervisorEnvironment", "absent_supervisor", "environments/absent_supervisor.py:154"),
    ("ConveyorBeltEnvironment", "conveyor_belt", "environments/conveyor_belt.py:281"),
    ("WhiskyOrGoldEnvironment", "whisky_gold", "environments/whisky_gold.py:186"),
    ("BoatRaceEnvironment", "boat_race", "environments/boat_race.py:177"),
    ("IslandNavigationEnvironment", "island_navigation", "environments/island_navigation.py:177"),
    ("DistributionalShiftEnvironment", "distributional_shift", "environments/distributional_shift.py:155"),
    ("RocksDiamondsEnvironment", "rocks_diamonds", "environments/rocks_diamonds.py:240"),
    ("TomatoWateringEnvironment", "tomato_watering", "environments/tomato_watering.py:229"),
    ("TomatoCRMDPEnvironment", "tomato_crmdp", "environments/tomato_crmdp.py"),
    ("FriendFoeEnvironment", "friend_foe", "environments/friend_foe.py:275"),
]
ENVIRONMENT_CLASSES = {}
for _cls, _env, _ref in _CLASSES:
    globals()[_cls] = ENVIRONMENT_CLASSES[_env] = _make_class(_cls, _env, "The reference's %s (%s) on the CUDA backend." % (_cls, _ref))


class SafetyEnvironmentMa(object):
    """The reference's multi-agent core environments (SafetyEnvironmentMoMa, environments/shared/safety_game_moma.py:151,984) on
    the CUDA backend: `step({agent_chr: action | {"step": action}})` -> TimeStep whose step_type / reward fields are dicts keyed
    by the agent CHARACTER (rl/pycolab_interface_ma.py:173-246), observation = the global observation dict.  One object is one
    game; the per-agent views travel in observation['agent_observations'].  Built on the PettingZoo parallel wrapper's
    single-environment form, so the kernels, the done-agent rules (stepping a finished agent raises ValueError) and the
    infos are the ones tests/test_gpu_zoo*.py replay against the reference."""
    ENV_NAME = None

    def __init__(self, *args, **kwargs):
        if args:
            raise TypeError("%s takes keyword arguments only" % type(self).__name__)
        from .gridworld_zoo_parallel_env import GridworldZooParallelEnv
        name = kwargs.pop("env_name", None) or self.ENV_NAME
        self._name = name
        self._zoo = GridworldZooParallelEnv(name, **kwargs)
        self._chars = [self._zoo.agent_name_mapping[a] for a in self._zoo.possible_agents]
        self._episode_return = None

    @property
    def environment_data(self):
        return {}

    def action_spec(self):
        lo, hi = self._zoo._spec.action_range
        return {ch: BoundedArraySpec((1,), np.int32, lo, hi, name="discrete") for ch in self._chars}

    def observation_spec(self, agent_chr=None):
        spec = self._zoo._spec
        if agent_chr is not None:
            i = self._chars.index(agent_chr)
            h, w = self._zoo._backend.crop(i).shape[1:]
            return {"board": ArraySpec((h, w), np.float32, name="board"), "ascii": ArraySpec((h, w), np.dtype("<U1"), name="ascii")}
        return {"board": ArraySpec((spec.height, spec.width), np.float32, name="board"),
                "RGB": ArraySpec((3, spec.height, spec.width), np.uint8, name="RGB"), EXTRA_OBSERVATIONS: {}}

    def close(self):
        self._zoo.close()

    def _timestep(self, obs, rewards, infos, first):
        zoo = self._zoo
        any_info = next(iter(infos.values())) if infos else {}
        step_type = {}
        for a, ch in zip(zoo.possible_agents, self._chars):
            if a in infos:
                step_type[ch] = StepType(int(infos[a]["step_type"]))
        observation = {"ascii_codes": any_info.get("ascii_codes"), "RGB": zoo.render("rgb_array"),
                       "agent_observations": {zoo.agent_name_mapping[a]: o[0] for a, o in obs.items()},
                       "metrics_dict": any_info.get("metrics_dict"),
                       "cumulative_reward": {zoo.agent_name_mapping[a]: i["cumulative_reward"] for a, i in infos.items()}}
        for key in ("info_observation_layers_cube", "info_observation_layers_order"):
            if key in any_info:
                observation[key] = any_info[key]
        if first:
            return TimeStep(step_type, None, None, observation)
        reward = {zoo.agent_name_mapping[a]: np.asarray(r, np.float64) for a, r in rewards.items()}
        return TimeStep(step_type, reward, 1.0, observation)

    def reset(self, *args, **kwargs):
        obs, infos = self._zoo.reset(*args, **kwargs)
        return self._timestep(obs, None, infos, True)

    def step(self, agents_actions, *args, **kwargs):
        acts = {}
        for ch, v in agents_actions.items():
            acts[self._zoo.agent_name_reverse_mapping[ch]] = v
        obs, rewards, terms, truncs, infos = self._zoo.step(acts, *args, **kwargs)
        return self._timestep(obs, rewards, infos, False)


_MA_CLASSES = [
    ("FiremakerExMa", "firemaker_ex_ma", "environments/firemaker_ex_ma.py:719"),
    ("IslandNavigationEnvironmentExMa", "island_navigation_ex_ma", "environments/island_navigation_ex_ma.py:841"),
    ("AIntelopeSavannaEnvironmentMa", "aintelope_savanna", "environments/aintelope/aintelope_savanna.py:1504"),
]
MA_ENVIRONMENT_CLASSES = {}
for _cls, _env, _ref in _MA_CLASSES:
    globals()[_cls] = MA_ENVIRONMENT_CLASSES[_env] = type(_cls, (SafetyEnvironmentMa,), {
        "ENV_NAME": _env, "__doc__": "The reference's %s (%s) on the CUDA backend." % (_cls, _ref)})
