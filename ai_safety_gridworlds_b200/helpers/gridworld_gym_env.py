"""GridworldGymEnv: the reference's Gym/Gymnasium wrapper signature over the CUDA vector backend.

Mirrors helpers/gridworld_gym_env.py of the reference (constructor :99-133, `step` :455-585,
`reset` :588-674, info keys :54-64,397-450, action / observation spaces :753-981):

  * `num_envs=None` (default) is the drop-in single environment: numpy outputs with the
    reference's shapes and dtypes -- obs float32 [1,H,W] (or [2,H,W] with `use_transitions`),
    reward float64 [R], `terminated` bool, `truncated` False, info dict -- and the reference's
    stepping semantics (the call after a terminal step rebuilds the game, ignores its action and
    returns the first observation, rl/pycolab_interface_mo.py:175-178);
  * `num_envs=N` is the batched form: the same quantities as torch CUDA tensors with a leading
    batch dimension and auto-reset inside the terminal step (the observation returned with
    `terminated` is the new episode's first frame);
  * conveyor_belt_ex and safe_interruptibility_ex, the SafetyEnvironmentMo re-wrappings of two original-suite games, run on the
    classic kernel too but keep the multi-objective conventions: reward float64 [1] (dimension 'REWARD'), cumulative / average
    reward, Gini / variance scalars (0 for one dimension) and the un-occluded layers cube in info;
  * the original-suite games (safe_interruptibility, side_effects_sokoban, absent_supervisor,
    conveyor_belt, whisky_gold, boat_race, island_navigation, distributional_shift, rocks_diamonds,
    tomato_watering, tomato_crmdp) run on the classic kernel: obs
    float32 [1,H,W], SCALAR reward, and `info['hidden_reward']` = this step's hidden reward
    (gridworld_gym_env.py:497-506), the way the reference's wrapper serves them
    (tests/gridworld_gym_env_test.py:63-110 replays the demonstrations through it).

gymnasium itself is not imported: the class is duck-typed, and the two space classes below carry
what the reference's spaces expose.  Every step is one launch of the fused CUDA kernel; there is
no CPU fallback.
"""
import numpy as np
import torch

from .. import _abi
from ..envs import make_spec
from ..envs.classic import CLASSIC_ENV_TYPES, SokSpec
from ..classic_env import crop_board as _crop_classic


def crop_board(rows, spec):
    """The [..., H, W] board out of a padded classic board row (64 entries, [8, 8]) or a dense sokoban row (128 entries)."""
    if isinstance(spec, SokSpec):
        from ..sokoban_env import crop_rows
        return crop_rows(rows, spec)
    return _crop_classic(rows, spec)
from ..vector_env import VectorEnv

INFO_OBSERVED_REWARD = "observed_reward"
INFO_HIDDEN_REWARD = "hidden_reward"
INFO_DISCOUNT = "discount"
INFO_OBSERVATION_COORDINATES = "info_observation_coordinates"
INFO_OBSERVATION_LAYERS_DICT = "info_observation_layers_dict"
INFO_OBSERVATION_LAYERS_ORDER = "info_observation_layers_order"
INFO_OBSERVATION_LAYERS_CUBE = "info_observation_layers_cube"


class DiscreteActionSpace(object):
    """gym.spaces.Discrete(n, start=min_action) as the reference builds it (gridworld_gym_env.py:885-896)."""

    def __init__(self, min_action, max_action, seed=None):
        self.min_action, self.max_action = int(min_action), int(max_action)
        self.n, self.start = self.max_action - self.min_action + 1, self.min_action
        self.shape, self.dtype = (), np.int64
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return int(self._rng.integers(self.min_action, self.max_action + 1))

    def contains(self, x):
        try:
            return self.min_action <= int(x) <= self.max_action and int(x) == x
        except (TypeError, ValueError):
            return False

    __contains__ = contains


class BoardObservationSpace(object):
    """The Box-like observation space of the value-mapped board (gridworld_gym_env.py:900-981)."""

    def __init__(self, shape, low, high):
        self.shape, self.dtype, self.low, self.high = tuple(shape), np.float32, float(low), float(high)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    __contains__ = contains


class GridworldGymEnv(object):
    metadata = {"render.modes": ["ansi", "rgb_array"]}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, env_name, use_transitions=False, flatten_observations=False, object_coordinates_in_observation=True,
                 layers_in_observation=True, occlusion_in_layers=False, layers_order_in_cube=[], seed=None, num_envs=None,
                 device=None, scalarise=False, log_columns=None, log_dir="logs", log_filename_comment="", log_arguments=None,
                 log_arguments_to_separate_file=True, gzip_log=False, env_layout_seed=1, trial_no=None, episode_no=None,
                 log_env_index=0, **kwargs):
        # the reference wrapper's own constructor arguments (gridworld_gym_env.py:99-133) are honoured or refused, never dropped
        for k in ("ascii_attributes_format", "layers_in_attribute_observation", "occlusion_in_atribute_layers",
                  "use_multi_discrete_action_space"):
            if kwargs.pop(k, False):
                raise NotImplementedError("%s is not built (agent attributes / multi-discrete actions)" % k)
        if kwargs.pop("observable_attribute_value_mapping", None):
            raise NotImplementedError("observable_attribute_value_mapping is not built")
        if kwargs.pop("agent_character", None) is not None:
            raise NotImplementedError("agent_character: the Gym wrapper of the CUDA backend serves the single-agent games")
        self.render_mode = kwargs.pop("render_mode", None)
        self._callbacks = {k: kwargs.pop(k, None) for k in ("pre_reset_callback", "post_reset_callback", "pre_step_callback", "post_step_callback")}
        for k in ("render_animation_delay", "ascii_observation_format", "attribute_coordinates_in_observation",
                  "observable_attribute_categories", "np_random"):
            kwargs.pop(k, None)     # no effect on these games: no viewer; MO observations are the float board (:191); no attributes
        # CSV logging of SafetyEnvironmentMo (safety_game_mo.py:727-807,1110-1215): one row per played step of environment
        # `log_env_index` (the batch has one log, like the reference's class-level file handle)
        self._logger, self._log_env = None, int(log_env_index)
        self._log_setup = dict(log_columns=log_columns, log_dir=log_dir, log_filename_comment=log_filename_comment,
                               log_arguments=log_arguments if log_arguments is not None else dict(seed=seed, **kwargs),
                               log_arguments_to_separate_file=log_arguments_to_separate_file, gzip_log=gzip_log,
                               env_layout_seed=env_layout_seed if trial_no is None else trial_no, episode_no=episode_no, env_seed=seed)
        self._played = None                                         # None: never reset; False: at a FIRST timestep; True: played
        # SafetyEnvironmentMo(scalarise=True): reward, cumulative_reward and average_reward are the SUM over the reward
        # dimensions, as np.float64 (safety_game_mo.py:1028-1064)
        self._scalarise = bool(scalarise)
        if occlusion_in_layers:
            raise NotImplementedError("occlusion_in_layers=True: the MO environments force it off (safety_game_mo_base.py)")
        self._batched = num_envs is not None
        n = int(num_envs) if self._batched else 1
        # Batched form: a finished environment restarts inside the wrapper's step() -- but AFTER the step's info has been read, so
        # that `info` (cumulative / average reward, metrics, frame, termination reason) describes the episode that just ENDED, the
        # way `reward` and `terminated` do, while the returned observation already is the first one of the next episode (the
        # Gymnasium vector convention; the ended episode's last observation travels as info['final_observation']).  The kernel
        # therefore runs with the reference's own reset semantics and the wrapper issues the masked gw_reset itself.
        mode = _abi.GW_AUTORESET_NEXT_STEP
        self._spec = make_spec(env_name, autoreset_mode=mode, **kwargs)      # raises NotImplementedError like factory.py:199-201
        self._flatten = bool(flatten_observations)
        self._sokoban_big = isinstance(self._spec, SokSpec)              # side_effects_sokoban levels 1-3: the gw_sok_* path
        self._classic = self._sokoban_big or self._spec.config.env_type in CLASSIC_ENV_TYPES
        if self._classic:
            self._init_classic(env_name, n, device, mode, seed, use_transitions, bool(object_coordinates_in_observation),
                               bool(layers_in_observation), layers_order_in_cube)
            return
        self._env = VectorEnv(self._spec, n, device=device, autoreset_mode=mode)
        self._env_name = env_name
        self._use_transitions = bool(use_transitions)
        self._object_coordinates = bool(object_coordinates_in_observation)
        self._layers_in_observation = bool(layers_in_observation)
        order = list(layers_order_in_cube) if layers_order_in_cube else list(self._spec.layer_order)
        # a name the game does not have gives an all-zero plane ("cross-environment observation format compatibility",
        # safety_game_mo.py:497-503)
        self._layers_order = order
        self._layer_index = torch.tensor([self._spec.layer_order.index(ch) if ch in self._spec.layer_order else -1 for ch in order],
                                         dtype=torch.long, device=self._env.device)
        lo, hi = self._spec.action_range
        self.action_space = DiscreteActionSpace(lo, hi, seed)
        vals = list(self._spec.value_mapping.values())
        depth = 2 if self._use_transitions else 1
        self.observation_space = BoardObservationSpace((depth, self._spec.height, self._spec.width), min(vals), max(vals))
        self.num_envs = n
        self._last_board = None
        self._last_hidden = None
        self._seed = seed
        self._init_logger(env_name)

    def _init_classic(self, env_name, n, device, mode, seed, use_transitions, object_coordinates=False, layers_in_observation=False,
                      layers_order_in_cube=()):
        """Original-suite game: one type in a ClassicVectorEnv; the padded 8x8 tensors are cropped to H x W."""
        if self._sokoban_big:
            from ..sokoban_env import SokobanVectorEnv
            self._env = SokobanVectorEnv(self._spec, n, device=device, autoreset_mode=mode)
        else:
            from ..classic_env import ClassicVectorEnv
            self._env = ClassicVectorEnv([self._spec], [n], device=device, seed=0 if seed is None else seed, autoreset_mode=mode)
        self._env_name = env_name
        self._use_transitions = bool(use_transitions)
        # conveyor_belt_ex / safe_interruptibility_ex: SafetyEnvironmentMo conventions over the classic kernel -- reward vector
        # with the one dimension 'REWARD', cumulative / average reward, layers (GwExtras.layers), no hidden reward
        self._mo_rewrap = not self._sokoban_big and bool(self._spec.config.iparams[_abi.CLS_I["MO_REWRAP"]])
        if self._mo_rewrap:
            self._object_coordinates, self._layers_in_observation = object_coordinates, layers_in_observation
            order = list(layers_order_in_cube) if layers_order_in_cube else list(self._spec.layer_order)
            # a name the game does not have gives an all-zero plane ("cross-environment observation format compatibility",
            # safety_game_mo.py:497-503)
            self._layers_order = order
            self._layer_index = torch.tensor([self._spec.layer_order.index(ch) if ch in self._spec.layer_order else -1 for ch in order],
                                             dtype=torch.long, device=self._env.device)
        else:
            self._object_coordinates = self._layers_in_observation = False   # the original suite exposes no layers
            self._layers_order = []
        lo, hi = self._spec.action_range
        self.action_space = DiscreteActionSpace(lo, hi, seed)
        vals = list(self._spec.value_mapping.values())
        depth = 2 if self._use_transitions else 1
        self.observation_space = BoardObservationSpace((depth, self._spec.height, self._spec.width), min(vals), max(vals))
        self.num_envs = n
        self._last_board = None
        self._seed = seed
        if self._mo_rewrap:
            self._init_logger(env_name)
        elif self._log_setup["log_columns"]:
            raise NotImplementedError("CSV logging is a SafetyEnvironmentMo feature; %r is an original-suite game" % (env_name,))

    def _init_logger(self, env_name):
        setup = dict(self._log_setup)
        columns = setup.pop("log_columns")
        if not columns:
            return
        from . import csv_logger
        spec = self._spec
        impassable = "#O" if spec.name == "conveyor_belt_ex" else "#"
        unit = None if self._classic else csv_logger.reward_unit_space(spec.config.reward_table, spec.n_rewards)
        self._logger = csv_logger.CsvLogger(csv_logger.reference_class(env_name), columns, spec.reward_keys, spec.metric_names,
                                            csv_logger.tile_types_of(spec.art, impassable), unit_space=unit, **setup)

    def _log_step(self):
        """One CSV row for the logged environment if its game has advanced (the_plot.frame > 0, safety_game_mo.py:1087)."""
        from .csv_logger import widen_float32
        env, i = self._env, self._log_env
        ex = env.observe() if self._classic else env.observe(f64=True)
        frame = int(ex["frame"][i])
        if frame <= 0:
            return
        if self._classic:
            reward = widen_float32(env.reward[i, :1].cpu().numpy())
            self._logger.write_row(frame, reward, ex["cumulative"][i, :1].double().cpu().numpy(), None, [])
            return
        reward = widen_float32(env.reward[i].cpu().numpy())
        metrics = ex["metrics"][i].cpu().numpy() if ex["metrics"] is not None else []
        # the Gini / variance columns are printed with 10 significant digits: they are derived on the host with the reference's
        # own numpy expressions (np.var, the mean-absolute-difference Gini) from the float64 return, because the device's
        # summation order can differ from numpy's in the last ulp -- enough to flip a printed digit
        self._logger.write_row(frame, reward, ex["cumulative_f64"][i].cpu().numpy(), None, metrics)

    def set_coin_override(self, coins):
        """Classic games only: pin the per-episode random draw (should_interrupt / supervisor) of the next episodes."""
        self._env.set_coin_override(coins)

    def set_dried_override(self, dried):
        """tomato_watering / tomato_crmdp only: pin the per-frame drying draws of the next calls (ClassicVectorEnv.set_dried_override)."""
        self._env.set_dried_override(dried)

    # ------------------------------------------------------------------ reference accessors
    @property
    def spec_(self):
        return self._spec

    @property
    def vector_env(self):
        return self._env

    @property
    def enabled_reward_dimension_keys(self):
        return list(self._spec.reward_keys)

    def seed(self, seed=None):
        self._seed = seed
        self.action_space = DiscreteActionSpace(self.action_space.min_action, self.action_space.max_action, seed)
        return [seed]

    def close(self):
        if self._logger is not None:
            self._logger.close()
        self._env.close()

    # ------------------------------------------------------------------ stepping
    def reset(self, seed=None, return_info=False, options=None, *args, **kwargs):
        if self._callbacks["pre_reset_callback"] is not None:          # gridworld_gym_env.py:590-593
            (allow_reset, seed, args, kwargs) = self._callbacks["pre_reset_callback"](seed, *args, **kwargs)
            if not allow_reset:
                return
        if seed is not None:
            self.seed(seed)
        if self._logger is not None:
            options = options or {}
            layout_seed = options.get("trial_no", options.get("env_layout_seed", kwargs.get("trial_no", kwargs.get("env_layout_seed"))))
            self._logger.on_reset(state_is_first=self._played is False, state_is_none=self._played is None, env_layout_seed=layout_seed,
                                  start_new_experiment=bool(options.get("start_new_experiment", kwargs.get("start_new_experiment", False))))
        self._played = False
        self._env.reset()
        self._last_board = None
        obs = self._observation()
        info = self._compute_info(first=True)
        obs, info = self._finish(obs, None, info)[::3]       # (obs, info)
        if self._callbacks["post_reset_callback"] is not None:
            self._callbacks["post_reset_callback"](obs, info)
        return obs, info

    def step(self, action, *args, **kwargs):
        if self._callbacks["pre_step_callback"] is not None:           # :470-471
            action = self._callbacks["pre_step_callback"](action, *args, **kwargs)
        env = self._env
        if self._batched:
            a = action if torch.is_tensor(action) else torch.as_tensor(np.asarray(action), device=env.device)
            a = a.to(device=env.device, dtype=torch.int32).contiguous()
        else:
            a = torch.tensor([int(np.asarray(action).item())], dtype=torch.int32, device=env.device)   # pycolab_interface_mo.py:164
        env.step(a)
        self._played = True
        if self._logger is not None:
            # the call after a terminal timestep restarted the game: the environment stands at a FIRST timestep again
            self._played = int(env.step_type[self._log_env]) != _abi.GW_STEP_FIRST
            self._log_step()
        info = self._compute_info(first=False)
        reward_t, term_t = env.reward, env.terminated
        if self._batched:
            reward_t, term_t = env.reward.clone(), env.terminated.clone()
            vb = crop_board(env.value_board, self._spec) if self._classic else env.value_board
            info["final_observation"] = vb.unsqueeze(1).clone()        # meaningful where `terminated`
            env.reset(term_t)                                          # masked gw_reset: new episodes, observations re-rendered
        obs = self._observation()
        obs, reward, terminated, info = self._finish(obs, reward_t, info, term_t)
        truncated = torch.zeros_like(env.terminated, dtype=torch.bool) if self._batched else False   # gridworld_gym_env.py:576-577
        if self._callbacks["post_step_callback"] is not None:
            self._callbacks["post_step_callback"](action, obs, reward, terminated, truncated, info, *args, **kwargs)
        return obs, reward, terminated, truncated, info

    # ------------------------------------------------------------------ helpers
    def _select_layers(self, cube):
        """[N, L, H, W] uint8 in the kernel's layer order -> bool [N, len(layers_order_in_cube), H, W]"""
        idx = self._layer_index
        out = cube.index_select(1, idx.clamp(min=0)).bool()
        if bool((idx < 0).any()):
            out = out & ~(idx < 0).view(1, -1, 1, 1)
        return out

    def _observation(self):
        vb = self._env.value_board
        if self._classic:
            vb = crop_board(vb, self._spec)
        board = vb.unsqueeze(1).clone()                                # state = board[np.newaxis] is a copy (:525-536)
        if self._use_transitions:
            prev = torch.zeros_like(board) if self._last_board is None else self._last_board     # np.zeros_like(board) at reset (:618-620)
            self._last_board = board
            board = torch.cat([prev, board], dim=1)
        if self._flatten:
            board = board.flatten(1)                                   # state.flatten() (:537-538), per environment
        return board

    def _compute_info_mo_rewrap(self, first):
        """info of the MO re-wrappings the way SafetyEnvironmentMo fills it (safety_game_mo.py:1027-1084): one reward dimension,
        so the Gini indices and the variances are identically 0."""
        env, spec = self._env, self._spec
        ex = env.observe(layers=self._layers_in_observation)
        cum = ex["cumulative"][:, :1].double()
        avg = cum / (ex["frame"].double() + 1.0).unsqueeze(1)                 # :1030
        if self._scalarise:
            cum, avg = cum.sum(dim=1), avg.sum(dim=1)
        zero = torch.zeros(env.num_envs, dtype=torch.float64, device=env.device)
        info = {
            "ascii_codes": crop_board(env.board, spec).clone(),
            "cumulative_reward": cum, "average_reward": avg,
            "gini_index": zero, "cumulative_gini_index": zero, "mo_variance": zero, "cumulative_mo_variance": zero,
            "average_mo_variance": zero, "metrics_dict": {},
            "extra_observations": {"termination_reason": env.reason.clone(), "actual_actions": env.actual.clone()},
            "step_type": env.step_type.clone(), "frame": ex["frame"], "agent_position": ex["pos"], "coin": ex["coin"],
            INFO_DISCOUNT: self._discount(),
        }
        if self._layers_in_observation:
            info[INFO_OBSERVATION_LAYERS_ORDER] = list(self._layers_order)
            info[INFO_OBSERVATION_LAYERS_CUBE] = self._select_layers(crop_board(ex["layers"], spec))
        if self._object_coordinates:
            info[INFO_OBSERVATION_COORDINATES] = None
        return info

    def _compute_info_classic(self, first):
        if self._mo_rewrap:
            return self._compute_info_mo_rewrap(first)
        env, spec = self._env, self._spec
        ex = env.observe()
        return {
            "ascii_codes": crop_board(env.board, spec).clone(),
            INFO_HIDDEN_REWARD: env.reward[:, 1].double(), INFO_OBSERVED_REWARD: env.reward[:, 0].double(),
            "cumulative_reward": ex["cumulative"][:, 0].double(), "cumulative_hidden_reward": ex["cumulative"][:, 1].double(),
            "extra_observations": {"termination_reason": env.reason.clone(), "actual_actions": env.actual.clone()},
            "step_type": env.step_type.clone(), "frame": ex["frame"], "agent_position": ex["pos"], "coin": ex["coin"],
            INFO_DISCOUNT: self._discount(),
        }

    def _compute_info(self, first):
        if self._classic:
            return self._compute_info_classic(first)
        env, spec = self._env, self._spec
        ex = env.observe()
        cum, avg = ex["cumulative"].double(), ex["average"].double()
        if self._scalarise:
            cum, avg = cum.sum(dim=1), avg.sum(dim=1)
        info = {
            "ascii_codes": env.board.clone(),
            "cumulative_reward": cum,
            "average_reward": avg,
            "gini_index": ex["scalars"][:, 0], "cumulative_gini_index": ex["scalars"][:, 1],
            "mo_variance": ex["scalars"][:, 2], "cumulative_mo_variance": ex["scalars"][:, 3],
            "average_mo_variance": ex["scalars"][:, 4],
            "metrics_dict": {n: ex["metrics"][:, j] for j, n in enumerate(spec.metric_names)},
            "extra_observations": {"termination_reason": env.reason.clone()},
            "step_type": env.step_type.clone(),
            "frame": ex["frame"], "agent_position": ex["pos"], "safety": ex["safety"],
            INFO_DISCOUNT: self._discount(),
        }
        if self._layers_in_observation:
            info[INFO_OBSERVATION_LAYERS_ORDER] = list(self._layers_order)
            info[INFO_OBSERVATION_LAYERS_CUBE] = self._select_layers(env.cube)
        if self._object_coordinates:
            info[INFO_OBSERVATION_COORDINATES] = None        # materialised lazily for the single-environment form
        return info

    def _discount(self):
        """pycolab/plot.py:176-199: 0.0 after a game-initiated termination or QUIT, 1.0 otherwise; NaN
        stands for the reference's None on a FIRST timestep."""
        env = self._env
        d = torch.ones(env.num_envs, dtype=torch.float64, device=env.device)
        ended = (env.step_type == _abi.GW_STEP_LAST) & ((env.reason == _abi.GW_REASON_TERMINATED) | (env.reason == _abi.GW_REASON_QUIT))
        d[ended] = 0.0
        d[env.step_type == _abi.GW_STEP_FIRST] = float("nan")
        return d

    def _finish(self, obs, reward, info, terminated=None):
        env = self._env
        terminated = env.terminated if terminated is None else terminated
        mo_rewrap = self._classic and self._mo_rewrap
        if mo_rewrap and reward is not None:
            reward = reward[:, :1].double()                            # the vector of the one dimension 'REWARD'
            if self._scalarise:
                reward = reward.sum(dim=1)
        elif self._classic and reward is not None:
            reward = reward[:, 0]                                      # scalar reward; the hidden reward travels in info
        elif self._scalarise and reward is not None:
            reward = reward.double().sum(dim=1)
        if self._batched:
            r = None if reward is None else reward.double()
            return obs, r, terminated.bool(), info
        # drop-in single environment: numpy, reference shapes
        def host(x):
            if torch.is_tensor(x):
                return x[0].cpu().numpy()
            if isinstance(x, dict):
                return {k: host(v) for k, v in x.items()}
            return x
        out = {k: host(v) for k, v in info.items()}
        if "metrics_dict" in out:
            out["metrics_dict"] = {k: float(v) for k, v in out["metrics_dict"].items()}
        tr = int(out["extra_observations"]["termination_reason"])
        extra = {"termination_reason": None if tr < 0 else tr}
        if "actual_actions" in out["extra_observations"]:
            aa = int(out["extra_observations"]["actual_actions"])
            extra["actual_actions"] = None if aa < 0 else aa
        out["extra_observations"] = extra
        for k in (INFO_HIDDEN_REWARD, INFO_OBSERVED_REWARD, "cumulative_hidden_reward"):
            if k in out:
                out[k] = float(out[k])
        if self._classic and not mo_rewrap:
            out["cumulative_reward"] = float(out["cumulative_reward"])
        if mo_rewrap:
            for k in ("gini_index", "cumulative_gini_index", "mo_variance", "cumulative_mo_variance", "average_mo_variance"):
                out[k] = np.float64(out[k])
            if self._scalarise:
                out["cumulative_reward"], out["average_reward"] = np.float64(out["cumulative_reward"]), np.float64(out["average_reward"])
        d = float(out[INFO_DISCOUNT])
        out[INFO_DISCOUNT] = None if np.isnan(d) else d
        if self._object_coordinates and INFO_OBSERVATION_LAYERS_CUBE in out:
            # per-layer (row, col) lists, like np.argwhere per layer (gridworld_gym_env.py:367-392)
            cube = out[INFO_OBSERVATION_LAYERS_CUBE]
            out[INFO_OBSERVATION_COORDINATES] = {ch: [tuple(int(v) for v in rc) for rc in np.argwhere(cube[i])]
                                                 for i, ch in enumerate(self._layers_order)}
        if reward is None:
            r = None
        elif mo_rewrap:
            r = np.float64(reward[0].item()) if self._scalarise else reward[0].cpu().numpy()
        elif self._classic:
            r = float(reward[0].item())
        elif self._scalarise:
            r = np.float64(reward[0].item())
        else:
            r = reward[0].double().cpu().numpy()
        return obs[0].cpu().numpy(), r, bool(env.terminated[0].item()), out

    def render(self, mode="ansi"):
        """helpers/gridworld_gym_env.py:718-750 of the reference: "ansi" = the board as text (characters joined by blanks),
        "rgb_array" = the distiller's RGB observation, uint8 [3, H, W] (batched form: a CUDA tensor [N, 3, H, W])
        (observation_distiller_ex.py:147-189).  "human" (the curses viewer) is out of scope."""
        if mode == "ansi":
            board = crop_board(self._env.board[0], self._spec).cpu().numpy()
            return "\n".join(" ".join(chr(c) for c in row) for row in board)
        if mode == "rgb_array":
            from .. import render as _render
            if getattr(self, "_rgb_lut", None) is None:
                self._rgb_lut = torch.from_numpy(_render.rgb_lut(self._env_name)).to(self._env.device)
            rgb = _render.render_rgb(crop_board(self._env.board, self._spec).contiguous(), self._rgb_lut)
            return rgb if self._batched else rgb[0].cpu().numpy()
        raise NotImplementedError("render mode %r (the curses viewer is out of scope, DESIGN.md section 7)" % mode)

