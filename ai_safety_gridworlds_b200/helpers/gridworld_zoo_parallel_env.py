"""GridworldZooParallelEnv: the reference's PettingZoo parallel signature over the CUDA backend.

Mirrors helpers/gridworld_zoo_parallel_env.py of the reference (constructor :100-135, `reset`
:618-702, `step` :429-615, `agents` :252-253) for firemaker_ex_ma (agents '1', '2', 'S') and
island_navigation_ex_ma (agents '1', '2'):

  * `num_envs=None` is the drop-in single environment: dicts keyed 'agent_<chr>' of numpy values
    with the reference's shapes -- observations '<U1' [1,5,5] (the firemaker supervisor: [1,33,33];
    float32 value-mapped with `ascii_observation_format=False`), rewards float64 [R], terminateds bool,
    truncateds False, infos -- done agents leave `agents` and the returned dicts, stepping a done
    agent raises ValueError (rl/pycolab_interface_ma.py:217-218) and a finished game needs `reset()`;
  * `num_envs=N` is the batched form: the same dicts of torch CUDA tensors with a leading batch
    dimension (observations as uint8 ASCII codes or float32), auto-reset inside the step that ends
    a game; the action entries of agents that are done in an environment are ignored.

pettingzoo itself is not imported (duck-typed).  Every step is one kernel launch; there is no CPU
fallback.
"""
import numpy as np
import torch

from .. import _abi
from ..envs import make_spec
from ..firemaker_env import FiremakerVectorEnv
from ..island_ma_env import IslandMaVectorEnv
from .gridworld_gym_env import (DiscreteActionSpace, INFO_OBSERVATION_LAYERS_CUBE, INFO_OBSERVATION_LAYERS_ORDER)

INFO_OBSERVATION_COORDINATES = "info_observation_coordinates"
INFO_OBSERVATION_LAYERS_DICT = "info_observation_layers_dict"
INFO_AGENT_OBSERVATIONS = "info_agent_observations"
INFO_AGENT_OBSERVATION_COORDINATES = "info_agent_observation_coordinates"
INFO_AGENT_OBSERVATION_LAYERS_DICT = "info_agent_observation_layers_dict"
INFO_AGENT_OBSERVATION_LAYERS_ORDER = "info_agent_observation_layers_order"
INFO_AGENT_OBSERVATION_LAYERS_CUBE = "info_agent_observation_layers_cube"


class GridworldsObservationSpace(object):
    """The per-agent observation space of the reference wrapper (helpers/gridworld_zoo_parallel_env.py:944-1013), duck-typed
    (gymnasium is not imported): shape (1 | 2, h, w) of the agent's view -- (2, ...) with `use_transitions`, flattened with
    `flatten_observations` -- dtype '<U1' for the ascii format, float32 for the value-mapped board."""

    def __init__(self, view_shape, ascii_format, use_transitions, flatten_observations):
        view_shape = tuple(int(v) for v in view_shape)
        self.use_transitions, self.flatten_observations = bool(use_transitions), bool(flatten_observations)
        cells = int(np.prod(view_shape))
        if self.flatten_observations:
            self.shape = (2, cells) if self.use_transitions else (cells,)
        else:
            self.shape = ((2,) if self.use_transitions else (1,)) + view_shape
        self.dtype = np.dtype("<U1") if ascii_format else np.dtype(np.float32)

    def sample(self):
        """Not a random sample: an example observation, as the reference returns (:973-1000)."""
        if self.use_transitions:
            raise NotImplementedError("Sampling from transition-based envs not yet supported.")
        return np.zeros(self.shape, self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return tuple(x.shape) == tuple(self.shape) and (x.dtype.kind == "U" if self.dtype.kind == "U" else x.dtype == self.dtype)

    __contains__ = contains


class _FiremakerBackend(object):
    """Per-agent views of FiremakerVectorEnv's tensors (workers share [N,2,...] tensors, the supervisor has its own).  The kernel
    always has three agent columns ('1', '2', 'S'); with amount_agents = 2 (the reference's default: one worker + the
    supervisor) column 1 is unused and agent i of the wrapper lives in column cols[i]."""
    n_cols = 3

    def __init__(self, n, device, seed, mode, spec):
        self.env = FiremakerVectorEnv(n, device=device, seed=seed, autoreset_mode=mode, spec=spec)
        self.agent_chars = ["1", "2", "S"] if spec.config.amount_agents == 3 else ["1", "S"]
        self.cols = [0, 1, 2] if spec.config.amount_agents == 3 else [0, 2]
        from ..envs.firemaker_ex_ma import METRIC_NAMES
        self._metric_cols = [METRIC_NAMES.index(m) for m in spec.metric_names]

    def crop(self, i):
        k = self.cols[i]
        return self.env.crop_supervisor if k == 2 else self.env.crop_workers[:, k]

    def lcrop(self, i):
        k = self.cols[i]
        return self.env.lcrop_supervisor if k == 2 else self.env.lcrop_workers[:, k]

    def reward(self, i):
        k = self.cols[i]
        return self.env.reward_supervisor if k == 2 else self.env.reward_workers[:, k]

    def step(self, act, order, draws):
        self.env.step(act, order, draws)

    def extras(self):
        ex = self.env.observe()
        cum = ex["cumulative"]
        per = [cum[:, 0:2], cum[:, 2:4], cum[:, 4:7]]
        ex["cumulative_per_agent"] = [per[k] for k in self.cols]
        ex["metrics"] = ex["metrics"][:, self._metric_cols]
        ex["external_fires"] = ex["ext_fires"]
        return ex


class _IslandMaBackend(object):
    agent_chars = ["1", "2"]
    cols = [0, 1]
    n_cols = 2

    def __init__(self, n, device, seed, mode, spec):
        self.env = IslandMaVectorEnv(n, device=device, seed=seed, autoreset_mode=mode, spec=spec)

    def crop(self, i):
        return self.env.crop[:, i]

    def lcrop(self, i):
        return self.env.lcrop[:, i]

    def reward(self, i):
        return self.env.reward[:, i]

    def step(self, act, order, draws):
        if draws is not None:
            raise ValueError("island_navigation_ex_ma draws no random numbers inside a step")
        self.env.step(act, order)

    def extras(self):
        ex = self.env.observe()
        ex["cumulative_per_agent"] = [ex["cumulative"][:, 0], ex["cumulative"][:, 1]]
        return ex


class _SavannaBackend(object):
    """aintelope_savanna and its experiment overlays: agents '0' (and '1' with amount_agents = 2); the kernel always has two agent
    columns."""
    n_cols = 2

    def __init__(self, n, device, seed, mode, spec):
        from ..savanna_env import SavannaVectorEnv
        self.env = SavannaVectorEnv(n, device=device, seed=seed, autoreset_mode=mode, spec=spec)
        self.agent_chars = ["0", "1"][:spec.n_agents]
        self.cols = [0, 1][:spec.n_agents]

    def crop(self, i):
        return self.env.crop[:, i]

    def lcrop(self, i):
        return self.env.lcrop[:, i]

    def reward(self, i):
        return self.env.reward[:, i]

    def step(self, act, order, draws):
        self.env.step(act, order, None if draws is None else draws[:, :4 * _abi.GW_SAV_MAX_DRAWS].contiguous())   # predator draws + the resource drapes' tile picks

    def extras(self):
        ex = self.env.observe()
        ex["cumulative_per_agent"] = [ex["cumulative"][:, k] for k in self.cols]
        return ex


_BACKENDS = {"firemaker_ex_ma": _FiremakerBackend, "island_navigation_ex_ma": _IslandMaBackend, "aintelope_savanna": _SavannaBackend}


def _backend_for(env_name):
    """The backend class of a factory name; the aintelope experiment overlays are aintelope_savanna games."""
    name = env_name.lower()
    if name in _BACKENDS:
        return _BACKENDS[name]
    from ..envs import savanna_experiments
    if name in savanna_experiments.OVERLAYS:
        return _SavannaBackend
    return None


class GridworldZooParallelEnv(object):
    metadata = {"render.modes": ["ansi", "rgb_array"], "name": "gridworld_zoo_parallel_env_b200"}

    def __init__(self, env_name, use_transitions=False, render_animation_delay=0.1, flatten_observations=False,
                 ascii_observation_format=True, object_coordinates_in_observation=None, layers_in_observation=True,
                 occlusion_in_layers=False, layers_order_in_cube=[], layers_order_in_cube_per_agent={},
                 ascii_attributes_format=False, attribute_coordinates_in_observation=True, layers_in_attribute_observation=False,
                 occlusion_in_atribute_layers=False, observable_attribute_categories=None, observable_attribute_value_mapping=None,
                 use_multi_discrete_action_space=False, np_random=None, seed=None, test_death=False, test_death_probability=0.33,
                 pre_reset_callback=None, post_reset_callback=None, pre_step_callback=None, post_step_callback=None,
                 render_mode=None, num_envs=None, device=None, final_info=True, **kwargs):
        # the wrapper-side options of the reference's constructor (:100-135) are honoured or refused -- never dropped
        if occlusion_in_layers:
            raise NotImplementedError("occlusion_in_layers=True is not built (the reference's own occluded branch is unfinished: "
                                      "safety_game_moma.py:605-618 raises NameError / NotImplementedError)")
        if use_multi_discrete_action_space:
            raise NotImplementedError("use_multi_discrete_action_space: the CUDA backend takes one discrete action per agent")
        if ascii_attributes_format or layers_in_attribute_observation or occlusion_in_atribute_layers:
            raise NotImplementedError("observable agent attributes (expression / numeric_message layers) are not built")
        if observable_attribute_value_mapping:
            raise NotImplementedError("observable_attribute_value_mapping is not built")
        self.render_mode = render_mode
        self._render_animation_delay = render_animation_delay
        if _backend_for(env_name) is None:
            raise NotImplementedError("the multi-agent CUDA backend is built for " + ", ".join(sorted(_BACKENDS)) + " and the "
                                      "aintelope experiment overlays")
        self._env_name = env_name
        self._batched = num_envs is not None
        n = int(num_envs) if self._batched else 1
        # Batched form with final_info (default): a finished game restarts inside step(), but after the step's infos were read -- the
        # infos (cumulative rewards, metrics, frame, the boards and views they carry) describe the game that just ended, like the
        # rewards and terminateds do, and the returned observations already belong to the next game.  The kernel then runs with the
        # reference's own reset semantics and the wrapper issues the masked reset.  final_info=False leaves the restart to the kernel
        # (auto-reset inside the ending launch; the AEC wrapper steps the backend that way).
        self._final_info = self._batched and bool(final_info)
        mode = _abi.GW_AUTORESET_SAME_STEP if (self._batched and not self._final_info) else _abi.GW_AUTORESET_NEXT_STEP
        self._spec = make_spec(env_name, autoreset_mode=mode, **kwargs)
        self._backend = _backend_for(env_name)(n, device, 0 if seed is None else seed, mode, self._spec)
        self._env = self._backend.env
        self._ascii = bool(ascii_observation_format)
        self._use_transitions = bool(use_transitions)
        self._flatten = bool(flatten_observations)
        self._layers_in_observation = bool(layers_in_observation)
        # per-cell coordinate LISTS are a single-environment convenience (python lists of tuples); the batched form carries the
        # same information in the layer cubes
        if object_coordinates_in_observation is None:
            object_coordinates_in_observation = not self._batched
        if object_coordinates_in_observation and self._batched:
            raise NotImplementedError("object_coordinates_in_observation is built for the single-environment form; the batched "
                                      "form returns the layer cubes (same information, tensor-shaped)")
        self._object_coordinates = bool(object_coordinates_in_observation)
        chars = self._backend.agent_chars
        self.possible_agents = ["agent_" + ch for ch in chars]
        self.agent_name_mapping = dict(zip(self.possible_agents, chars))
        self.agent_name_reverse_mapping = {v: k for k, v in self.agent_name_mapping.items()}
        dev = self._env.device
        # layers_order_in_cube (:110) / layers_order_in_cube_per_agent (:111): the cube's channel order; [] = every layer, sorted;
        # a name the game does not have gives an all-zero plane ("for cross-environment observation format compatibility",
        # safety_game_moma.py:661-665); None = no cube at all (:296, :310)
        self._layers_order = None if layers_order_in_cube is None else self._order(layers_order_in_cube)
        if layers_order_in_cube_per_agent is None:
            self._agent_layers_order = None
        else:
            unknown_agents = [a for a in layers_order_in_cube_per_agent if a not in self.possible_agents]
            if unknown_agents:
                raise ValueError("layers_order_in_cube_per_agent names unknown agents %r (agents: %r)" % (unknown_agents, self.possible_agents))
            self._agent_layers_order = {a: self._order(layers_order_in_cube_per_agent.get(a, [])) for a in self.possible_agents}
        lo, hi = self._spec.action_range
        self.action_spaces = {a: DiscreteActionSpace(lo, hi, seed) for a in self.possible_agents}
        self.observation_spaces = {a: GridworldsObservationSpace(tuple(self._backend.crop(i).shape[1:]), self._ascii, self._use_transitions,
                                                                 self._flatten) for i, a in enumerate(self.possible_agents)}
        self.num_envs = n
        self._dones = {a: False for a in self.possible_agents}
        # test_death (:124-125, :577-586): every step each live agent "dies" with test_death_probability -- a wrapper-side fault
        # injection for the consumers' handling of agents that disappear.  Single form: the reference's own generator calls
        # (`np_random.random()`, a numpy Generator: `np_random` or one seeded with `seed`); batched form: one Philox4x32-10 draw
        # per (seed, global environment, agent, step), sticky until the game restarts.
        self._test_death = bool(test_death)
        self._test_death_probability = float(test_death_probability)
        self._test_deads = {a: False for a in self.possible_agents}
        self._np_random = np_random if np_random is not None else np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        self._death_seed = 0 if seed is None else int(seed)
        self._death_step = 0
        self._death_mask = torch.zeros((n, len(chars)), dtype=torch.bool, device=dev) if self._batched else None
        self._replay_death_draws = None
        self._pre_reset_callback, self._post_reset_callback = pre_reset_callback, post_reset_callback
        self._pre_step_callback, self._post_step_callback = pre_step_callback, post_step_callback
        self._last_agent_boards = {a: None for a in self.possible_agents}
        lut = torch.zeros(256, dtype=torch.float32, device=dev)
        for ch, v in self._spec.value_mapping.items():
            lut[ord(ch)] = v
        self._lut = lut
        self._rgb_lut = None

    def _order(self, names):
        """(layer names, index tensor into the kernel's layer axis with -1 for names the game does not have)"""
        have = list(self._spec.layer_order)
        names = list(names) if names else list(have)
        idx = [have.index(ch) if ch in have else -1 for ch in names]
        return names, torch.tensor(idx, dtype=torch.long, device=self._env.device)

    def _select_layers(self, cube, order):
        """cube [N, L, h, w] uint8 in the kernel's (sorted) layer order -> bool [N, len(order), h, w]"""
        names, idx = order
        out = cube.index_select(1, idx.clamp(min=0)).bool()
        missing = idx < 0
        if bool(missing.any()):
            out = out & ~missing.view(1, -1, 1, 1)
        return out

    # ------------------------------------------------------------------ PettingZoo surface
    @property
    def agents(self):
        return [a for a in self.possible_agents if not self._dones[a]]

    @property
    def num_agents(self):
        return len(self.agents)

    @property
    def max_num_agents(self):
        return len(self.possible_agents)

    def action_space(self, agent):
        return self.action_spaces[agent]

    def observation_space(self, agent):
        return self.observation_spaces[agent]

    @property
    def vector_env(self):
        return self._env

    def close(self):
        self._env.close()

    def seed(self, seed=None):
        for sp in self.action_spaces.values():
            sp._rng = np.random.default_rng(seed)

    def render(self, mode="ansi"):
        """"ansi": environment 0's board as text; "rgb_array": the distiller's RGB observation of the global board, uint8 [3, H, W]
        (batched form: CUDA tensor [N, 3, H, W]) -- helpers/gridworld_zoo_parallel_env.py `render`, observation_distiller_ex.py:147-189."""
        if mode == "ansi":
            return "\n".join(" ".join(chr(c) for c in row) for row in self._env.board[0].cpu().numpy())
        if mode == "rgb_array":
            from .. import render as _render
            if self._rgb_lut is None:
                self._rgb_lut = torch.from_numpy(_render.rgb_lut(self._env_name)).to(self._env.device)
            rgb = _render.render_rgb(self._env.board, self._rgb_lut)
            return rgb if self._batched else rgb[0].cpu().numpy()
        raise NotImplementedError("render mode %r (the curses viewer is out of scope)" % mode)

    def reset(self, seed=None, *args, **kwargs):
        if self._pre_reset_callback is not None:
            (allow_reset, seed, args, kwargs) = self._pre_reset_callback(seed, *args, **kwargs)
            if not allow_reset:
                return
        if seed is not None:
            self.seed(seed)
        self._env.reset()
        self._dones = {a: False for a in self.possible_agents}
        self._test_deads = {a: False for a in self.possible_agents}
        if self._death_mask is not None:
            self._death_mask.zero_()
        self._last_agent_boards = {a: None for a in self.possible_agents}
        obs, infos = self._observations(first=True), self._infos()
        if self._post_reset_callback is not None:
            self._post_reset_callback(obs, infos)
        return obs, infos

    def step(self, actions, *args, replay_order=None, replay_draws=None, replay_death_draws=None, **kwargs):
        """`replay_order` (KERNEL agent columns in execution order -- '1', '2', 'S' = 0, 1, 2 -- with -1 = no frame), `replay_draws` (the
        FireDrape uniform draws of this step, in call order) and `replay_death_draws` (the test_death draws of this step, in the
        order the reference's wrapper makes them) replay a recorded reference run -- test hooks of the single-environment form."""
        if self._pre_step_callback is not None:
            actions = self._pre_step_callback(actions, *args, **kwargs)
        env, A, be = self._env, self._backend.n_cols, self._backend
        if self._batched:
            act = torch.zeros((self.num_envs, A), dtype=torch.int32, device=env.device)
            for i, a in enumerate(self.possible_agents):
                v = actions[a]
                v = v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v), device=env.device)
                act[:, be.cols[i]] = v.to(device=env.device, dtype=torch.int32).reshape(-1)
            lo, hi = self._spec.action_range
            if bool(((act < lo) | (act > hi)).any()):
                # QUIT (9) and anything else outside the game's action range: the kernels have no such frame; the reference ends
                # the agent's episode on QUIT (safety_game_ma.py), so a silent "non-NOOP non-move" would diverge from it
                raise NotImplementedError("actions outside %d..%d (QUIT included) are not supported by the multi-agent CUDA backend" % (lo, hi))
            stepped = list(self.possible_agents)
        else:
            for a in actions:
                if self._dones.get(a, False):
                    raise ValueError("Agent %s is done" % self.agent_name_mapping[a])     # pycolab_interface_ma.py:217-218
            if not self.agents:
                raise ValueError("all agents are done: call reset()")
            stepped = [a for a in self.possible_agents if not self._dones[a]]
            vals = [0] * A
            for i, a in enumerate(self.possible_agents):
                v = actions.get(a, 0)
                v = v["step"] if isinstance(v, dict) else v                               # {"step": a} modality (safety_game_ma.py:412-423)
                vals[be.cols[i]] = int(np.asarray(v).item())
            if any(v == 9 for v in vals):
                raise NotImplementedError("QUIT is not supported by the multi-agent CUDA backend")
            act = torch.tensor([vals], dtype=torch.int32, device=env.device)
        order = draws = None
        if replay_order is None and not self._batched and any(self._test_deads.values()):
            # a virtually dead agent is alive inside the game but gets no action any more: the reference's core plays frames for the
            # acting agents only (pycolab_interface_ma.py:183-230), so the kernel is given the execution order without it
            acting = [be.cols[i] for i, a in enumerate(self.possible_agents) if a in stepped]
            if len(acting) > 1 and self._spec.flags.get("randomize_agent_actions_order", True):
                self._np_random.shuffle(acting)
            replay_order = acting + [-1] * (A - len(acting))
        if replay_order is not None:
            order = torch.tensor(np.asarray(replay_order, np.int32).reshape(1, A), device=env.device)
        if replay_draws is not None:
            dr = np.full((1, _abi.GW_FM_MAX_DRAWS), 2.0)
            dr[0, :len(replay_draws)] = replay_draws
            draws = torch.from_numpy(dr).to(env.device)
        self._backend.step(act, order, draws)
        infos = self._infos(stepped, clone=self._final_info)
        rewards, terms, truncs = {}, {}, {}
        for i, a in enumerate(self.possible_agents):
            if a not in stepped:
                continue
            r = self._backend.reward(i)
            t = env.terminated[:, be.cols[i]].bool()
            if self._batched:
                rewards[a], terms[a], truncs[a] = r.double(), t, torch.zeros_like(t)
            else:
                rewards[a], terms[a], truncs[a] = r[0].double().cpu().numpy(), bool(t[0].item()), False
        if self._test_death:
            self._apply_test_death(stepped, rewards, terms, replay_death_draws)
        if self._batched:
            over = (env.step_type[:, be.cols] >= 2).all(dim=1)          # every agent of the game is done (LAST / DEAD)
            if self._death_mask is not None:
                self._death_mask[over] = False                           # a new game: nobody is (virtually) dead
            if self._final_info:
                env.reset(over)                                          # restart it now that its final infos are out
        obs = self._observations(stepped)
        if not self._batched:
            self._dones.update(terms)
        result = (obs, rewards, terms, truncs, infos)
        if self._post_step_callback is not None:
            self._post_step_callback(actions, *result, *args, **kwargs)
        return result

    def _apply_test_death(self, stepped, rewards, terms, replay):
        """helpers/gridworld_zoo_parallel_env.py:577-586.  Single form: for every possible agent in order -- a virtually dead agent
        loses its reward entry (and the mark, once it is really done); a live one dies with `test_death_probability` (one
        `np_random.random()` draw, made only for agents that are not done, exactly as the reference's `elif` does)."""
        if not self._batched:
            it = iter(replay) if replay is not None else None
            for a in self.possible_agents:
                done = terms.get(a, True)                  # agents removed earlier are done
                if self._test_deads[a]:
                    rewards.pop(a, None)
                    if done:
                        self._test_deads[a] = False
                elif not done:
                    u = next(it) if it is not None else float(self._np_random.random())
                    if u < self._test_death_probability:
                        if a in terms:
                            terms[a] = True
                        self._test_deads[a] = True
            return
        # batched: Philox4x32-10 keyed by (seed, global environment index, step * agents + agent); a dead mark is sticky until the
        # agent's game restarts (its step type returns to FIRST) and shows as terminated = True with a zeroed reward row
        env = self._env
        self._death_step += 1
        n, A = self.num_envs, len(self.possible_agents)
        for i, a in enumerate(self.possible_agents):
            col = self._backend.cols[i]
            u = _philox_uniform(self._death_seed ^ 0x7e57dea7, getattr(env, "env_index_base", 0), n, self._death_step * A + i, env.device)
            newly = (~self._death_mask[:, i]) & (~terms[a]) & (u < self._test_death_probability)
            rewards[a] = torch.where(self._death_mask[:, i].unsqueeze(-1), torch.zeros_like(rewards[a]), rewards[a])
            self._death_mask[:, i] |= newly
            terms[a] = terms[a] | self._death_mask[:, i]

    # ------------------------------------------------------------------ helpers
    def _crop(self, i):
        return self._backend.crop(i)

    def _lcrop(self, i):
        return self._backend.lcrop(i)

    def _observations(self, agents=None, first=False):
        """Per-agent states (:530-560, :660-684): the agent's view [1, h, w]; with use_transitions [2, h, w] = (previous view, view)
        with zeros before the first step ('' for the ascii format: np.zeros_like of a '<U1' array); flattened on request."""
        out = {}
        for i, a in enumerate(self.possible_agents):
            if agents is not None and a not in agents:
                continue
            codes = self._crop(i)
            if self._batched:
                board = codes.clone() if self._ascii else self._lut[codes.long()]
            elif self._ascii:
                board = np.vectorize(chr)(codes[0].cpu().numpy())                             # '<U1' [h, w]
            else:
                board = self._lut[codes[0].long()].cpu().numpy()
            if self._use_transitions:
                last = self._last_agent_boards[a]
                if first or last is None:
                    last = torch.zeros_like(board) if self._batched else np.zeros_like(board)
                state = torch.stack([last, board], dim=1) if self._batched else np.stack([last, board], axis=0)
                self._last_agent_boards[a] = board
            else:
                state = board.unsqueeze(1) if self._batched else board[np.newaxis, :]
            if self._flatten:
                state = state.flatten(1) if self._batched else state.flatten()
            out[a] = state
        return out

    @staticmethod
    def _coordinates(layers, names):
        """{layer: [(row, col), ...]} of a bool [L, h, w] cube (calculate_observation_coordinates, safety_game_moma.py:583-601)"""
        return {ch: [tuple(c) for c in np.argwhere(layers[k]).tolist()] for k, ch in enumerate(names)}

    def _infos(self, agents=None, clone=False):
        """`clone`: the board / view tensors are copied (the caller is about to restart finished games in place)"""
        env, spec = self._env, self._spec
        keep = (lambda t: t.clone()) if clone else (lambda t: t)
        ex = self._backend.extras()
        have = list(spec.layer_order)
        infos = {}
        global_cube = env.cube
        for i, a in enumerate(self.possible_agents):
            if agents is not None and a not in agents:
                continue
            info = {
                "ascii_codes": keep(env.board),
                INFO_AGENT_OBSERVATIONS: keep(self._crop(i)),
                "cumulative_reward": ex["cumulative_per_agent"][i].double(),
                "metrics_dict": {n: ex["metrics"][:, j] for j, n in enumerate(spec.metric_names)},
                "frame": ex["frame"], "agent_positions": ex["pos"],
                "step_type": keep(env.step_type[:, self._backend.cols[i]]),
            }
            if self._layers_order is not None:
                info[INFO_OBSERVATION_LAYERS_ORDER] = list(self._layers_order[0])
                info[INFO_OBSERVATION_LAYERS_CUBE] = self._select_layers(global_cube, self._layers_order)
            if self._agent_layers_order is not None:
                info[INFO_AGENT_OBSERVATION_LAYERS_ORDER] = list(self._agent_layers_order[a][0])
                info[INFO_AGENT_OBSERVATION_LAYERS_CUBE] = self._select_layers(self._lcrop(i), self._agent_layers_order[a])
            if "external_fires" in ex:
                info["external_fires"] = ex["external_fires"]
            if "directions" in ex:
                info["action_direction"] = ex["directions"][:, self._backend.cols[i], 0]
                info["observation_direction"] = ex["directions"][:, self._backend.cols[i], 1]
            if not self._batched:
                def host(x):
                    if torch.is_tensor(x):
                        return x[0].cpu().numpy()
                    if isinstance(x, dict):
                        return {k: host(v) for k, v in x.items()}
                    return x
                info = {k: host(v) for k, v in info.items()}
                info["metrics_dict"] = {k: float(v) for k, v in info["metrics_dict"].items()}
                g_layers = global_cube[0].bool().cpu().numpy()
                a_layers = self._lcrop(i)[0].bool().cpu().numpy()
                if self._layers_in_observation:
                    info[INFO_OBSERVATION_LAYERS_DICT] = {ch: g_layers[k] for k, ch in enumerate(have)}
                    info[INFO_AGENT_OBSERVATION_LAYERS_DICT] = {ch: a_layers[k] for k, ch in enumerate(have)}
                if self._object_coordinates:
                    # global: absolute (row, col); per agent: (x - agent x, y - agent y) relative to the agent's own cell in its view
                    # (calculate_agents_observation_coordinates, safety_game_moma.py:528-580); [] if the agent is not in its view
                    info[INFO_OBSERVATION_COORDINATES] = self._coordinates(g_layers, have)
                    ch = self.agent_name_mapping[a]
                    own = np.argwhere(a_layers[have.index(ch)]) if ch in have else np.zeros((0, 2), int)
                    if len(own) > 0:
                        ay, ax = int(own[0][0]), int(own[0][1])
                        info[INFO_AGENT_OBSERVATION_COORDINATES] = {
                            key: [(x - ax, y - ay) for (y, x) in coords] for key, coords in self._coordinates(a_layers, have).items()}
                    else:
                        info[INFO_AGENT_OBSERVATION_COORDINATES] = []
            infos[a] = info
        return infos


def _philox_uniform(seed, env_index_base, n, step, device):
    """float64 [n] uniforms in [0, 1): Philox4x32-10 keyed by `seed`, counter (global environment index, step) -- the library's
    stream definition (gw_random_actions / or_philox), evaluated with torch integer arithmetic for this wrapper-side draw."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    mask = 0xFFFFFFFF
    g = torch.arange(n, dtype=torch.int64, device=device) + int(env_index_base)
    c0, c1 = g & mask, (g >> 32) & mask
    c2 = torch.full_like(g, int(step) & mask)
    c3 = torch.full_like(g, (int(step) >> 32) & mask)
    k0, k1 = int(seed) & mask, (int(seed) >> 32) & mask
    for _ in range(10):
        # 32 x 32 -> 64-bit products reach 2^64: the multiplicand is split into 16-bit halves to stay inside int64
        hi0 = ((c0 >> 16) * M0 + (((c0 & 0xFFFF) * M0) >> 16)) >> 16
        lo0 = (((c0 & 0xFFFF) * M0) + (((c0 >> 16) * M0 & 0xFFFF) << 16)) & mask
        hi1 = ((c2 >> 16) * M1 + (((c2 & 0xFFFF) * M1) >> 16)) >> 16
        lo1 = (((c2 & 0xFFFF) * M1) + (((c2 >> 16) * M1 & 0xFFFF) << 16)) & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0, k1 = (k0 + W0) & mask, (k1 + W1) & mask
    bits = (c0 << 21) | (c1 >> 11)                          # ((c0 << 32) | c1) >> 11 without leaving int64
    return bits.double() * (1.0 / 9007199254740992.0)
