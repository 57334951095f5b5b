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

INFO_AGENT_OBSERVATIONS = "info_agent_observations"
INFO_AGENT_OBSERVATION_LAYERS_ORDER = "info_agent_observation_layers_order"
INFO_AGENT_OBSERVATION_LAYERS_CUBE = "info_agent_observation_layers_cube"

_WRAPPER_ONLY = ("use_transitions", "render_animation_delay", "flatten_observations", "object_coordinates_in_observation",
                 "layers_in_observation", "occlusion_in_layers", "layers_order_in_cube", "layers_order_in_cube_per_agent",
                 "ascii_attributes_format", "attribute_coordinates_in_observation", "layers_in_attribute_observation",
                 "occlusion_in_atribute_layers", "observable_attribute_categories", "observable_attribute_value_mapping",
                 "use_multi_discrete_action_space", "np_random", "test_death", "test_death_probability", "pre_reset_callback",
                 "post_reset_callback", "pre_step_callback", "post_step_callback", "render_mode")


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
    metadata = {"render.modes": ["ansi"], "name": "gridworld_zoo_parallel_env_b200"}

    def __init__(self, env_name, ascii_observation_format=True, seed=None, num_envs=None, device=None, **kwargs):
        if kwargs.get("test_death"):
            raise NotImplementedError("test_death fault injection is a wrapper-side debugging aid and is not built")
        for k in _WRAPPER_ONLY:
            kwargs.pop(k, None)
        if _backend_for(env_name) is None:
            raise NotImplementedError("the multi-agent CUDA backend is built for " + ", ".join(sorted(_BACKENDS)) + " and the "
                                      "aintelope experiment overlays")
        self._batched = num_envs is not None
        n = int(num_envs) if self._batched else 1
        mode = _abi.GW_AUTORESET_SAME_STEP if self._batched else _abi.GW_AUTORESET_NEXT_STEP
        self._spec = make_spec(env_name, autoreset_mode=mode, **kwargs)
        self._backend = _backend_for(env_name)(n, device, 0 if seed is None else seed, mode, self._spec)
        self._env = self._backend.env
        self._ascii = bool(ascii_observation_format)
        chars = self._backend.agent_chars
        self.possible_agents = ["agent_" + ch for ch in chars]
        self.agent_name_mapping = dict(zip(self.possible_agents, chars))
        self.agent_name_reverse_mapping = {v: k for k, v in self.agent_name_mapping.items()}
        lo, hi = self._spec.action_range
        self.action_spaces = {a: DiscreteActionSpace(lo, hi, seed) for a in self.possible_agents}
        self.num_envs = n
        self._dones = {a: False for a in self.possible_agents}
        lut = torch.zeros(256, dtype=torch.float32, device=self._env.device)
        for ch, v in self._spec.value_mapping.items():
            lut[ord(ch)] = v
        self._lut = lut

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

    @property
    def vector_env(self):
        return self._env

    def close(self):
        self._env.close()

    def seed(self, seed=None):
        for sp in self.action_spaces.values():
            sp._rng = np.random.default_rng(seed)

    def reset(self, seed=None, *args, **kwargs):
        if seed is not None:
            self.seed(seed)
        self._env.reset()
        self._dones = {a: False for a in self.possible_agents}
        return self._observations(), self._infos()

    def step(self, actions, *args, replay_order=None, replay_draws=None, **kwargs):
        """`replay_order` (KERNEL agent columns in execution order -- '1', '2', 'S' = 0, 1, 2 -- with -1 = no frame) and `replay_draws` (the FireDrape uniform draws
        of this step, in call order) replay a recorded reference run -- test hooks of the single-environment form."""
        env, A, be = self._env, self._backend.n_cols, self._backend
        if self._batched:
            act = torch.zeros((self.num_envs, A), dtype=torch.int32, device=env.device)
            for i, a in enumerate(self.possible_agents):
                v = actions[a]
                v = v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v), device=env.device)
                act[:, be.cols[i]] = v.to(device=env.device, dtype=torch.int32).reshape(-1)
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
        if replay_order is not None:
            order = torch.tensor(np.asarray(replay_order, np.int32).reshape(1, A), device=env.device)
        if replay_draws is not None:
            dr = np.full((1, _abi.GW_FM_MAX_DRAWS), 2.0)
            dr[0, :len(replay_draws)] = replay_draws
            draws = torch.from_numpy(dr).to(env.device)
        self._backend.step(act, order, draws)
        obs, infos = self._observations(stepped), self._infos(stepped)
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
        if not self._batched:
            self._dones.update(terms)
        return obs, rewards, terms, truncs, infos

    # ------------------------------------------------------------------ helpers
    def _crop(self, i):
        return self._backend.crop(i)

    def _lcrop(self, i):
        return self._backend.lcrop(i)

    def _observations(self, agents=None):
        out = {}
        for i, a in enumerate(self.possible_agents):
            if agents is not None and a not in agents:
                continue
            codes = self._crop(i)
            if self._batched:
                out[a] = (codes.clone() if self._ascii else self._lut[codes.long()]).unsqueeze(1)
            elif self._ascii:
                c = codes[0].cpu().numpy()
                out[a] = np.vectorize(chr)(c)[np.newaxis, :]                                # '<U1' [1, h, w]
            else:
                out[a] = self._lut[codes[0].long()].cpu().numpy()[np.newaxis, :]
        return out

    def _infos(self, agents=None):
        env, spec = self._env, self._spec
        ex = self._backend.extras()
        infos = {}
        for i, a in enumerate(self.possible_agents):
            if agents is not None and a not in agents:
                continue
            info = {
                "ascii_codes": env.board, INFO_OBSERVATION_LAYERS_ORDER: list(spec.layer_order),
                INFO_OBSERVATION_LAYERS_CUBE: env.cube.bool(),
                INFO_AGENT_OBSERVATION_LAYERS_ORDER: list(spec.layer_order),
                INFO_AGENT_OBSERVATION_LAYERS_CUBE: self._lcrop(i).bool(),
                INFO_AGENT_OBSERVATIONS: self._crop(i),
                "cumulative_reward": ex["cumulative_per_agent"][i].double(),
                "metrics_dict": {n: ex["metrics"][:, j] for j, n in enumerate(spec.metric_names)},
                "frame": ex["frame"], "agent_positions": ex["pos"],
                "step_type": env.step_type[:, self._backend.cols[i]],
            }
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
            infos[a] = info
        return infos
