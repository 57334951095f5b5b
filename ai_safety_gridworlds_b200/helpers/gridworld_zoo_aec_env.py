"""GridworldZooAecEnv: the reference's PettingZoo AEC (agent-by-agent) signature over the CUDA backend.

Mirrors helpers/gridworld_zoo_aec_env.py of the reference (constructor :100-262, `agent_iter`
:327-333, `observe` :380-411, `last` :503-518, `step` :607-806, `reset` :809-905) for
firemaker_ex_ma and island_navigation_ex_ma.  One AEC step is ONE single-agent `EnvironmentMa.step({agent: action})`
(rl/pycolab_interface_ma.py:173-246), i.e. one engine frame: the wrapper launches the
kernel with the execution order {agent, -1, ...} (include/gwsim_fm.h, include/gwsim_ima.h).

  * `num_envs=None` is the drop-in single environment with the reference's bookkeeping: `rewards`
    holds the frame's reward vector of every agent, `_cumulative_rewards[agent]` restarts at the
    agent's own step and accumulates over the other agents' steps (the value `last()` returns), a
    finished agent stays selected once more and must be stepped with `None` (its "dead step",
    :623-648), after which it leaves `agents`.  After the frame cut-off the next live agent's step starts a
    new game (rl/pycolab_interface_ma.py:206-213).  Deviation: the reference then fails with a
    KeyError as soon as a live agent steps after another one was removed (:757-759 credit the
    removed agent); here removed agents are skipped.
  * `num_envs=N` is the batched form: torch CUDA tensors with a leading batch dimension, every
    environment stepped by the same agent, auto-reset inside the frame that ends a game, no agent
    removal (`agent_selection` cycles '1', '2', 'S').

pettingzoo itself is not imported (duck-typed).  There is no CPU fallback.
"""
import numpy as np
import torch

from .. import _abi
from .gridworld_zoo_parallel_env import GridworldZooParallelEnv


class GridworldZooAecEnv(object):
    metadata = {"render.modes": ["ansi"], "name": "gridworld_zoo_aec_env_b200"}

    class ZooAECAgentIter(object):                     # :362-377
        def __init__(self, env, max_iter):
            self.env, self.num_iterations, self.max_iter = env, 0, max_iter

        def __iter__(self):
            return self

        def __next__(self):
            if self.num_iterations < self.max_iter and not self.env._all_agents_done:
                self.num_iterations += 1
                return self.env._next_agent
            raise StopIteration

    def __init__(self, env_name, ascii_observation_format=True, seed=None, num_envs=None, device=None, **kwargs):
        self._par = GridworldZooParallelEnv(env_name, ascii_observation_format=ascii_observation_format, seed=seed,
                                            num_envs=num_envs, device=device, final_info=False, **kwargs)
        self._batched = num_envs is not None
        self.possible_agents = list(self._par.possible_agents)
        self.agent_name_mapping = dict(self._par.agent_name_mapping)
        self.agent_name_reverse_mapping = dict(self._par.agent_name_reverse_mapping)
        self.num_envs = self._par.num_envs
        self._init_bookkeeping()

    # ------------------------------------------------------------------ bookkeeping
    def _zero_reward(self, i):
        env = self._par._env
        if self._batched:
            return torch.zeros((self.num_envs, self._par._backend.reward(i).shape[-1]), dtype=torch.float64, device=env.device)
        return 0.0                                      # the reference initialises scalars (:899-901)

    def _init_bookkeeping(self):
        self._agents = list(self.possible_agents)
        self._next_agent, self._next_agent_index, self._all_agents_done = self.possible_agents[0], 0, False
        self._rewards = {a: self._zero_reward(i) for i, a in enumerate(self.possible_agents)}
        self._cumulative_rewards = {a: self._zero_reward(i) for i, a in enumerate(self.possible_agents)}
        if self._batched:
            z = torch.zeros((self.num_envs,), dtype=torch.bool, device=self._par._env.device)
            self.terminations = {a: z.clone() for a in self.possible_agents}
            self.truncations = {a: z.clone() for a in self.possible_agents}
        else:
            self.terminations = {a: False for a in self.possible_agents}
            self.truncations = {a: False for a in self.possible_agents}
        self._infos = self._par._infos()

    def _move_to_next_agent(self):                      # :336-357
        for _ in range(len(self.possible_agents)):
            self._next_agent_index = (self._next_agent_index + 1) % len(self.possible_agents)
            agent = self.possible_agents[self._next_agent_index]
            if agent in self._agents:
                self._next_agent = agent
                return
        self._next_agent_index, self._next_agent, self._all_agents_done = -1, None, True

    # ------------------------------------------------------------------ PettingZoo surface
    @property
    def agents(self):
        return self._agents

    @property
    def num_agents(self):
        return len(self._agents)

    @property
    def max_num_agents(self):
        return len(self.possible_agents)

    @property
    def agent_selection(self):
        return self._next_agent

    @property
    def rewards(self):
        return self._rewards

    @property
    def infos(self):
        return self._infos

    @property
    def vector_env(self):
        return self._par._env

    @property
    def state(self):
        board = self._par._env.board
        return board.clone().unsqueeze(1) if self._batched else board.cpu().numpy().copy()

    def action_space(self, agent):
        return self._par.action_spaces[agent]

    @property
    def action_spaces(self):
        return self._par.action_spaces

    def agent_iter(self, max_iter=2 ** 63):
        return GridworldZooAecEnv.ZooAECAgentIter(self, max_iter)

    def close(self):
        self._par.close()

    def seed(self, seed=None):
        self._par.seed(seed)

    def get_step_no(self):
        f = self._par._env.observe()["frame"]
        return f if self._batched else int(f[0].item())

    def reset(self, seed=None, *args, **kwargs):
        if seed is not None:
            self.seed(seed)
        self._par._env.reset()
        self._par._dones = {a: False for a in self.possible_agents}
        self._init_bookkeeping()

    def observe(self, agent):
        """The agent's CURRENT perspective, whichever agent moved last (:380-411)."""
        return self._par._observations()[agent]

    def observe_info(self, agent):
        return self._par._infos()[agent]

    def last_for_agent(self, agent=None, observe=True):
        agent = self._next_agent if agent is None else agent
        state = self.observe(agent) if observe else None
        return (state, self._cumulative_rewards[agent], self.terminations[agent], self.truncations[agent], self._infos[agent])

    def last(self, observe=True):
        return self.last_for_agent(self._next_agent, observe)

    def step(self, action, *args, replay_draws=None, **kwargs):
        """`replay_draws` (the FireDrape uniform draws of this frame, in call order) replays a recorded reference
        run -- test hook of the single-environment form."""
        sel = self._next_agent
        if sel is None:
            raise ValueError("all agents are done: call reset()")
        idx = self.possible_agents.index(sel)
        env = self._par._env
        if not self._batched and (self.terminations[sel] or self.truncations[sel]):
            step_action = action["step"] if isinstance(action, dict) else action
            if step_action is not None:
                raise ValueError("When an agent is dead, the only valid action is None")
            del self.terminations[sel], self.truncations[sel], self._cumulative_rewards[sel], self._infos[sel]
            self._agents.remove(sel)
            self._rewards = {a: 0.0 for a in self._agents}
            self._move_to_next_agent()
            return
        A, col = self._par._backend.n_cols, self._par._backend.cols[idx]       # the kernel's agent columns
        order = torch.tensor([[col] + [-1] * (A - 1)], dtype=torch.int32, device=env.device).expand(self.num_envs, A).contiguous()
        if self._batched:
            v = action if torch.is_tensor(action) else torch.as_tensor(np.asarray(action), device=env.device)
            act = torch.zeros((self.num_envs, A), dtype=torch.int32, device=env.device)
            act[:, col] = v.to(device=env.device, dtype=torch.int32).reshape(-1)
            draws = None
        else:
            v = action["step"] if isinstance(action, dict) else action
            if v is None:
                raise ValueError("None is only valid as the action of a dead agent")
            v = int(np.asarray(v).item())
            if v == 9:
                raise NotImplementedError("QUIT is not supported by the multi-agent CUDA backend")
            row = [0] * A
            row[col] = v
            act = torch.tensor([row], dtype=torch.int32, device=env.device)
            draws = None
            if replay_draws is not None:
                dr = np.full((1, _abi.GW_FM_MAX_DRAWS), 2.0)
                dr[0, :len(replay_draws)] = replay_draws
                draws = torch.from_numpy(dr).to(env.device)
        self._par._backend.step(act, order, draws)
        self._infos[sel] = self._par._infos()[sel]
        rewards = {}
        for i, a in enumerate(self.possible_agents):
            r = self._par._backend.reward(i).double()
            rewards[a] = r.clone() if self._batched else r[0].cpu().numpy()
        self._cumulative_rewards[sel] = self._zero_reward(idx)                      # :757
        for a, r in rewards.items():
            if a in self._cumulative_rewards:
                self._cumulative_rewards[a] = self._cumulative_rewards[a] + r
        done = env.step_type[:, col] == 2                                           # StepType.LAST (:763)
        if self._batched:
            self._rewards.update(rewards)
            self.terminations[sel] = done
        else:
            done = bool(done[0].item())
            self._rewards.update({a: r for a, r in rewards.items() if a in self._agents})
            for a in self._agents:                                                  # :777-780
                if self.terminations[a] or self.truncations[a]:
                    self._rewards[a] = 0.0
            self.terminations[sel] = done
            self._par._dones[sel] = False               # the single-agent frames are gated here, not by the parallel wrapper
        self._move_to_next_agent()
