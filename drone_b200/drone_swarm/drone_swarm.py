"""DroneSwarm -- the PufferEnv-shaped wrapper of the multi-drone swarm env, stepping on the GPU.

Mirrors pufferlib/ocean/drone_swarm/drone_swarm.py:7-75 (same constructor arguments and
defaults, same attributes, same reset/step/close flow through the `binding` module); see
drone_b200/drone_race/drone_race.py for the `buffers` choice (NumPy contract vs zero-copy
torch CUDA tensors).
"""
import numpy as np

from ..drone_race.drone_race import _bind_stream, _pinned
from ..pufferenv import Box, PufferEnv
from . import binding


class DroneSwarm(PufferEnv):
    def __init__(self, num_envs=16, num_drones=64, max_rings=5, render_mode=None, report_interval=1024,
                 buf=None, seed=0, buffers="host", device=0, math="fast", env_id_base=0, per_env_init=False):
        self.single_observation_space = Box(low=-1, high=1, shape=(41,), dtype=np.float32)
        self.single_action_space = Box(low=-1, high=1, shape=(4,), dtype=np.float32)
        self.num_agents = num_envs * num_drones
        self.render_mode = render_mode
        self.report_interval = report_interval
        self.tick = 0
        self.buffers = buffers
        self._binding = binding
        n = self.num_agents

        if buffers == "device":
            import torch
            if buf is not None:
                raise ValueError("buf= slices are host memory; use buffers='host'")
            dev = self.device = torch.device("cuda", device)
            self.observations = torch.zeros((n, 41), dtype=torch.float32, device=dev)
            self.actions = torch.zeros((n, 4), dtype=torch.float32, device=dev)
            self.rewards = torch.zeros(n, dtype=torch.float32, device=dev)
            self.terminals = torch.zeros(n, dtype=torch.bool, device=dev)
            self.truncations = torch.zeros(n, dtype=torch.bool, device=dev)
            self.masks = torch.ones(n, dtype=torch.bool, device=dev)
            self.agent_ids = torch.arange(n, device=dev)
            self.action_space = Box(low=-1, high=1, shape=(n, 4), dtype=np.float32)
            self.observation_space = Box(low=-1, high=1, shape=(n, 41), dtype=np.float32)
        elif buffers == "host":
            if buf is None:
                buf = dict(observations=_pinned((n, 41), np.float32), actions=_pinned((n, 4), np.float32),
                           rewards=_pinned((n,), np.float32), terminals=_pinned((n,), bool),
                           truncations=_pinned((n,), bool), masks=np.ones(n, dtype=bool))
            super().__init__(buf)
            self.actions = self.actions.astype(np.float32, copy=False)
        else:
            raise ValueError("buffers must be 'host' or 'device'")

        if per_env_init:
            # the reference's own construction path (drone_swarm.py:35-48): one env_init per env
            c_envs = []
            for i in range(num_envs):
                sl = slice(i * num_drones, (i + 1) * num_drones)
                c_envs.append(binding.env_init(self.observations[sl], self.actions[sl], self.rewards[sl],
                                               self.terminals[sl], self.truncations[sl], i,
                                               num_agents=num_drones, max_rings=max_rings, device=device, math=math,
                                               env_id_base=env_id_base))
            self._env_handles = c_envs
            self.c_envs = binding.vectorize(*c_envs)
        else:
            self._env_handles = []
            self.c_envs = binding.vec_init(self.observations, self.actions, self.rewards, self.terminals,
                                           self.truncations, num_envs, seed, num_agents=num_drones,
                                           max_rings=max_rings, device=device, math=math, env_id_base=env_id_base)

    def reset(self, seed=None):
        self.tick = 0
        if self.buffers == "device":
            _bind_stream(self)
        binding.vec_reset(self.c_envs, seed)
        return self.observations, []

    def step(self, actions):
        if self.buffers == "device":
            _bind_stream(self)
            if actions is not self.actions:
                self.actions.copy_(actions)
        self.tick += 1
        if self.buffers == "device":
            binding.vec_step(self.c_envs)
        elif (isinstance(actions, np.ndarray) and actions.dtype == np.float32 and actions.flags.c_contiguous
                and actions.shape == self.actions.shape):
            # `self.actions[:] = actions; vec_step` in one call (the copy runs on several cores)
            binding.vec_step_actions(self.c_envs, actions)
        else:
            self.actions[:] = actions
            binding.vec_step(self.c_envs)

        info = []
        if self.tick % self.report_interval == 0:
            log_data = binding.vec_log(self.c_envs)
            if log_data:
                info.append(log_data)

        return (self.observations, self.rewards, self.terminals, self.truncations, info)

    def pinned_actions(self):
        """A page-locked float32 action array of this env's shape.  `step()` uploads such an array by DMA from where
        it is (no staging copy on the CPU); any other NumPy array works too and is copied into `self.actions` first,
        like the reference's wrapper does."""
        return _pinned(tuple(self.actions.shape), np.float32)

    def render(self):
        binding.vec_render(self.c_envs, 0)

    def close(self):
        binding.vec_close(self.c_envs)
        for h in self._env_handles:
            binding.env_close(h)
        self._env_handles = []
