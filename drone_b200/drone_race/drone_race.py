"""DroneRace -- the PufferEnv-shaped wrapper of the ring-race env, stepping on the GPU.

Mirrors pufferlib/ocean/drone_race/drone_race.py:7-75 (same constructor
arguments and defaults, same attributes, same reset/step/close flow through
the `binding` module) with one extra choice, `buffers`:

  buffers="host"   the reference's contract: NumPy observations / actions /
                   rewards / terminals / truncations (pinned), or the caller's
                   `buf=` slices.  step() copies actions H2D, runs the kernel and
                   copies the results D2H before returning.
  buffers="device" the same five buffers as torch CUDA tensors aliasing the
                   memory the kernel reads and writes (zero-copy; DLPack-able).
                   step() only launches; nothing synchronises unless vec_log is due.
"""
import numpy as np

from ..pufferenv import Box, PufferEnv
from . import binding


def _pinned(shape, dtype):
    """NumPy array over page-locked memory when torch+CUDA is there (faster H2D/D2H)."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(bool): torch.bool}[np.dtype(dtype)]
            # the ndarray's base is the tensor (torch sets it in .numpy()), so the pinned allocation lives
            # exactly as long as the array: nothing to keep alive on the side, nothing leaks per env
            return torch.zeros(shape, dtype=tdt, pin_memory=True).numpy()
    except Exception:  # noqa: BLE001 - pinning is an optimisation only
        pass
    return np.zeros(shape, dtype=dtype)


def _bind_stream(env):
    """Device buffers: the kernels run on torch's CURRENT stream of the env's device, the stream the
    caller's action writes and observation reads are ordered on (side streams and graph capture included)."""
    import torch
    binding_mod = env._binding
    binding_mod.vec_set_stream(env.c_envs, torch.cuda.current_stream(env.device).cuda_stream)


class DroneRace(PufferEnv):
    def __init__(self, num_envs=16, render_mode=None, report_interval=1, buf=None, seed=0,
                 max_rings=10, max_moves=1000, buffers="host", device=0, math="fast",
                 env_id_base=0, per_env_init=False, write_clamped_actions=-1):
        self.single_observation_space = Box(low=-1, high=1, shape=(29,), dtype=np.float32)
        self.single_action_space = Box(low=-1, high=1, shape=(4,), dtype=np.float32)
        self.num_agents = num_envs
        self.render_mode = render_mode
        self.report_interval = report_interval
        self.tick = 0
        self.buffers = buffers
        self._binding = binding

        if buffers == "device":
            import torch
            if buf is not None:
                raise ValueError("buf= slices are host memory; use buffers='host'")
            self.device = torch.device("cuda", device)
            self._init_device_buffers(torch)
        elif buffers == "host":
            if buf is None:
                buf = dict(
                    observations=_pinned((num_envs, 29), np.float32),
                    actions=_pinned((num_envs, 4), np.float32),
                    rewards=_pinned((num_envs,), np.float32),
                    terminals=_pinned((num_envs,), bool),
                    truncations=_pinned((num_envs,), bool),
                    masks=np.ones(num_envs, dtype=bool),
                )
            super().__init__(buf)
            self.actions = self.actions.astype(np.float32, copy=False)
        else:
            raise ValueError("buffers must be 'host' or 'device'")

        # -1: the caller-visible NumPy action buffer holds clamp(action, -1, 1) after a step like the
        # reference's (dronelib.h:437); device buffers are left alone unless write_clamped_actions=1
        kwargs = dict(max_rings=max_rings, max_moves=max_moves, write_clamped_actions=write_clamped_actions)
        if per_env_init:
            # the reference's own construction path (drone_race.py:37-51): one env_init per env
            c_envs = []
            for env_num in range(num_envs):
                c_envs.append(binding.env_init(
                    self.observations[env_num:(env_num + 1)],
                    self.actions[env_num:(env_num + 1)],
                    self.rewards[env_num:(env_num + 1)],
                    self.terminals[env_num:(env_num + 1)],
                    self.truncations[env_num:(env_num + 1)],
                    env_num, report_interval=self.report_interval, device=device, math=math,
                    env_id_base=env_id_base, **kwargs))
            self._env_handles = c_envs
            self.c_envs = binding.vectorize(*c_envs)
        else:
            self._env_handles = []
            self.c_envs = binding.vec_init(
                self.observations, self.actions, self.rewards, self.terminals, self.truncations,
                num_envs, seed, report_interval=self.report_interval, device=device, math=math,
                env_id_base=env_id_base, **kwargs)

    def _init_device_buffers(self, torch):
        n = self.num_agents
        dev = self.device
        self.observations = torch.zeros((n, 29), dtype=torch.float32, device=dev)
        self.actions = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros(n, dtype=torch.float32, device=dev)
        self.terminals = torch.zeros(n, dtype=torch.bool, device=dev)
        self.truncations = torch.zeros(n, dtype=torch.bool, device=dev)
        self.masks = torch.ones(n, dtype=torch.bool, device=dev)
        self.agent_ids = torch.arange(n, device=dev)
        self.action_space = Box(low=-1, high=1, shape=(n, 4), dtype=np.float32)
        self.observation_space = Box(low=-1, high=1, shape=(n, 29), dtype=np.float32)

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


def test_performance(timeout=10, atn_cache=1024, num_envs=1000, **kwargs):
    """The reference's own perf loop (drone_race.py:77-92), with reset(seed=0) so it runs."""
    import time
    env = DroneRace(num_envs=num_envs, report_interval=1 << 30, **kwargs)
    env.reset(0)
    tick = 0
    actions = [env.action_space.sample() for _ in range(atn_cache)]
    start = time.time()
    while time.time() - start < timeout:
        env.step(actions[tick % atn_cache])
        tick += 1
    sps = env.num_agents * tick / (time.time() - start)
    print(f"SPS: {sps}")
    env.close()
    return sps


if __name__ == "__main__":
    test_performance()
