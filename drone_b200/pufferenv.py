"""The PufferEnv buffer contract, restated for environments whose step runs on the GPU.

Reference: pufferlib/pufferlib.py:22-111 (`set_buffers`, `PufferEnv`) and
pufferlib/spaces.py:12-25 (`joint_space`).  pufferlib itself (and gymnasium) is
not importable in the build image, so the handful of attributes the trainers
and vectorisers touch are restated here with the same names and meanings; when
gymnasium is present its `Box` is used so the spaces compare equal to the
reference's.
"""
import numpy as np

try:  # pragma: no cover - gymnasium is absent in the build image
    from gymnasium.spaces import Box
except Exception:  # noqa: BLE001
    class Box:
        """Minimal stand-in for gymnasium.spaces.Box (low/high/shape/dtype/sample)."""

        def __init__(self, low, high, shape, dtype=np.float32):
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __eq__(self, other):
            return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                    and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class APIUsageError(RuntimeError):
    """pufferlib.APIUsageError (pufferlib/pufferlib.py)."""


def joint_space(space, n):
    """pufferlib/spaces.py:12-25 for Box spaces."""
    return Box(low=float(np.min(space.low)), high=float(np.max(space.high)),
               shape=(n, *space.shape), dtype=space.dtype)


def set_buffers(env, buf=None):
    """pufferlib/pufferlib.py:22-43: allocate the flat buffers, or adopt the caller's slices."""
    if buf is None:
        obs_space = env.single_observation_space
        env.observations = np.zeros((env.num_agents, *obs_space.shape), dtype=obs_space.dtype)
        env.rewards = np.zeros(env.num_agents, dtype=np.float32)
        env.terminals = np.zeros(env.num_agents, dtype=bool)
        env.truncations = np.zeros(env.num_agents, dtype=bool)
        env.masks = np.ones(env.num_agents, dtype=bool)
        atn_space = joint_space(env.single_action_space, env.num_agents)
        env.actions = np.zeros(atn_space.shape, dtype=atn_space.dtype)
    else:
        env.observations = buf["observations"]
        env.rewards = buf["rewards"]
        env.terminals = buf["terminals"]
        env.truncations = buf["truncations"]
        env.masks = buf["masks"]
        env.actions = buf["actions"]


class PufferEnv:
    """pufferlib/pufferlib.py:45-111."""

    def __init__(self, buf=None):
        for attr in ("single_observation_space", "single_action_space", "num_agents"):
            if not hasattr(self, attr):
                raise APIUsageError(f"Environment missing required attribute {attr}")
        if self.num_agents < 1:
            raise APIUsageError("num_agents must be >= 1")
        set_buffers(self, buf)
        self.action_space = joint_space(self.single_action_space, self.num_agents)
        self.observation_space = joint_space(self.single_observation_space, self.num_agents)
        self.agent_ids = np.arange(self.num_agents)

    @property
    def agent_per_batch(self):
        return self.num_agents

    @property
    def emulated(self):
        return False

    @property
    def done(self):
        return False

    @property
    def driver_env(self):
        return self

    def reset(self, seed=None):
        raise NotImplementedError

    def step(self, actions):
        raise NotImplementedError

    def close(self):
        raise NotImplementedError

    def async_reset(self, seed=None):
        _, self.infos = self.reset(seed)
        assert isinstance(self.infos, list), "PufferEnvs must return info as a list of dicts"

    def send(self, actions):
        _, _, _, _, self.infos = self.step(actions)
        assert isinstance(self.infos, list), "PufferEnvs must return info as a list of dicts"

    def recv(self):
        return (self.observations, self.rewards, self.terminals, self.truncations, self.infos,
                self.agent_ids, self.masks)
