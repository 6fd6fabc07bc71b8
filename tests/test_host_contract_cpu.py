"""The PufferEnv buffer contract restated in drone_b200/pufferenv.py
(reference: pufferlib/pufferlib.py:22-111, pufferlib/spaces.py:12-25)."""
import numpy as np
import pytest

from drone_b200 import pufferenv as pe


class _Env(pe.PufferEnv):
    def __init__(self, n, buf=None):
        self.single_observation_space = pe.Box(low=-1, high=1, shape=(29,), dtype=np.float32)
        self.single_action_space = pe.Box(low=-1, high=1, shape=(4,), dtype=np.float32)
        self.num_agents = n
        super().__init__(buf)

    def reset(self, seed=None):
        return self.observations, []

    def step(self, actions):
        self.actions[:] = actions
        return self.observations, self.rewards, self.terminals, self.truncations, []


def test_set_buffers_allocates_the_flat_contract():
    env = _Env(6)
    assert env.observations.shape == (6, 29) and env.observations.dtype == np.float32
    assert env.actions.shape == (6, 4) and env.actions.dtype == np.float32
    assert env.rewards.shape == (6,) and env.rewards.dtype == np.float32
    assert env.terminals.dtype == bool and env.truncations.dtype == bool
    assert env.masks.all() and env.masks.shape == (6,)
    assert env.action_space.shape == (6, 4) and env.observation_space.shape == (6, 29)
    assert list(env.agent_ids) == list(range(6))
    assert env.emulated is False and env.done is False and env.driver_env is env


def test_set_buffers_adopts_caller_slices_without_copy():
    n = 4
    buf = dict(observations=np.zeros((n, 29), np.float32), actions=np.zeros((n, 4), np.float32),
               rewards=np.zeros(n, np.float32), terminals=np.zeros(n, bool), truncations=np.zeros(n, bool),
               masks=np.ones(n, bool))
    env = _Env(n, buf)
    for k in buf:
        assert getattr(env, k) is buf[k]


def test_async_api_and_missing_attributes():
    env = _Env(3)
    env.async_reset(0)
    env.send(np.ones((3, 4), np.float32))
    o, r, t, tr, infos, ids, masks = env.recv()
    assert o is env.observations and infos == [] and masks is env.masks
    assert env.actions.sum() == 12

    class Bad(pe.PufferEnv):
        pass
    with pytest.raises(pe.APIUsageError):
        Bad()


def test_box_space_behaviour():
    b = pe.Box(low=-1, high=1, shape=(4,), dtype=np.float32)
    s = b.sample()
    assert s.shape == (4,) and s.dtype == np.float32 and b.contains(s)
    j = pe.joint_space(b, 5)
    assert j.shape == (5, 4) and float(j.low.min()) == -1.0


def test_registry_names_and_native_backend_rules():
    """pufferlib/ocean/environment.py:167-177 and pufferlib/vector.py:618-639 for the drone envs."""
    from drone_b200 import registry
    assert registry.env_creator("puffer_drone_race").__name__ == "DroneRace"
    assert registry.env_creator("puffer_drone_swarm").__name__ == "DroneSwarm"
    with pytest.raises(pe.APIUsageError):
        registry.env_creator("drone_race")
    with pytest.raises(pe.APIUsageError):
        registry.env_creator("puffer_breakout")
    with pytest.raises(pe.APIUsageError):
        registry.make("puffer_drone_race", backend="Multiprocessing")
    with pytest.raises(pe.APIUsageError):
        registry.make("puffer_drone_race", num_envs=2)
    with pytest.raises(pe.APIUsageError):
        registry.make("puffer_drone_race", num_envs=0)
    assert registry.ENV_DEFAULTS["drone_swarm"] == dict(num_envs=16, num_drones=64, max_rings=10)
