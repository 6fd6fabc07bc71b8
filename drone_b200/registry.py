"""PufferLib registration adapter: `puffer_drone_race` / `puffer_drone_swarm` -> the B200 classes.

Mirrors the two entry points PufferLib's CLI and scripts go through:
  pufferlib/ocean/environment.py:119-177   MAKE_FUNCTIONS + env_creator('puffer_<name>')
  pufferlib/vector.py:618-639              make(creator, env_kwargs=..., backend=PufferEnv): the
                                           native path (the env vectorises itself, num_envs == 1)
and the `[env]` defaults of config/ocean/drone_race.ini / drone_swarm.ini.

With pufferlib installed, `install()` points its Ocean registry at these classes so that
`puffer train puffer_drone_race --vec.backend PufferEnv` steps on the GPU unmodified.
"""
from .pufferenv import APIUsageError, PufferEnv

MAKE_FUNCTIONS = {"drone_race": "DroneRace", "drone_swarm": "DroneSwarm"}

# config/ocean/drone_race.ini:11-12, config/ocean/drone_swarm.ini:17-20
ENV_DEFAULTS = {
    "drone_race": dict(num_envs=1024),
    "drone_swarm": dict(num_envs=16, num_drones=64, max_rings=10),
}


def env_creator(name="puffer_drone_race"):
    """pufferlib.ocean.environment.env_creator for the two drone envs."""
    if "puffer_" not in name:
        raise APIUsageError(f"Invalid environment name: {name}")
    short = name.replace("puffer_", "")
    if short not in MAKE_FUNCTIONS:
        raise APIUsageError(f"{name} is not provided by drone_b200 (only {sorted(MAKE_FUNCTIONS)})")
    import importlib
    module = importlib.import_module(f"drone_b200.{short}.{short}")
    return getattr(module, MAKE_FUNCTIONS[short])


def make(env_creator_or_name, env_args=None, env_kwargs=None, backend="PufferEnv", num_envs=1, seed=0, **kwargs):
    """pufferlib.vector.make restricted to the native backend (vector.py:618-639): these envs
    vectorise themselves on the GPU, so `num_envs` here (the number of Python env instances) must
    be 1 and the env-level `num_envs` goes in env_kwargs."""
    if num_envs < 1:
        raise APIUsageError("num_envs must be at least 1")
    if num_envs != int(num_envs):
        raise APIUsageError("num_envs must be an integer")
    name = backend if isinstance(backend, str) else getattr(backend, "__name__", "")
    if name not in ("PufferEnv", "native"):
        raise APIUsageError(f"Invalid backend: {backend}: GPU-resident envs use native vectorization "
                            "(Serial / Multiprocessing fan a CPU env out over processes)")
    if num_envs != 1:
        raise APIUsageError("Native vectorization is for PufferEnvs that handle all per-process vectorization "
                            "internally: pass num_envs inside env_kwargs")
    creator = env_creator(env_creator_or_name) if isinstance(env_creator_or_name, str) else env_creator_or_name
    env_kwargs = dict(env_kwargs or {})
    if isinstance(env_creator_or_name, str):
        short = env_creator_or_name.replace("puffer_", "")
        env_kwargs = {**ENV_DEFAULTS[short], **env_kwargs}
    env_kwargs.setdefault("seed", seed)
    vecenv = creator(*(env_args or []), **env_kwargs)
    if not isinstance(vecenv, PufferEnv):
        raise APIUsageError("Native vectorization requires a native PufferEnv")
    return vecenv


def install():
    """Route pufferlib's own registry to the B200 envs (no-op error if pufferlib is absent)."""
    import importlib
    import sys
    try:
        importlib.import_module("pufferlib.ocean.environment")
    except Exception as e:  # noqa: BLE001
        raise ImportError("pufferlib is not importable here; use drone_b200.registry.make directly") from e
    for short in MAKE_FUNCTIONS:
        sys.modules[f"pufferlib.ocean.{short}.{short}"] = importlib.import_module(f"drone_b200.{short}.{short}")
