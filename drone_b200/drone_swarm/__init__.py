from .drone_swarm import DroneSwarm  # noqa: F401
