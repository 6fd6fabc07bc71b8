from .drone_race import DroneRace  # noqa: F401
