"""drone_b200 -- B200-native batched environment step for the tensaur/drone (PufferLib Ocean) envs.

Only the data-parallel hot path lives here: the CUDA step kernels behind the C
ABI of include/b200drone.h (drone_b200/csrc), and the host-side mirror of the
reference's env interface (binding module, DroneRace / DroneSwarm wrappers).
There is no CPU implementation in this package.
"""
__version__ = "0.1.0"
