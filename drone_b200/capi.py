"""ctypes view of the C ABI in include/b200drone.h (libb200drone.so).

This is the only way Python reaches the CUDA kernels; there is no CPU path.
Loading fails loudly when the library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C drone_b200/csrc`).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2D_LIBRARY") or os.path.join(HERE, "lib", "libb200drone.so")

B2D_OK, B2D_EINVAL, B2D_ENOMEM, B2D_ECUDA, B2D_ESTATE = 0, -1, -2, -3, -4
MATH_FAST, MATH_STRICT = 0, 1
RESET_PHILOX, RESET_INJECT = 0, 1
MEM_DEVICE, MEM_HOST = 0, 1
POLICY_FP32, POLICY_TF32 = 0, 1
LOG_FIELDS = ("episode_return", "episode_length", "rings_passed", "collision_rate", "oob",
              "timeout", "score", "perf", "n")  # DR/dronelib.h:52-63


class Buffers(C.Structure):
    _fields_ = [("observations", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p),
                ("terminals", C.c_void_p), ("truncations", C.c_void_p), ("location", C.c_int)]


class RaceCfg(C.Structure):
    _fields_ = [("num_envs", C.c_int), ("max_rings", C.c_int), ("max_moves", C.c_int),
                ("device", C.c_int), ("seed", C.c_uint64), ("env_id_base", C.c_uint32),
                ("math", C.c_int), ("write_clamped_actions", C.c_int)]


class SwarmCfg(C.Structure):
    _fields_ = [("num_envs", C.c_int), ("num_agents", C.c_int), ("max_rings", C.c_int),
                ("device", C.c_int), ("seed", C.c_uint64), ("env_id_base", C.c_uint32),
                ("math", C.c_int), ("write_clamped_actions", C.c_int)]


class PolicyWeights(C.Structure):
    _fields_ = [("encoder_weight", C.c_void_p), ("encoder_bias", C.c_void_p), ("decoder_mean_weight", C.c_void_p),
                ("decoder_mean_bias", C.c_void_p), ("decoder_logstd", C.c_void_p), ("value_weight", C.c_void_p),
                ("value_bias", C.c_void_p), ("hidden", C.c_int), ("precision", C.c_int)]


class PolicyIO(C.Structure):
    _fields_ = [("observations", C.c_void_p), ("rewards", C.c_void_p), ("terminals", C.c_void_p),
                ("env_actions", C.c_void_p), ("store_observations", C.c_void_p), ("store_actions", C.c_void_p),
                ("store_logprobs", C.c_void_p), ("store_rewards", C.c_void_p), ("store_terminals", C.c_void_p),
                ("store_values", C.c_void_p), ("rows", C.c_int), ("obs_dim", C.c_int), ("row_id_base", C.c_uint32)]


class RefDrone(C.Structure):
    """Memory layout of the reference's `Drone` (dronelib.h:191-247), 208 bytes."""
    _fields_ = [("pos", C.c_float * 3), ("vel", C.c_float * 3), ("quat", C.c_float * 4), ("omega", C.c_float * 3),
                ("rpms", C.c_float * 4),
                ("mass", C.c_float), ("ixx", C.c_float), ("iyy", C.c_float), ("izz", C.c_float), ("arm_len", C.c_float),
                ("k_thrust", C.c_float), ("k_ang_damp", C.c_float), ("k_drag", C.c_float), ("b_drag", C.c_float),
                ("gravity", C.c_float), ("max_rpm", C.c_float), ("max_vel", C.c_float), ("max_omega", C.c_float),
                ("k_mot", C.c_float), ("j_mot", C.c_float),
                ("spawn_pos", C.c_float * 3), ("prev_pos", C.c_float * 3), ("target_pos", C.c_float * 3),
                ("target_vel", C.c_float * 3),
                ("last_abs_reward", C.c_float), ("last_target_reward", C.c_float), ("last_collision_reward", C.c_float),
                ("episode_return", C.c_float), ("collisions", C.c_float), ("episode_length", C.c_int),
                ("score", C.c_float), ("ring_idx", C.c_int)]


class RefRing(C.Structure):
    """Memory layout of the reference's `Ring` (dronelib.h:161-166), 44 bytes."""
    _fields_ = [("pos", C.c_float * 3), ("orientation", C.c_float * 4), ("normal", C.c_float * 3), ("radius", C.c_float)]


class RolloutStore(C.Structure):
    _fields_ = [("observations", C.c_void_p), ("actions", C.c_void_p), ("logprobs", C.c_void_p), ("rewards", C.c_void_p),
                ("terminals", C.c_void_p), ("values", C.c_void_p)]


# every symbol include/b200drone.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "b2d_race_create": (C.c_int, [C.POINTER(_P), C.POINTER(RaceCfg), C.POINTER(Buffers)]),
    "b2d_swarm_create": (C.c_int, [C.POINTER(_P), C.POINTER(SwarmCfg), C.POINTER(Buffers)]),
    "b2d_vec_close": (C.c_int, [_P]),
    "b2d_vec_reset": (C.c_int, [_P, C.c_uint64, _P]),
    "b2d_vec_step": (C.c_int, [_P, _P]),
    "b2d_vec_step_from": (C.c_int, [_P, _P, _P]),
    "b2d_vec_step_tape": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "b2d_vec_step_host": (C.c_int, [_P, _P]),
    "b2d_vec_step_host_from": (C.c_int, [_P, _P, _P]),
    "b2d_vec_reset_host": (C.c_int, [_P, C.c_uint64, _P]),
    "b2d_vec_log": (C.c_int, [_P, C.POINTER(C.c_float), _P]),
    "b2d_vec_log_begin": (C.c_int, [_P, _P, C.POINTER(_P), C.POINTER(C.c_int)]),
    "b2d_vec_log_end": (C.c_int, [_P, C.POINTER(C.c_float), _P]),
    "b2d_vec_log_reduce": (C.c_int, [_P, C.POINTER(C.c_float), _P, _P]),
    "b2d_log_average": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_float)]),
    "b2d_get_buffers": (C.c_int, [_P, C.POINTER(Buffers)]),
    "b2d_num_agents": (C.c_int, [_P]),
    "b2d_obs_dim": (C.c_int, [_P]),
    "b2d_state_blob_floats": (C.c_int, [_P]),
    "b2d_kernel_launches": (C.c_longlong, [_P]),
    "b2d_step_count": (C.c_int, [_P, C.POINTER(C.c_uint32), _P]),
    "b2d_guard_replays": (C.c_int, [_P, C.POINTER(C.c_ulonglong), _P]),
    "b2d_get_state": (C.c_int, [_P, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_float)]),
    "b2d_put_state": (C.c_int, [_P, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_float)]),
    "b2d_observe": (C.c_int, [_P, _P]),
    "b2d_race_blob_to_ref": (C.c_int, [C.POINTER(C.c_float), C.c_int, C.POINTER(RefDrone), C.POINTER(RefRing), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "b2d_swarm_blob_to_ref": (C.c_int, [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(RefDrone), C.POINTER(RefRing),
                                        C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b2d_export_ref": (C.c_int, [_P, C.c_int, C.POINTER(RefDrone), C.POINTER(RefRing), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_float)]),
    "b2d_set_math": (C.c_int, [_P, C.c_int]),
    "b2d_set_reset_mode": (C.c_int, [_P, C.c_int]),
    "b2d_set_reset_payload": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "b2d_set_step_count": (C.c_int, [_P, C.c_uint32]),
    "b2d_profile_kernels": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "b2d_puff_advantage": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                                     C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _P]),
    "b2d_policy_act": (C.c_int, [C.POINTER(PolicyWeights), C.POINTER(PolicyIO), C.c_uint64, _P, C.c_int, _P]),
    "b2d_race_rollout": (C.c_int, [_P, C.POINTER(PolicyWeights), C.POINTER(RolloutStore), C.c_int, C.c_uint64, _P, C.c_int, _P]),
    "b2d_last_error": (C.c_char_p, []),
    "b2d_version": (C.c_int, []),
}

_LIB = None


class B2DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200drone error {code}: {msg}")
        self.code = code


def lib():
    """Load libb200drone.so (once).  No fallback: a missing library is an error."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA library has not been built and there is no "
                "CPU fallback. Run `python -c 'import __graft_entry__ as g; g.build()'`.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc < 0:
        msg = lib().b2d_last_error().decode("utf-8", "replace")
        if rc == B2D_EINVAL:
            raise ValueError(msg)
        if rc == B2D_ENOMEM:
            raise MemoryError(msg)
        raise B2DError(rc, msg)
    return rc
