"""Handle-level Python API over the C ABI: device-resident vectorised drone envs.

`RaceVec` / `SwarmVec` own one `b2d_vec` handle each and expose the contract
buffers as torch CUDA tensors that alias the memory the kernels read and write
(zero-copy; `torch.utils.dlpack.to_dlpack(vec.observations)` hands the same
memory to any DLPack consumer).  The PufferLib-shaped wrappers in
drone_race.py / drone_swarm.py sit on top of these.
"""
import ctypes as C

import numpy as np
import torch

from . import capi


def _stream_ptr(stream=None, device=None):
    if stream is None:
        stream = torch.cuda.current_stream(device)
    return C.c_void_p(stream.cuda_stream)


class _Vec:
    obs_dim = 0

    def _adopt(self, handle, device):
        self.h = handle
        self.device = device
        L = capi.lib()
        self.num_agents = L.b2d_num_agents(self.h)
        self.blob_floats = L.b2d_state_blob_floats(self.h)
        self.payload_floats = self.blob_floats

    # ---- the hot path -----------------------------------------------------
    def reset(self, seed=0, stream=None):
        capi.check(capi.lib().b2d_vec_reset(self.h, int(seed) & (2**64 - 1), _stream_ptr(stream, self.device)))

    def step(self, actions=None, stream=None):
        """One env step on the current stream (async).  `actions`: optional CUDA float32
        tensor [num_agents, 4] to read instead of `self.actions` (no copy)."""
        L = capi.lib()
        if actions is None:
            capi.check(L.b2d_vec_step(self.h, _stream_ptr(stream, self.device)))
        else:
            if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous():
                raise ValueError("actions must be a contiguous float32 CUDA tensor")
            if actions.numel() != self.num_agents * 4:
                raise ValueError("actions must have shape [num_agents, 4]")
            capi.check(L.b2d_vec_step_from(self.h, C.c_void_p(actions.data_ptr()), _stream_ptr(stream, self.device)))

    def step_tape(self, tape, first, steps, stream=None):
        """`steps` consecutive steps reading actions from a CUDA tape [T, num_agents, 4]
        (slice (first + k) % T at step k); one kernel launch per step, no host work in between."""
        if tape.dtype != torch.float32 or not tape.is_cuda or not tape.is_contiguous() or tape.dim() != 3:
            raise ValueError("tape must be a contiguous float32 CUDA tensor [T, num_agents, 4]")
        if tape.shape[1] * tape.shape[2] != self.num_agents * 4:
            raise ValueError("tape slices must have shape [num_agents, 4]")
        capi.check(capi.lib().b2d_vec_step_tape(self.h, C.c_void_p(tape.data_ptr()), int(tape.shape[0]), int(first),
                                                int(steps), _stream_ptr(stream, self.device)))

    def step_host(self, stream=None):
        capi.check(capi.lib().b2d_vec_step_host(self.h, _stream_ptr(stream, self.device)))

    def reset_host(self, seed=0, stream=None):
        capi.check(capi.lib().b2d_vec_reset_host(self.h, int(seed) & (2**64 - 1), _stream_ptr(stream, self.device)))

    # ---- statistics -------------------------------------------------------------
    def log(self, stream=None, group=None):
        """vec_log: dict of averaged episode statistics ({} when nothing finished).
        With a torch.distributed `group` (or an initialised default group and
        group=True) the integer sums are all-reduced over NCCL first."""
        L = capi.lib()
        out = (C.c_float * 9)()
        if group is None:
            capi.check(L.b2d_vec_log(self.h, out, _stream_ptr(stream, self.device)))
        else:
            from .shard import reduce_log_sums
            ptr, cnt = C.c_void_p(), C.c_int()
            capi.check(L.b2d_vec_log_begin(self.h, _stream_ptr(stream, self.device), C.byref(ptr), C.byref(cnt)))
            sums = _alias(ptr.value, (cnt.value,), torch.int64, self.device, self)
            reduce_log_sums(sums, None if group is True else group)
            capi.check(L.b2d_vec_log_end(self.h, out, _stream_ptr(stream, self.device)))
        vals = [float(x) for x in out]
        if vals[8] == 0.0:
            return {}
        return self._log_dict(vals)

    # ---- state hooks ---------------------------------------------------------------
    def get_state(self, env_ids=None):
        n = self.num_envs if env_ids is None else len(env_ids)
        out = np.zeros((n, self.blob_floats), np.float32)
        ids = None if env_ids is None else (C.c_int * n)(*[int(i) for i in env_ids])
        capi.check(capi.lib().b2d_get_state(self.h, ids, n, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def put_state(self, blobs, env_ids=None):
        blobs = np.ascontiguousarray(blobs, np.float32)
        n = blobs.shape[0]
        ids = None if env_ids is None else (C.c_int * n)(*[int(i) for i in env_ids])
        capi.check(capi.lib().b2d_put_state(self.h, ids, n, blobs.ctypes.data_as(C.POINTER(C.c_float))))

    def observe(self, stream=None):
        capi.check(capi.lib().b2d_observe(self.h, _stream_ptr(stream, self.device)))

    def set_math(self, math):
        capi.check(capi.lib().b2d_set_math(self.h, _math(math)))

    def set_reset_mode(self, mode):
        capi.check(capi.lib().b2d_set_reset_mode(self.h, int(mode)))

    def set_reset_payload(self, payload):
        payload = np.ascontiguousarray(payload, np.float32)
        assert payload.shape == (self.num_envs, self.payload_floats)
        capi.check(capi.lib().b2d_set_reset_payload(self.h, payload.ctypes.data_as(C.POINTER(C.c_float))))

    @property
    def step_count(self):
        v = C.c_uint32()
        capi.check(capi.lib().b2d_step_count(self.h, C.byref(v), _stream_ptr(None, self.device)))
        return int(v.value)

    @step_count.setter
    def step_count(self, v):
        capi.check(capi.lib().b2d_set_step_count(self.h, int(v)))

    @property
    def guard_replays(self):
        """Steps the fast kernel re-did in the reference's arithmetic (near-threshold guard)."""
        v = C.c_ulonglong()
        capi.check(capi.lib().b2d_guard_replays(self.h, C.byref(v), _stream_ptr(None, self.device)))
        return int(v.value)

    def profile_kernels(self, enable):
        """Start (True) or stop (False) per-kernel CUDA-event timing; stopping returns
        {'step_us', 'adopt_us', 'steps'} averaged over the steps issued in between."""
        out = (C.c_float * 3)()
        capi.check(capi.lib().b2d_profile_kernels(self.h, int(bool(enable)), out))
        if not enable:
            return {"step_us": float(out[0]), "adopt_us": float(out[1]), "steps": int(out[2])}
        return None

    @property
    def kernel_launches(self):
        return int(capi.lib().b2d_kernel_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            capi.check(capi.lib().b2d_vec_close(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _math(m):
    if isinstance(m, str):
        return {"fast": capi.MATH_FAST, "strict": capi.MATH_STRICT}[m]
    return int(m)


def _alias(ptr, shape, dtype, device, owner):
    """torch tensor aliasing raw device memory (kept alive by `owner`)."""
    n = int(np.prod(shape))
    itemsize = torch.empty((), dtype=dtype).element_size()

    class _Cai:
        pass

    holder = _Cai()
    holder.__cuda_array_interface__ = {
        "shape": (n * itemsize,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    holder._owner = owner
    t = torch.as_tensor(holder, device=device)
    return t.view(dtype).view(*shape)


class RaceVec(_Vec):
    """Device-resident DroneRace envs (reference: pufferlib/ocean/drone_race)."""
    obs_dim = 29

    def __init__(self, num_envs, max_rings=10, max_moves=1000, seed=0, device="cuda:0", math="fast",
                 env_id_base=0, write_clamped_actions=False, host_buffers=None):
        if not torch.cuda.is_available():
            raise RuntimeError("drone_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.num_envs = int(num_envs)
        self.max_rings, self.max_moves = int(max_rings), int(max_moves)
        self.row_id_base = int(env_id_base)  # global id of row 0 (policy noise stream, drone_b200.policy)
        cfg = capi.RaceCfg(self.num_envs, self.max_rings, self.max_moves, idx, int(seed) & (2**64 - 1),
                           int(env_id_base), _math(math), int(bool(write_clamped_actions)))
        n = self.num_envs
        bufs = capi.Buffers()
        if host_buffers is None:
            # torch owns the contract buffers; the library adopts the pointers (zero-copy)
            self.observations = torch.zeros((n, self.obs_dim), dtype=torch.float32, device=self.device)
            self.actions = torch.zeros((n, 4), dtype=torch.float32, device=self.device)
            self.rewards = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.terminals = torch.zeros(n, dtype=torch.uint8, device=self.device)
            self.truncations = torch.zeros(n, dtype=torch.uint8, device=self.device)
            bufs.observations = self.observations.data_ptr()
            bufs.actions = self.actions.data_ptr()
            bufs.rewards = self.rewards.data_ptr()
            bufs.terminals = self.terminals.data_ptr()
            bufs.truncations = self.truncations.data_ptr()
            bufs.location = capi.MEM_DEVICE
        else:
            self.host = host_buffers  # dict of numpy arrays (kept alive here)
            bufs.observations = host_buffers["observations"].ctypes.data
            bufs.actions = host_buffers["actions"].ctypes.data
            bufs.rewards = host_buffers["rewards"].ctypes.data
            bufs.terminals = host_buffers["terminals"].ctypes.data
            bufs.truncations = host_buffers["truncations"].ctypes.data
            bufs.location = capi.MEM_HOST
        h = C.c_void_p()
        with torch.cuda.device(idx):
            torch.cuda.current_stream().synchronize()
            capi.check(capi.lib().b2d_race_create(C.byref(h), C.byref(cfg), C.byref(bufs)))
        self._adopt(h, self.device)
        if host_buffers is not None:
            db = capi.Buffers()
            capi.check(capi.lib().b2d_get_buffers(self.h, C.byref(db)))
            self.observations = _alias(db.observations, (n, self.obs_dim), torch.float32, self.device, self)
            self.actions = _alias(db.actions, (n, 4), torch.float32, self.device, self)
            self.rewards = _alias(db.rewards, (n,), torch.float32, self.device, self)
            self.terminals = _alias(db.terminals, (n,), torch.uint8, self.device, self)
            self.truncations = _alias(db.truncations, (n,), torch.uint8, self.device, self)

    def _log_dict(self, v):
        # keys and order of DR/binding.c:13-23
        return {"perf": v[7], "score": v[6], "collision_rate": v[3], "oob": v[4], "timeout": v[5],
                "episode_return": v[0], "episode_length": v[1], "n": v[8]}


def _make_buffers(vec, rows, obs_dim, host_buffers):
    """Contract buffers: torch-owned device tensors adopted by the library (zero-copy), or the
    caller's NumPy arrays mirrored by the *_host entry points."""
    bufs = capi.Buffers()
    if host_buffers is None:
        vec.observations = torch.zeros((rows, obs_dim), dtype=torch.float32, device=vec.device)
        vec.actions = torch.zeros((rows, 4), dtype=torch.float32, device=vec.device)
        vec.rewards = torch.zeros(rows, dtype=torch.float32, device=vec.device)
        vec.terminals = torch.zeros(rows, dtype=torch.uint8, device=vec.device)
        vec.truncations = torch.zeros(rows, dtype=torch.uint8, device=vec.device)
        for name in ("observations", "actions", "rewards", "terminals", "truncations"):
            setattr(bufs, name, getattr(vec, name).data_ptr())
        bufs.location = capi.MEM_DEVICE
    else:
        vec.host = host_buffers
        for name in ("observations", "actions", "rewards", "terminals", "truncations"):
            setattr(bufs, name, host_buffers[name].ctypes.data)
        bufs.location = capi.MEM_HOST
    return bufs


class SwarmVec(_Vec):
    """Device-resident DroneSwarm envs (reference: pufferlib/ocean/drone_swarm)."""
    obs_dim = 41

    def __init__(self, num_envs, num_drones=64, max_rings=5, seed=0, device="cuda:0", math="fast",
                 env_id_base=0, write_clamped_actions=False, host_buffers=None):
        if not torch.cuda.is_available():
            raise RuntimeError("drone_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.num_envs, self.num_drones, self.max_rings = int(num_envs), int(num_drones), int(max_rings)
        rows = self.num_envs * self.num_drones
        self.row_id_base = int(env_id_base) * self.num_drones
        cfg = capi.SwarmCfg(self.num_envs, self.num_drones, self.max_rings, idx, int(seed) & (2**64 - 1),
                            int(env_id_base), _math(math), int(bool(write_clamped_actions)))
        bufs = _make_buffers(self, rows, self.obs_dim, host_buffers)
        h = C.c_void_p()
        with torch.cuda.device(idx):
            torch.cuda.current_stream().synchronize()
            capi.check(capi.lib().b2d_swarm_create(C.byref(h), C.byref(cfg), C.byref(bufs)))
        self._adopt(h, self.device)
        self.payload_floats = self.num_drones * 41 + 2 + 6 * self.max_rings
        if host_buffers is not None:
            db = capi.Buffers()
            capi.check(capi.lib().b2d_get_buffers(self.h, C.byref(db)))
            self.observations = _alias(db.observations, (rows, self.obs_dim), torch.float32, self.device, self)
            self.actions = _alias(db.actions, (rows, 4), torch.float32, self.device, self)
            self.rewards = _alias(db.rewards, (rows,), torch.float32, self.device, self)
            self.terminals = _alias(db.terminals, (rows,), torch.uint8, self.device, self)
            self.truncations = _alias(db.truncations, (rows,), torch.uint8, self.device, self)

    def split_state(self, blobs):
        """state blobs [num_envs, blob] -> (env blobs [n, 2+6R], agent blobs [n, A, 47])"""
        n, A = blobs.shape[0], self.num_drones
        return blobs[:, A * 47:].copy(), blobs[:, :A * 47].reshape(n, A, 47).copy()

    def join_state(self, env, ag):
        return np.concatenate([np.asarray(ag, np.float32).reshape(len(env), -1), np.asarray(env, np.float32)], axis=1)

    def _log_dict(self, v):
        # keys and order of DS/binding.c:13-23
        return {"perf": v[7], "score": v[6], "rings_passed": v[2], "collision_rate": v[3], "oob": v[4],
                "episode_return": v[0], "episode_length": v[1], "n": v[8]}
