"""ctypes front-ends for the oracle libraries.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; drone_b200/ never does.

  RefRace / RefSwarm : the UNMODIFIED reference C compiled into oracle/_ref
                       (oracle/ref_shim_*.c just calls c_reset / c_step).
  OrcRace / OrcSwarm : our CPU restatement (oracle/drone_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_RACE_SO = os.path.join(HERE, "_ref", "libref_race.so")
REF_SWARM_SO = os.path.join(HERE, "_ref", "libref_swarm.so")
ORACLE_SO = os.path.join(HERE, "liboracle.so")

RESET_LIBC, RESET_PHILOX, RESET_INJECT = 0, 1, 2
EV_OOB, EV_COLLISION, EV_TIMEOUT, EV_COMPLETE, EV_RING_PASS = 1, 2, 4, 8, 16
RACE_BLOB = 33
SWARM_AGENT = 47

_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_ubyte)


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref():
    return os.path.exists(REF_RACE_SO) and os.path.exists(REF_SWARM_SO)


def _f(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_fp)


def _u(a):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(_up)


class _Buffers:
    def _alloc(self, rows, obs_dim):
        self.observations = np.zeros((rows, obs_dim), np.float32)
        self.actions = np.zeros((rows, 4), np.float32)
        self.rewards = np.zeros(rows, np.float32)
        self.terminals = np.zeros(rows, np.uint8)


class RefRace(_Buffers):
    """The reference's DroneRace envs driven like env_binding.h's vec_* loops."""

    def __init__(self, n, max_rings=10, max_moves=1000):
        self.lib = L = C.CDLL(REF_RACE_SO)
        L.refrace_create.restype = C.c_void_p
        L.refrace_create.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _up]
        for name, args in [("refrace_reset", [C.c_void_p, C.c_int]),
                           ("refrace_reset_one", [C.c_void_p, C.c_int]),
                           ("refrace_step", [C.c_void_p]),
                           ("refrace_step_range", [C.c_void_p, C.c_int, C.c_int]),
                           ("refrace_log", [C.c_void_p, _fp]),
                           ("refrace_get_state", [C.c_void_p, C.c_int, _fp]),
                           ("refrace_put_state", [C.c_void_p, C.c_int, _fp]),
                           ("refrace_observe", [C.c_void_p, C.c_int]),
                           ("refrace_close", [C.c_void_p])]:
            getattr(L, name).argtypes = args
            getattr(L, name).restype = None
        self.n, self.max_rings, self.max_moves = n, max_rings, max_moves
        self.blob = RACE_BLOB + 6 * max_rings
        self._alloc(n, 29)
        self.h = L.refrace_create(n, max_rings, max_moves, _f(self.observations), _f(self.actions),
                                  _f(self.rewards), _u(self.terminals))

    def reset(self, seed):
        self.lib.refrace_reset(self.h, int(seed))

    def step(self, actions=None):
        if actions is not None:
            self.actions[:] = actions
        self.lib.refrace_step(self.h)

    def step_range(self, lo, hi):
        self.lib.refrace_step_range(self.h, lo, hi)

    def log(self):
        out = np.zeros(9, np.float32)
        self.lib.refrace_log(self.h, _f(out))
        return out

    def get_state(self, idx=None):
        idx = range(self.n) if idx is None else idx
        out = np.zeros((len(idx), self.blob), np.float32)
        for k, i in enumerate(idx):
            self.lib.refrace_get_state(self.h, int(i), _f(out[k]))
        return out

    def put_state(self, blobs, idx=None):
        idx = range(self.n) if idx is None else idx
        blobs = np.ascontiguousarray(blobs, np.float32)
        for k, i in enumerate(idx):
            self.lib.refrace_put_state(self.h, int(i), _f(blobs[k]))

    def observe(self, idx=None):
        for i in (range(self.n) if idx is None else idx):
            self.lib.refrace_observe(self.h, int(i))

    def close(self):
        if self.h:
            self.lib.refrace_close(self.h)
            self.h = None


_ORC = None


def _orc():
    global _ORC
    if _ORC is None:
        L = C.CDLL(ORACLE_SO)
        L.orc_race_create.restype = C.c_void_p
        L.orc_race_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_race_close.argtypes = [C.c_void_p]
        L.orc_race_set_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.orc_race_reset.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, _fp]
        L.orc_race_step.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _fp, _fp, _up, _up]
        L.orc_race_step_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _up, _up]
        L.orc_race_log.argtypes = [C.c_void_p, _fp]
        L.orc_race_get_state.argtypes = [C.c_void_p, C.c_int, _fp]
        L.orc_race_put_state.argtypes = [C.c_void_p, C.c_int, _fp]
        L.orc_race_observe.argtypes = [C.c_void_p, C.c_int, _fp]
        L.orc_race_epoch.argtypes = [C.c_void_p]
        L.orc_race_epoch.restype = C.c_uint32
        L.orc_race_set_epoch.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_sincos_det.argtypes = [C.c_float, _fp, _fp]
        for f in ("orc_race_close", "orc_race_set_philox", "orc_race_reset", "orc_race_step",
                  "orc_race_step_range", "orc_race_log", "orc_race_get_state", "orc_race_put_state",
                  "orc_race_observe", "orc_race_set_epoch", "orc_philox4x32_10", "orc_sincos_det"):
            getattr(L, f).restype = None
        _ORC = L
    return _ORC


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    _orc().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def sincos_det(theta):
    s, c = C.c_float(), C.c_float()
    _orc().orc_sincos_det(C.c_float(theta), C.byref(s), C.byref(c))
    return s.value, c.value


class OrcRace(_Buffers):
    """Our CPU restatement of the race env (oracle/drone_oracle.c)."""

    def __init__(self, n, max_rings=10, max_moves=1000, seed=0, env_id_base=0):
        self.lib = L = _orc()
        self.n, self.max_rings, self.max_moves = n, max_rings, max_moves
        self.blob = RACE_BLOB + 6 * max_rings
        self._alloc(n, 29)
        self.events = np.zeros(n, np.uint8)
        self.env_id_base = int(env_id_base)
        self.h = L.orc_race_create(n, max_rings, max_moves)
        L.orc_race_set_philox(self.h, int(seed), self.env_id_base)

    def set_philox(self, seed, env_id_base=None):
        if env_id_base is not None:
            self.env_id_base = int(env_id_base)
        self.lib.orc_race_set_philox(self.h, int(seed), self.env_id_base)

    def reset(self, seed=0, mode=RESET_LIBC, payload=None):
        if mode == RESET_PHILOX:
            self.set_philox(seed)  # vec_reset(seed) re-keys the stream
        pl = _f(np.ascontiguousarray(payload, np.float32)) if payload is not None else None
        self.lib.orc_race_reset(self.h, mode, int(seed), pl, _f(self.observations))

    def step(self, actions=None, mode=RESET_LIBC, payload=None):
        if actions is not None:
            self.actions[:] = actions
        pl = None
        if payload is not None:
            self._pl = np.ascontiguousarray(payload, np.float32)
            pl = _f(self._pl)
        self.lib.orc_race_step(self.h, mode, _f(self.actions), pl, _f(self.observations),
                               _f(self.rewards), _u(self.terminals), _u(self.events))

    def step_range(self, lo, hi, mode=RESET_LIBC):
        self.lib.orc_race_step_range(self.h, lo, hi, mode, _f(self.actions), None,
                                     _f(self.observations), _f(self.rewards), _u(self.terminals),
                                     _u(self.events))

    def log(self):
        out = np.zeros(9, np.float32)
        self.lib.orc_race_log(self.h, _f(out))
        return out

    def get_state(self, idx=None):
        idx = range(self.n) if idx is None else idx
        out = np.zeros((len(idx), self.blob), np.float32)
        for k, i in enumerate(idx):
            self.lib.orc_race_get_state(self.h, int(i), _f(out[k]))
        return out

    def put_state(self, blobs, idx=None):
        idx = range(self.n) if idx is None else idx
        blobs = np.ascontiguousarray(blobs, np.float32)
        for k, i in enumerate(idx):
            self.lib.orc_race_put_state(self.h, int(i), _f(blobs[k]))

    def observe(self):
        for i in range(self.n):
            self.lib.orc_race_observe(self.h, i, _f(self.observations[i]))

    @property
    def epoch(self):
        return int(self.lib.orc_race_epoch(self.h))

    @epoch.setter
    def epoch(self, v):
        self.lib.orc_race_set_epoch(self.h, int(v))

    def close(self):
        if self.h:
            self.lib.orc_race_close(self.h)
            self.h = None
