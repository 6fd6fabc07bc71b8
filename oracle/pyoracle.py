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
REF_ADVANTAGE_SO = os.path.join(HERE, "_ref", "libref_advantage.so")
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


SWARM_AGENT_PAYLOAD = 41


class RefSwarm(_Buffers):
    """The reference's DroneSwarm envs (oracle/_ref/libref_swarm.so), driven like env_binding.h."""

    def __init__(self, n, num_agents, max_rings=5):
        self.lib = L = C.CDLL(REF_SWARM_SO)
        L.refswarm_create.restype = C.c_void_p
        L.refswarm_create.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _up]
        for name, args in [("refswarm_reset", [C.c_void_p, C.c_int]), ("refswarm_step", [C.c_void_p]),
                           ("refswarm_step_range", [C.c_void_p, C.c_int, C.c_int]),
                           ("refswarm_log", [C.c_void_p, _fp]),
                           ("refswarm_get_env", [C.c_void_p, C.c_int, _fp]),
                           ("refswarm_put_env", [C.c_void_p, C.c_int, _fp]),
                           ("refswarm_get_agent", [C.c_void_p, C.c_int, C.c_int, _fp]),
                           ("refswarm_put_agent", [C.c_void_p, C.c_int, C.c_int, _fp]),
                           ("refswarm_observe", [C.c_void_p, C.c_int]), ("refswarm_close", [C.c_void_p])]:
            getattr(L, name).argtypes = args
            getattr(L, name).restype = None
        self.n, self.A, self.max_rings = n, num_agents, max_rings
        self.env_blob = 2 + 6 * max_rings
        self._alloc(n * num_agents, 41)
        self.h = L.refswarm_create(n, num_agents, max_rings, _f(self.observations), _f(self.actions),
                                   _f(self.rewards), _u(self.terminals))

    def reset(self, seed):
        self.lib.refswarm_reset(self.h, int(seed))

    def step(self, actions=None):
        if actions is not None:
            self.actions[:] = actions
        self.lib.refswarm_step(self.h)

    def log(self):
        out = np.zeros(9, np.float32)
        self.lib.refswarm_log(self.h, _f(out))
        return out

    def get_state(self):
        """(env blobs [n, 2+6R], agent blobs [n, A, 47])"""
        env = np.zeros((self.n, self.env_blob), np.float32)
        ag = np.zeros((self.n, self.A, SWARM_AGENT), np.float32)
        for e in range(self.n):
            self.lib.refswarm_get_env(self.h, e, _f(env[e]))
            for a in range(self.A):
                self.lib.refswarm_get_agent(self.h, e, a, _f(ag[e, a]))
        return env, ag

    def put_state(self, env, ag):
        env = np.ascontiguousarray(env, np.float32)
        ag = np.ascontiguousarray(ag, np.float32)
        for e in range(self.n):
            self.lib.refswarm_put_env(self.h, e, _f(env[e]))
            for a in range(self.A):
                self.lib.refswarm_put_agent(self.h, e, a, _f(ag[e, a]))

    def observe(self):
        for e in range(self.n):
            self.lib.refswarm_observe(self.h, e)

    def close(self):
        if self.h:
            self.lib.refswarm_close(self.h)
            self.h = None


def _orc_swarm():
    L = _orc()
    if not getattr(L, "_swarm_bound", False):
        L.orc_swarm_create.restype = C.c_void_p
        L.orc_swarm_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_swarm_close.argtypes = [C.c_void_p]
        L.orc_swarm_set_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.orc_swarm_set_payload.argtypes = [C.c_void_p, _fp, _up, _fp, _up]
        L.orc_swarm_reset.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp]
        L.orc_swarm_step.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _fp, _up]
        L.orc_swarm_observe.argtypes = [C.c_void_p, C.c_int, _fp]
        L.orc_swarm_log.argtypes = [C.c_void_p, _fp]
        L.orc_swarm_get_env.argtypes = [C.c_void_p, C.c_int, _fp]
        L.orc_swarm_put_env.argtypes = [C.c_void_p, C.c_int, _fp]
        L.orc_swarm_get_agent.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp]
        L.orc_swarm_put_agent.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp]
        L.orc_swarm_formation_target.argtypes = [C.c_int, C.c_int, C.c_int, _fp]
        for f in ("orc_swarm_close", "orc_swarm_set_philox", "orc_swarm_set_payload", "orc_swarm_reset",
                  "orc_swarm_step", "orc_swarm_observe", "orc_swarm_log", "orc_swarm_get_env", "orc_swarm_put_env",
                  "orc_swarm_get_agent", "orc_swarm_put_agent", "orc_swarm_formation_target"):
            getattr(L, f).restype = None
        L._swarm_bound = True
    return L


class OrcSwarm(_Buffers):
    """Our CPU restatement of the swarm env (oracle/drone_oracle.c)."""

    def __init__(self, n, num_agents, max_rings=5, seed=0, env_id_base=0):
        self.lib = L = _orc_swarm()
        self.n, self.A, self.max_rings = n, num_agents, max_rings
        self.env_blob = 2 + 6 * max_rings
        self._alloc(n * num_agents, 41)
        self.h = L.orc_swarm_create(n, num_agents, max_rings)
        self.env_id_base = int(env_id_base)
        L.orc_swarm_set_philox(self.h, int(seed), self.env_id_base)
        # draw results of the latest reset/step (recorded in LIBC/PHILOX mode, consumed in INJECT mode)
        self.pay_agent = np.zeros((n * num_agents, SWARM_AGENT_PAYLOAD), np.float32)
        self.flag_agent = np.zeros(n * num_agents, np.uint8)
        self.pay_env = np.zeros((n, self.env_blob), np.float32)
        self.flag_env = np.zeros(n, np.uint8)
        L.orc_swarm_set_payload(self.h, _f(self.pay_agent), _u(self.flag_agent), _f(self.pay_env), _u(self.flag_env))

    def reset(self, seed=0, mode=RESET_LIBC):
        if mode == RESET_PHILOX:
            self.lib.orc_swarm_set_philox(self.h, int(seed), self.env_id_base)
        self.lib.orc_swarm_reset(self.h, mode, int(seed), _f(self.observations))

    def step(self, actions=None, mode=RESET_LIBC):
        if actions is not None:
            self.actions[:] = actions
        self.lib.orc_swarm_step(self.h, mode, _f(self.actions), _f(self.observations), _f(self.rewards),
                                _u(self.terminals))

    def log(self):
        out = np.zeros(9, np.float32)
        self.lib.orc_swarm_log(self.h, _f(out))
        return out

    def get_state(self):
        env = np.zeros((self.n, self.env_blob), np.float32)
        ag = np.zeros((self.n, self.A, SWARM_AGENT), np.float32)
        for e in range(self.n):
            self.lib.orc_swarm_get_env(self.h, e, _f(env[e]))
            for a in range(self.A):
                self.lib.orc_swarm_get_agent(self.h, e, a, _f(ag[e, a]))
        return env, ag

    def put_state(self, env, ag):
        env = np.ascontiguousarray(env, np.float32)
        ag = np.ascontiguousarray(ag, np.float32)
        for e in range(self.n):
            self.lib.orc_swarm_put_env(self.h, e, _f(env[e]))
            for a in range(self.A):
                self.lib.orc_swarm_put_agent(self.h, e, a, _f(ag[e, a]))

    def observe(self):
        for e in range(self.n):
            self.lib.orc_swarm_observe(self.h, e, _f(self.observations[e * self.A:(e + 1) * self.A]))

    def close(self):
        if self.h:
            self.lib.orc_swarm_close(self.h)
            self.h = None


def formation_target(task, idx, num_agents):
    out = np.zeros(3, np.float32)
    _orc_swarm().orc_swarm_formation_target(int(task), int(idx), int(num_agents), _f(out))
    return out


def puff_advantage(values, rewards, dones, importance, gamma, lam, rho_clip, c_clip, time_major=False):
    """The reference's CPU advantage (pufferlib.cpp:28-41,63-72) on float32 NumPy arrays.
    Returns (advantages, sum_t |adv| per row)."""
    L = _orc()
    L.orc_puff_advantage.restype = None
    L.orc_puff_advantage.argtypes = [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                                     C.c_float, C.c_float, C.c_float, C.c_float]
    a = [np.ascontiguousarray(x, np.float32) for x in (values, rewards, dones, importance)]
    if time_major:
        horizon, rows = a[0].shape
        rs, ts = 1, rows
    else:
        rows, horizon = a[0].shape
        rs, ts = horizon, 1
    adv = np.zeros_like(a[0])
    prio = np.zeros(rows, np.float32)
    L.orc_puff_advantage(_f(a[0]), _f(a[1]), _f(a[2]), _f(a[3]), _f(adv), _f(prio), rows, horizon, rs, ts,
                         gamma, lam, rho_clip, c_clip)
    return adv, prio


def have_ref_advantage():
    return os.path.exists(REF_ADVANTAGE_SO)


_REF_ADV = None


def ref_puff_advantage(values, rewards, dones, importance, gamma, lam, rho_clip, c_clip):
    """The UNMODIFIED reference implementation (pufferlib/extensions/pufferlib.cpp puff_advantage, compiled into
    oracle/_ref/libref_advantage.so by oracle/Makefile) on row-major [num_steps, horizon] float32 arrays."""
    global _REF_ADV
    if _REF_ADV is None:
        import torch  # noqa: F401  (the reference source registers a torch op: its libraries must be loaded first)
        L = C.CDLL(REF_ADVANTAGE_SO)
        L.ref_puff_advantage.restype = None
        L.ref_puff_advantage.argtypes = [_fp, _fp, _fp, _fp, _fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
        _REF_ADV = L
    a = [np.ascontiguousarray(x, np.float32).copy() for x in (values, rewards, dones, importance)]
    rows, horizon = a[0].shape
    adv = np.zeros_like(a[0])
    _REF_ADV.ref_puff_advantage(_f(a[0]), _f(a[1]), _f(a[2]), _f(a[3]), _f(adv), gamma, lam, rho_clip, c_clip, rows, horizon)
    return adv
