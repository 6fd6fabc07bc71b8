"""CPU baseline worker.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and
--impl reference legs).

Runs the reference's C env step on host cores the way PufferLib itself uses more than
one core: one process per core, each owning a disjoint slice of the envs
(pufferlib/vector.py:226-488).  Each worker is the reference's own wrapper loop
(pufferlib/ocean/drone_race/drone_race.py:37-62, drone_swarm/drone_swarm.py:36-62): `env_init`
per env on NumPy slices, `vectorize`, then per step `self.actions[:] = actions` and the
reference's real `binding.vec_step` (env_binding.h:508-524) -- the unmodified CPython extension
compiled into oracle/_ref/ref_drone_{race,swarm} (kind "reference").  Where oracle/_ref is absent
the bit-exact CPU restatement steps instead (kind "port").

`policy=True` adds what PuffeRL.evaluate does between recv and send (pufferl.py:229-296) on the
CPU: the Default policy's forward (Linear 29->128, exact GELU, mean / value heads), Normal sampling
and the clip -- NumPy on the worker's core -- so the on-device rollout has a CPU figure beside it.

    python -m oracle.cpu_worker --procs P --envs N --steps K --warmup W [--kind reference|port]

prints one JSON line {"env_steps_per_s", "procs", "envs", "steps", "wall_s", "kind"}.
Wall time is that of the slowest worker after a common start barrier.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def have_ref_binding():
    import glob
    return bool(glob.glob(os.path.join(HERE, "_ref", "ref_drone_race", "binding*.so"))) and \
        bool(glob.glob(os.path.join(HERE, "_ref", "ref_drone_swarm", "binding*.so")))


class RefBindingEnv:
    """The reference's Python wrapper flow over its real CPython `binding` module."""

    def __init__(self, n, drones=0, max_rings=10, max_moves=1000):
        A = max(drones, 1)
        rows = n * A
        self.observations = np.zeros((rows, 41 if drones else 29), np.float32)
        self.actions = np.zeros((rows, 4), np.float32)
        self.rewards = np.zeros(rows, np.float32)
        self.terminals = np.zeros(rows, bool)
        self.truncations = np.zeros(rows, bool)
        if drones:
            from oracle._ref.ref_drone_swarm import binding
            kw = dict(num_agents=drones, max_rings=max_rings)
        else:
            from oracle._ref.ref_drone_race import binding
            kw = dict(max_rings=max_rings, max_moves=max_moves)
        self.binding = binding
        handles = []
        for i in range(n):
            s = slice(i * A, (i + 1) * A)
            handles.append(binding.env_init(self.observations[s], self.actions[s], self.rewards[s], self.terminals[s],
                                            self.truncations[s], i, **kw))
        self.c_envs = binding.vectorize(*handles)

    def reset(self, seed):
        self.binding.vec_reset(self.c_envs, int(seed))

    def step(self, actions):
        self.actions[:] = actions
        self.binding.vec_step(self.c_envs)

    def close(self):
        self.binding.vec_close(self.c_envs)


class CpuPolicy:
    """pufferlib.models.Default for a Box action space + sample_logits, in NumPy (float32)."""

    def __init__(self, obs_dim, rng, hidden=128):
        self.w1 = (rng.standard_normal((obs_dim, hidden)) * (2.0 / obs_dim) ** 0.5).astype(np.float32)
        self.b1 = np.zeros(hidden, np.float32)
        self.w2 = (rng.standard_normal((hidden, 5)) * 0.01).astype(np.float32)  # 4 means + value
        self.b2 = np.zeros(5, np.float32)
        self.std = np.ones(4, np.float32)
        self.rng = rng
        from scipy.special import erf
        self.erf = erf

    def act(self, obs):
        h = obs @ self.w1 + self.b1
        h = 0.5 * h * (1.0 + self.erf(h * np.float32(0.70710678)))
        out = h @ self.w2 + self.b2
        noise = self.rng.standard_normal((obs.shape[0], 4), dtype=np.float32)
        action = out[:, :4] + self.std * noise
        logp = (-0.5 * noise * noise - 0.9189385).sum(axis=1)
        return np.clip(action, -1.0, 1.0), logp, out[:, 4]


def _worker(rank, n, steps, warmup, kind, seed, barrier, out, drones=0, policy=False):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import pyoracle as po
    if kind == "reference":
        env = RefBindingEnv(n, drones)
        step = env.step
    else:
        env = po.OrcSwarm(n, drones, 10) if drones else po.OrcRace(n)
        step = lambda a: env.step(a, mode=po.RESET_LIBC)  # noqa: E731
    rng = np.random.default_rng(1234 + rank)
    tape = rng.uniform(-1.0, 1.0, size=(16, n * max(drones, 1), 4)).astype(np.float32)
    pol = CpuPolicy(env.observations.shape[1], rng) if policy else None
    env.reset(seed)

    def one(t):
        if pol is None:
            step(tape[t % 16])
        else:
            a, _, _ = pol.act(env.observations)
            step(a)

    for t in range(warmup):
        one(t)
    barrier.wait()
    t0 = time.perf_counter()
    for t in range(steps):
        one(t)
    dt = time.perf_counter() - t0
    out.put((rank, dt, int(np.asarray(env.terminals).sum())))
    env.close()


def run(procs, envs, steps, warmup, kind="reference", seed=0, drones=0, policy=False):
    """Total `envs` split over `procs` processes; returns a result dict.  drones > 0 selects the
    swarm env (envs x drones agents); env_steps_per_s then counts drone-steps."""
    from oracle import pyoracle as po
    if kind == "reference" and not have_ref_binding():
        kind = "port"
    if kind == "port" and not os.path.exists(po.ORACLE_SO):
        po.build()
    per = max(1, envs // procs)
    os.environ.setdefault("OMP_NUM_THREADS", "1")  # inherited by the spawned workers: one BLAS thread per core
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(procs)
    out = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, per, steps, warmup, kind, seed, barrier, out, drones, policy)) for r in range(procs)]
    for p in ps:
        p.start()
    res = [out.get() for _ in ps]
    for p in ps:
        p.join()
    wall = max(r[1] for r in res)
    total = per * procs
    api = ("binding.vec_step of the unmodified reference extension (oracle/_ref)" if kind == "reference"
           else "CPU restatement (oracle/liboracle.so)")
    r = {"env_steps_per_s": total * max(drones, 1) * steps / wall, "procs": procs, "envs": total, "steps": steps,
         "wall_s": wall, "kind": kind, "api": api}
    if drones:
        r["drones"] = drones
    if policy:
        r["policy"] = "numpy Default policy forward + Normal sample on the same cores"
    return r


def host_cores():
    """Cores the baseline may use: physical cores (the cap PufferLib enforces,
    pufferlib/vector.py:246-253) limited by this process's affinity mask."""
    try:
        import psutil
        phys = psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:  # noqa: BLE001
        phys = os.cpu_count()
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        aff = phys
    return max(1, min(phys, aff))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--kind", default="reference")
    ap.add_argument("--drones", type=int, default=0)
    ap.add_argument("--policy", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.procs or host_cores(), a.envs, a.steps, a.warmup, a.kind, drones=a.drones, policy=a.policy)))
