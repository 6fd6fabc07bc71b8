"""CPU baseline worker.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and
--impl reference legs).

Runs the reference's C env step on host cores the way PufferLib itself uses more than
one core: one process per core, each owning a disjoint slice of the envs
(pufferlib/vector.py:226-488).  Each worker mirrors DroneRace.step
(pufferlib/ocean/drone_race/drone_race.py:58-62): copy the action batch into the env's
action buffer, then vec_step (env_binding.h:520-522: a C loop over c_step).

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


def _worker(rank, n, steps, warmup, kind, seed, barrier, out, drones=0):
    from oracle import pyoracle as po
    if drones:  # swarm: n envs of `drones` agents (pufferlib/ocean/drone_swarm), max_rings=10
        env = po.RefSwarm(n, drones, 10) if kind == "reference" else po.OrcSwarm(n, drones, 10)
    else:
        env = po.RefRace(n) if kind == "reference" else po.OrcRace(n)
    if kind == "reference":
        step = env.step
    else:
        step = lambda a: env.step(a, mode=po.RESET_LIBC)  # noqa: E731
    rng = np.random.default_rng(1234 + rank)
    tape = rng.uniform(-1.0, 1.0, size=(16, n * max(drones, 1), 4)).astype(np.float32)
    env.reset(seed)
    for t in range(warmup):
        step(tape[t % 16])
    barrier.wait()
    t0 = time.perf_counter()
    for t in range(steps):
        step(tape[t % 16])
    dt = time.perf_counter() - t0
    out.put((rank, dt, int(env.terminals.sum())))
    env.close()


def run(procs, envs, steps, warmup, kind="reference", seed=0, drones=0):
    """Total `envs` split over `procs` processes; returns a result dict.  drones > 0 selects the
    swarm env (envs x drones agents); env_steps_per_s then counts drone-steps."""
    from oracle import pyoracle as po
    if kind == "reference" and not po.have_ref():
        kind = "port"
    if kind == "port" and not os.path.exists(po.ORACLE_SO):
        po.build()
    per = max(1, envs // procs)
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(procs)
    out = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, per, steps, warmup, kind, seed, barrier, out, drones)) for r in range(procs)]
    for p in ps:
        p.start()
    res = [out.get() for _ in ps]
    for p in ps:
        p.join()
    wall = max(r[1] for r in res)
    total = per * procs
    if drones:
        return {"env_steps_per_s": total * drones * steps / wall, "procs": procs, "envs": total, "steps": steps,
                "wall_s": wall, "kind": kind, "drones": drones}
    return {"env_steps_per_s": total * steps / wall, "procs": procs, "envs": total, "steps": steps,
            "wall_s": wall, "kind": kind}


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
    a = ap.parse_args()
    print(json.dumps(run(a.procs or host_cores(), a.envs, a.steps, a.warmup, a.kind)))
