#!/usr/bin/env python
"""bench.py -- drone env-steps/sec on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): the single-drone ring-race
env, 1,048,576 envs per GPU (max_rings=10, max_moves=1000), vec_reset(seed=0),
actions from a tape of 16 pre-generated [N,4] f32 U(-1,1) tensors (seed 1234)
cycled t % 16, step-only.  A "step" is one vec_step over all envs of all ranks.
Envs shard across ranks with no per-step communication (weak scaling: 1M envs per
GPU); the only collective is the episode-statistics all-reduce of vec_log.

value      = env-steps/s, whole job, device-resident buffers (zero-copy path),
             timed with CUDA events on the launching stream, max over ranks.
e2e        = the same metric through the public host-buffer API
             (DroneRace(buffers="host").step(numpy_actions)): H2D of the actions and
             D2H of observations/rewards/terminals inside the timed region.
roofline   = algorithmic bytes (373 B/env-step, SURVEY 8d / DESIGN.md) / kernel time
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline = the reference's own C step (oracle/_ref) on this box's host cores,
             one process per physical core, bounded sample (rank 0, N=1 only).
--impl reference prints the CPU reference arm in the same format.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 1 << 20
MAX_RINGS, MAX_MOVES = 10, 1000
TAPE_LEN, TAPE_SEED = 16, 1234
ALGO_BYTES_PER_ENV_STEP = 373  # reads 172 + writes 201, SURVEY.md 8(d)
METRIC = "drone_env_steps_per_sec"
UNIT = "env-steps/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the step kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("race_step_kernel_fast_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:  # noqa: BLE001
            return local_rank
    return local_rank


def bind_to_gpu_numa_node(gpu_index):
    """Multi-rank runs: keep this rank's host threads -- and therefore the first-touch placement of its
    pinned NumPy buffers -- on the CPUs NVML names as local to its GPU (what `numactl` would do for a
    launcher).  Only matters for the e2e leg (PCIe traffic from host memory).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = ideal & allowed
        if pick and pick != allowed:
            os.sched_setaffinity(0, pick)
            return f"nvml-local cpus ({len(pick)} of {len(allowed)})"
        return "unchanged (all allowed cpus are local)" if pick else "unchanged (no local cpu allowed)"
    except Exception as e:  # noqa: BLE001
        return f"unchanged ({type(e).__name__})"


def cpu_sample(steps_budget_s=20.0, kind="reference", drones=0):
    """Bounded CPU sample of the same workload shape (rank 0, N=1 only)."""
    from oracle import cpu_worker
    procs = cpu_worker.host_cores()
    per = 2048 if not drones else max(1, 2048 // drones)
    probe = cpu_worker.run(procs, procs * per, 20, 5, kind, drones=drones)
    rate = probe["env_steps_per_s"]
    envs = procs * per * 2
    units = envs * max(drones, 1)
    steps = int(max(50, min(2000, rate * steps_budget_s / units)))
    res = cpu_worker.run(procs, envs, steps, 20, kind, drones=drones)
    return res


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU step on the host cores, same metric/config."""
    if rank != 0:
        return
    from oracle import cpu_worker
    procs = cpu_worker.host_cores()
    kind = "reference"
    # bounded sample: each "step" steps `envs` envs of the 1M-env workload so that
    # warmup + steps finish within a few minutes whatever K the driver asks for
    probe = cpu_worker.run(procs, procs * 2048, 20, 5, kind)
    rate = probe["env_steps_per_s"]
    total_steps = max(1, args.steps + args.warmup)
    envs = int(rate * 120.0 / total_steps)
    envs = max(procs * 256, min(ENVS_PER_GPU * world, envs // (procs * 256) * (procs * 256)))
    res = cpu_worker.run(procs, envs, args.steps, args.warmup, kind)
    v = res["env_steps_per_s"]
    sample = f"{res['envs']} of {ENVS_PER_GPU * world} envs x {args.steps} steps, {procs} processes (one per physical core)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["wall_s"] / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, "host cores only (reference C c_step via oracle/_ref)"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": res["kind"], "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world, parallelism):
    return {"workload": "drone_race single-drone gate-race, 1,048,576 envs per GPU, max_rings=10, max_moves=1000, "
                        "fixed random action tape (16 x U(-1,1), seed 1234), step-only incl. auto-resets "
                        "(BASELINE.json configs[1])",
            "envs_per_gpu": ENVS_PER_GPU, "total_envs": ENVS_PER_GPU * world, "parallelism": parallelism,
            "l2": "working set 391 MB per step per GPU > 126 MB L2 (inputs larger than L2, no flush needed)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--launch", default="tape", choices=["tape", "single"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = "unchanged (single rank)"
    if world > 1:
        affinity = bind_to_gpu_numa_node(physical_gpu_index(local_rank))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from drone_b200.vec import RaceVec
    from drone_b200.drone_race import DroneRace

    n = args.envs_per_gpu
    vec = RaceVec(n, max_rings=MAX_RINGS, max_moves=MAX_MOVES, seed=0, device=dev, math=args.math,
                  env_id_base=rank * n)
    g = torch.Generator(device="cpu").manual_seed(TAPE_SEED + rank)
    tape = (torch.rand((TAPE_LEN, n, 4), generator=g) * 2.0 - 1.0).to(dev)
    vec.reset(0)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The timed loop is the reference's own cached-action perf loop (drone_race.py:83-90): one
    # vec_step per action batch of the tape, issued through b2d_vec_step_tape (one kernel launch
    # per step; `--launch single` issues them one b2d_vec_step_from call at a time instead).
    def run_steps(t0, k):
        if args.launch == "tape":
            vec.step_tape(tape, t0 % TAPE_LEN, k)
        else:
            for j in range(k):
                vec.step(tape[(t0 + j) % TAPE_LEN])

    t = 0
    run_steps(t, args.warmup)
    t += args.warmup
    barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = vec.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    term_count = torch.zeros((), dtype=torch.int64, device=dev)
    ev0.record(stream)
    run_steps(t, args.steps)
    t += args.steps
    ev1.record(stream)
    barrier()
    sampler.stop()
    launches = vec.kernel_launches - launches0
    ms = ev0.elapsed_time(ev1)
    # reset fraction of the workload, measured after the timed region (one extra step)
    vec.step(tape[t % TAPE_LEN])
    term_count = vec.terminals.sum()
    reset_frac = float(term_count.item()) / n
    stats = vec.log(group=True if world > 1 else None)

    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    total_envs = n * world
    value = total_envs * args.steps / (ms_max * 1e-3)

    # ---- e2e: public host-buffer API (numpy in / numpy out), PCIe inside the timed region
    e2e = None
    if not args.no_e2e:
        vec.close()
        del tape
        torch.cuda.empty_cache()
        env = DroneRace(num_envs=n, report_interval=1 << 30, seed=0, buffers="host", device=local_rank,
                        math=args.math, env_id_base=rank * n)
        env.reset(0)
        rng = np.random.default_rng(TAPE_SEED + rank)
        htape = rng.uniform(-1, 1, size=(4, n, 4)).astype(np.float32)
        e2e_steps = max(3, min(args.steps, 200))
        for k in range(3):
            env.step(htape[k % 4])
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            obs, rew, term, trunc, info = env.step(htape[k % 4])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": total_envs * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": total_envs * 16,
               "d2h_bytes_per_step": total_envs * (116 + 4 + 1), "steps": e2e_steps,
               "ms_per_step": dt / e2e_steps * 1e3,
               "api": "drone_b200.drone_race.DroneRace(buffers='host').step(np.ndarray) -> binding.vec_step_actions -> b2d_vec_step_host_from",
               "host_cpu_affinity": affinity,
               "checksum": float(np.abs(obs).sum(dtype=np.float64))}
        env.close()

    if rank == 0:
        peak, peak_src = measured_peaks()
        per_gpu_gbs = ALGO_BYTES_PER_ENV_STEP * n * args.steps / (ms_max * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, f"env-sharded x{world}, no per-step collective; vec_log all-reduce over NCCL"),
            "math": args.math, "launch": args.launch, "reset_fraction_per_step": reset_frac,
            "episode_stats": stats,
            "clocks": sampler.summary(),
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s",
                         "frac": per_gpu_gbs / peak, "traffic": ncu_traffic(),
                         "kernel": f"race_step_kernel<{'strict' if args.math == 'strict' else 'fast'}>",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * n,
                         "avg_launch_us": ms_max / args.steps * 1e3, "peak_source": peak_src, "per": "GPU"},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            try:
                res = cpu_sample()
                line["cpu_baseline"] = {
                    "value": res["env_steps_per_s"], "unit": UNIT, "cores": res["procs"], "kind": res["kind"],
                    "sample": f"{res['envs']} envs x {res['steps']} steps of the same workload, "
                              f"{res['procs']} processes (one per physical core), {res['wall_s']:.1f} s"}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
