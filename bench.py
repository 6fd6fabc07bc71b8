#!/usr/bin/env python
"""bench.py -- drone env-steps/sec on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload race|race4096|swarm16|swarm32|swarm64|rollout] [--math fast|strict]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1], SURVEY 8d "C2"): the single-drone ring-race env,
1,048,576 envs per GPU (max_rings=10, max_moves=1000), vec_reset(seed=0), actions from a tape of 16
pre-generated [N,4] f32 U(-1,1) tensors (seed 1234) cycled t % 16, step-only incl. auto-resets.  A
"step" is one vec_step over all envs of all ranks.  Envs shard across ranks with no per-step
communication (weak scaling: 1M envs per GPU); the only collective is the episode-statistics
all-reduce of vec_log.

value      = env-steps/s, whole job, device-resident buffers (zero-copy path), timed with CUDA
             events on the launching stream, max over ranks.
e2e        = the same metric through the public host-buffer API
             (DroneRace(buffers="host").step(numpy_actions)): H2D of the actions and D2H of
             observations/rewards/terminals inside the timed region; `pcie_floor_ms` is the same
             byte counts as bare pinned-memory copies, measured in the same run.
roofline   = algorithmic bytes (373 B/env-step, SURVEY 8d / DESIGN.md) / kernel time against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json (strict math: FP32 issue slots).
cpu_baseline = the reference's own C step (binding.vec_step of oracle/_ref) on this box's host
             cores, one process per physical core, bounded sample (rank 0, N=1 only).
parity     = what tests/test_fast_parity_gpu.py measured for the kernels timed here (committed
             copy: profiles/parity_r02.json).
extra      = (N=1, default run) the other BASELINE.json configs, each a full sub-result with its own
             value / ms_per_step / roofline / cpu_baseline / clocks: configs[0] race 4096 envs,
             configs[2] swarm 65,536 envs x 16 and x 64 drones, configs[3] on-device rollout K=128,
             plus the headline kernel in strict math and as stand-alone launches.
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
ALGO_BYTES_PER_ENV_STEP = 373    # reads 172 + writes 201, SURVEY.md 8(d)
ALGO_BYTES_PER_DRONE_STEP = 521  # swarm, SURVEY.md 8(d)
SWARM_OBS_DIM = 41
ALGO_FLOPS_PER_ENV_STEP = 1330   # as-written FP32 operations of the reference step, SURVEY.md 8(d)
# rollout (fused policy + env kernel): per env-step only the experience row leaves the SM:
# obs 116 + action 16 + logprob, value, reward, terminal 4 x 4 = 148 B (DESIGN.md "rollout")
ALGO_BYTES_PER_ROLLOUT_STEP = 148
INTERNAL_WARMUP = 64             # steps before the caller's warm-up: past the post-vec_reset transient
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


def ncu_traffic(key):
    """DRAM bytes per launch from the committed ncu --set full capture (not measured in this run)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:  # noqa: BLE001
            pass
    return None


def parity_summary():
    """The committed parity record of the fast kernels (tests/test_fast_parity_gpu.py on a B200)."""
    p = os.path.join(ROOT, "profiles", "parity_r02.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
    except Exception:  # noqa: BLE001
        return None
    out = {"source": "profiles/parity_r02.json (written by tests/test_fast_parity_gpu.py on a B200; not re-measured in this run)"}
    r = d.get("race_fast_per_step")
    if r:
        out["race_fast_vs_oracle"] = {
            "env_steps_compared": r.get("env_steps"), "shape": "slices of the 1,048,576-env vector, per-step resync",
            "integer_flips": sum(int(r.get(k, 0)) for k in ("terminal_flips", "reward_flips", "tick_flips", "ring_idx_flips", "return_flips")),
            "worst_obs_abs_err": r.get("worst_obs_abs"), "worst_pos_abs_err_m": r.get("worst_pos_abs"),
            "worst_err_in_tolerance_units": max(r.get("worst_obs_tol", 0.0), r.get("worst_state_tol", 0.0)),
            "tolerance": "1e-5 rel + 1e-6 abs per step", "guard_replay_rate": r.get("guard_replay_rate")}
    for rec in ("bench_uniform", "near_hover"):
        k = d.get(f"race_fast_drift_{rec}")
        if k:
            out[f"race_fast_drift_1000_steps_{rec}"] = {
                "max_abs_dpos_m": k.get("max_abs_dpos_m"), "mean_abs_dpos_m": k.get("mean_abs_dpos_m"),
                "max_abs_dquat": k.get("max_abs_dquat"), "mean_abs_dquat": k.get("mean_abs_dquat"),
                "identical_event_history": k.get("envs_with_identical_event_history_after_1000_steps")}
    for A in (16, 64):
        k = d.get(f"swarm_fast_per_step_A{A}")
        if k:
            out[f"swarm{A}_fast_vs_oracle"] = {
                "drone_steps_compared": k.get("drone_steps"),
                "integer_flips": sum(int(k.get(x, 0)) for x in ("terminal_flips", "ring_idx_flips", "episode_length_flips", "collision_count_flips")),
                "nearest_neighbour_flips": k.get("nearest_neighbour_flips"),
                "worst_err_in_tolerance_units": max(k.get("worst_obs_tol", 0.0), k.get("worst_state_tol", 0.0))}
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self._go = threading.Event()
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

    def arm(self):
        """Begin sampling.  Called by the timing code right AFTER it has queued the timed launches: the GPU is then
        busy with them for the rest of the region, and the NVML calls (which contend with the CUDA launch path for
        driver locks) do not sit in front of the launches -- that cost 2-3 us per step in a 20-step region."""
        self._go.set()

    def run(self):
        if not self.ok:
            return
        self._go.wait()
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
        self._go.set()
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


# ------------------------------------------------------------------------------------- CPU arm
def cpu_sample(budget_s=15.0, kind="reference", drones=0, envs=None, policy=False):
    """Bounded CPU sample of the same workload shape (rank 0, N=1 only): the reference's own
    binding.vec_step, one process per physical core."""
    from oracle import cpu_worker
    procs = cpu_worker.host_cores()
    per = 2048 if not drones else max(1, 2048 // drones)
    if envs is not None:  # a fixed-size workload (configs[0]: 4096 envs in all)
        per = max(1, envs // procs)
    probe = cpu_worker.run(procs, procs * per, 20, 5, kind, drones=drones, policy=policy)
    rate = probe["env_steps_per_s"]
    n = procs * per * (1 if envs is not None else 2)
    units = n * max(drones, 1)
    steps = int(max(50, min(4000, rate * budget_s / units)))
    return cpu_worker.run(procs, n, steps, 20, kind, drones=drones, policy=policy)


def cpu_baseline_dict(res, unit, what):
    return {"value": res["env_steps_per_s"], "unit": unit, "cores": res["procs"], "kind": res["kind"],
            "sample": f"{res['envs']} envs{what} x {res['steps']} steps of the same workload, {res['procs']} processes "
                      f"(one per physical core), {res['wall_s']:.1f} s; {res.get('api', '')}"}


def safe_cpu_baseline(unit, what="", **kw):
    try:
        return cpu_baseline_dict(cpu_sample(**kw), unit, what)
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": unit, "cores": 0, "kind": "unavailable", "sample": repr(e)}


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
    sample = (f"{res['envs']} of {ENVS_PER_GPU * world} envs x {args.steps} steps, {procs} processes (one per physical core); "
              f"{res.get('api', '')}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["wall_s"] / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": race_config(ENVS_PER_GPU, world, f"env-sharded x{world}, no per-step collective; vec_log all-reduce over NCCL"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": res["kind"], "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def race_config(n, world, parallelism, which="configs[1]"):
    ws = ALGO_BYTES_PER_ENV_STEP * n / 1e6
    return {"workload": f"drone_race single-drone gate-race, {n:,} envs per GPU, max_rings=10, max_moves=1000, "
                        f"fixed random action tape (16 x U(-1,1), seed 1234), step-only incl. auto-resets "
                        f"(BASELINE.json {which})",
            "envs_per_gpu": n, "total_envs": n * world, "parallelism": parallelism,
            "l2": (f"working set {ws:.0f} MB per step per GPU > 126 MB L2 (inputs larger than L2, no flush needed)" if ws > 126 else
                   f"working set {ws:.1f} MB per step: L2-resident (the reference's own CPU-sized case; timed as is and said so)")}


# ------------------------------------------------------------------------------------- GPU workloads
def roofline_hbm(bytes_per_launch, us, kernel, traffic=None, traffic_src=None):
    peak, peak_src = measured_peaks()
    gbs = bytes_per_launch / (us * 1e-6) / 1e9
    return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": traffic,
            "traffic_source": traffic_src, "kernel": kernel, "algorithmic_bytes_per_launch": bytes_per_launch,
            "avg_launch_us": us, "peak_source": peak_src, "per": "GPU"}


def roofline_fp32(flops_per_launch, us, kernel, sm_mhz, bytes_per_launch):
    """Strict math is bound by FP32 issue slots, not by HBM: peak = SMs x 128 lanes x clock (one non-FMA op
    per lane per cycle; the strict kernel may not contract into FMAs)."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mhz = sm_mhz or 1965
    peak = sms * 128 * mhz * 1e6 / 1e12
    ach = flops_per_launch / (us * 1e-6) / 1e12
    hbm_peak, _ = measured_peaks()
    return {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "kernel": kernel, "algorithmic_flops_per_launch": flops_per_launch, "avg_launch_us": us,
            "peak_source": f"{sms} SMs x 128 FP32 lanes x {mhz} MHz (as-written reference operations, no FMA; IEEE division and "
                           "square root expand to ~10 issue slots each, so 1.0 is not reachable)",
            "hbm_frac_for_comparison": bytes_per_launch / (us * 1e-6) / 1e9 / hbm_peak, "per": "GPU"}


def time_race(n, math, launch, steps, warmup, dev, rank=0, world=1, dist=None):
    """CUDA-event time of `steps` vec_steps of the race env (after INTERNAL_WARMUP + warmup untimed ones)."""
    import torch
    from drone_b200.vec import RaceVec
    vec = RaceVec(n, max_rings=MAX_RINGS, max_moves=MAX_MOVES, seed=0, device=dev, math=math, env_id_base=rank * n)
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
        if launch == "tape":
            vec.step_tape(tape, t0 % TAPE_LEN, k)
        else:
            for j in range(k):
                vec.step(tape[(t0 + j) % TAPE_LEN])

    # (NVML is initialised BEFORE the warm-up: its first initialisation takes ~0.1 s, and a GPU left idle that long between
    # the warm-up and a 20-step timed region starts the region below its boost clock)
    sampler = ClockSampler(physical_gpu_index(dev.index))
    t = 0
    run_steps(t, INTERNAL_WARMUP + warmup)  # episodes last ~40 steps: the reset rate is steady after 64
    t += INTERNAL_WARMUP + warmup
    barrier()
    sampler.start()
    launches0 = vec.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    run_steps(t, steps)
    t += steps
    ev1.record(stream)
    sampler.arm()
    barrier()
    sampler.stop()
    launches = vec.kernel_launches - launches0
    ms = ev0.elapsed_time(ev1)
    vec.step(tape[t % TAPE_LEN])  # reset fraction of the workload, measured after the timed region
    reset_frac = float(vec.terminals.sum().item()) / n
    replays = vec.guard_replays if math == "fast" else 0
    stats = vec.log(group=True if world > 1 else None)
    total_steps = INTERNAL_WARMUP + warmup + steps + 1
    vec.close()
    del tape
    torch.cuda.empty_cache()
    return {"ms": ms, "launches": int(launches), "reset_frac": reset_frac, "stats": stats, "clocks": sampler.summary(),
            "guard_replay_rate": replays / float(n * total_steps)}


def race_line(res, n, math, launch, steps, warmup, world, ms_max, which="configs[1]"):
    us = ms_max / steps * 1e3
    kernel = f"race_step_kernel<{'strict' if math == 'strict' else 'fast'}>"
    if math == "strict":
        roof = roofline_fp32(ALGO_FLOPS_PER_ENV_STEP * n, us, kernel, res["clocks"].get("sm_mhz"), ALGO_BYTES_PER_ENV_STEP * n)
    else:
        roof = roofline_hbm(ALGO_BYTES_PER_ENV_STEP * n, us, kernel,
                            ncu_traffic("race_step_kernel_fast_bytes_per_launch") if n == ENVS_PER_GPU else None,
                            "profiles/roofline_traffic.json (ncu --set full capture of this kernel at 1,048,576 envs; not measured in this run)"
                            if n == ENVS_PER_GPU else None)
    return {
        "metric": METRIC, "value": n * world * steps / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "internal_warmup": INTERNAL_WARMUP, "ms_per_step": ms_max / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": race_config(n, world, f"env-sharded x{world}, no per-step collective; vec_log all-reduce over NCCL", which),
        "math": math, "launch": launch, "reset_fraction_per_step": res["reset_frac"],
        "guard_replay_rate": res["guard_replay_rate"], "episode_stats": res["stats"], "clocks": res["clocks"],
        "gpu_launches": res["launches"], "roofline": roof}


def bench_swarm(drones, steps, warmup, dev, math="fast", envs=1 << 16, cpu=True, e2e=True):
    import torch
    from drone_b200.vec import SwarmVec
    rows = envs * drones
    vec = SwarmVec(envs, drones, 10, seed=0, device=dev, math=math)
    g = torch.Generator(device="cpu").manual_seed(TAPE_SEED)
    tape = (torch.rand((4, rows, 4), generator=g) * 2.0 - 1.0).to(dev)
    vec.reset(0)
    sampler = ClockSampler(physical_gpu_index(dev.index))
    for t in range(max(warmup, 20)):
        vec.step(tape[t % 4])
    torch.cuda.synchronize()
    sampler.start()
    l0 = vec.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for t in range(steps):
        vec.step(tape[t % 4])
    ev1.record()
    sampler.arm()
    torch.cuda.synchronize()
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = vec.kernel_launches - l0
    stats = vec.log()
    replays = vec.guard_replays if math == "fast" else 0
    value = rows * steps / (ms * 1e-3)
    us = ms / steps * 1e3
    traffic = ncu_traffic(f"swarm_kernel_fast_A{drones}_bytes_per_launch") if (envs == 1 << 16 and math == "fast") else None
    line = {"metric": "drone_steps_per_sec", "value": value, "unit": "drone-steps/s", "n_gpus": 1, "steps": steps,
            "warmup": max(warmup, 20), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "math": math,
            "config": {"workload": f"drone_swarm {envs:,} envs x {drones} drones, max_rings=10, fixed random action tape "
                                   f"(4 x U(-1,1)), step-only incl. respawns and the 1023-tick env-wide reset "
                                   f"(BASELINE.json configs[2])",
                       "rows": rows, "l2": f"working set {rows * ALGO_BYTES_PER_DRONE_STEP / 1e6:.0f} MB per step > 126 MB L2"},
            "episode_stats": stats, "clocks": sampler.summary(), "gpu_launches": int(launches),
            "guard_replay_env_rate": replays / float(envs * (steps + max(warmup, 20))),
            "roofline": roofline_hbm(ALGO_BYTES_PER_DRONE_STEP * rows, us, f"swarm_kernel<{math}>", traffic,
                                     "profiles/roofline_traffic.json (not measured in this run)" if traffic else None)}
    vec.close()
    del tape
    torch.cuda.empty_cache()
    if e2e:  # the reference's NumPy-buffer contract: host actions in, host observations / rewards / terminals out, every step
        import numpy as np
        from drone_b200.drone_swarm import DroneSwarm
        env = DroneSwarm(num_envs=envs, num_drones=drones, max_rings=10, seed=0, report_interval=1 << 30, buffers="host",
                         device=dev.index, math=math)
        env.reset(0)
        rng = np.random.default_rng(TAPE_SEED)
        htape = [env.pinned_actions() for _ in range(2)]  # inputs from pinned host memory (the bench contract)
        for h in htape:
            h[:] = rng.uniform(-1, 1, size=(rows, 4)).astype(np.float32)
        e2e_steps = max(3, min(steps, 4_000_000 * 12 // rows))
        for k in range(3):
            env.step(htape[k % 2])
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            obs, rew, term, trunc, info = env.step(htape[k % 2])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        checksum = float(np.abs(obs[:1 << 16]).sum(dtype=np.float64))
        env.close()
        del htape
        floor = pcie_floor_ms(rows, dev, down_bytes=SWARM_OBS_DIM * 4 + 5)
        line["e2e"] = {"value": rows * e2e_steps / dt, "unit": "drone-steps/s", "h2d_bytes_per_step": rows * 16,
                       "d2h_bytes_per_step": rows * (SWARM_OBS_DIM * 4 + 5), "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
                       "pcie_floor_ms": floor, "ms_per_step_over_pcie_floor": dt / e2e_steps * 1e3 / floor,
                       "api": "drone_b200.drone_swarm.DroneSwarm(buffers='host').step(np.ndarray) -> binding.vec_step_actions -> "
                              "b2d_vec_step_host_from (chunked H2D / kernel / D2H pipeline)", "checksum": checksum}
    if cpu:
        line["cpu_baseline"] = safe_cpu_baseline("drone-steps/s", f" x {drones} drones", budget_s=8.0, drones=drones)
    return line


def bench_rollout(replays, dev, envs=1 << 20, horizon=128, impl="auto", cpu=True):
    """BASELINE.json configs[3]: policy forward + sampling + experience stores + env step, K=128 per launch/graph."""
    import torch
    from drone_b200.rollout import DeviceRollout, DronePolicy
    from drone_b200.vec import RaceVec
    torch.manual_seed(0)
    vec = RaceVec(envs, seed=0, device=dev)
    vec.reset(0)
    policy = DronePolicy().to(dev)
    ro = DeviceRollout(vec, policy, horizon=horizon, policy_impl=impl)
    sampler = ClockSampler(physical_gpu_index(dev.index))
    ro.collect()
    torch.cuda.synchronize()
    sampler.start()
    l0 = ro.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(replays):
        ro.collect()
    ev1.record()
    sampler.arm()
    torch.cuda.synchronize()
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    steps = replays * horizon
    us = ms / steps * 1e3
    launches = ro.kernel_launches - l0
    fused = ro.policy_impl == "rollout_kernel"
    algo = ALGO_BYTES_PER_ROLLOUT_STEP if fused else (ALGO_BYTES_PER_ENV_STEP + 285)
    line = {"metric": "rollout_env_steps_per_sec", "value": envs * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
            "steps": steps, "warmup": horizon, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 env / tf32 policy GEMMs (the reference's matmul precision, pufferl.py:55)",
            "data": "synthetic", "policy_impl": ro.policy_impl, "gpu_launches": int(launches),
            "config": {"workload": f"on-device rollout: Default policy (29 -> 128 GELU -> 4 means + value, Normal sampling) + race env step, "
                                   f"{envs:,} envs, K={horizon} steps per collect() (BASELINE.json configs[3])",
                       "l2": f"experience written per step {envs * 148 / 1e6:.0f} MB > 126 MB L2"},
            "episode_stats": vec.log(), "clocks": sampler.summary(),
            "roofline": roofline_hbm(algo * envs, us, ro.kernel_name,
                                     ncu_traffic("race_rollout_kernel_bytes_per_step") if (fused and envs == ENVS_PER_GPU) else None,
                                     "profiles/roofline_traffic.json (per step of a K = 128 launch; not measured in this run)"
                                     if (fused and envs == ENVS_PER_GPU) else None)}
    line["roofline"]["note"] = ("bytes that must move per env-step: " +
                                ("the experience row only (state stays on chip for the K steps)" if fused else
                                 "env step 373 B + policy step 285 B (two kernels per step)") +
                                "; the kernel is FP32-issue bound (GELU x 128 + the env step per env-step: 2,180 warp-instructions per "
                                "warp-step, ncu issue-active 57 % at 3 CTAs per SM = the TMEM limit), see DESIGN.md")
    vec.close()
    del ro
    torch.cuda.empty_cache()
    if cpu:
        line["cpu_baseline"] = safe_cpu_baseline(UNIT, budget_s=8.0, policy=True)
    return line


def pcie_floor_ms(n, dev, iters=5, down_bytes=121):
    """The bytes of one host-buffer step (16 B/env up, 121 B/env down) as bare pinned-memory copies on two
    streams, nothing else: the floor the e2e figure sits on."""
    import torch
    up_h = torch.zeros(n * 16, dtype=torch.uint8, pin_memory=True)
    dn_h = torch.zeros(n * down_bytes, dtype=torch.uint8, pin_memory=True)
    up_d = torch.zeros(n * 16, dtype=torch.uint8, device=dev)
    dn_d = torch.zeros(n * down_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = None
    for _ in range(iters + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            up_d.copy_(up_h, non_blocking=True)
        with torch.cuda.stream(s2):
            dn_h.copy_(dn_d, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="race", choices=["race", "race4096", "swarm16", "swarm32", "swarm64", "rollout"])
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--launch", default="tape", choices=["tape", "single"])
    ap.add_argument("--rollout-impl", default="auto", choices=["auto", "rollout_kernel", "fused", "torch"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the sub-results for the other BASELINE configs")
    ap.add_argument("--e2e-clamp", type=int, default=-1, help=argparse.SUPPRESS)  # A/B aid: write_clamped_actions of the e2e env
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
    cpu = not args.no_cpu_baseline

    # ---- secondary workloads as the main line (single GPU)
    if args.workload != "race":
        if world > 1:
            raise SystemExit("--workload other than race runs on one GPU")
        if args.workload == "race4096":
            res = time_race(4096, args.math, args.launch, args.steps, args.warmup, dev)
            line = race_line(res, 4096, args.math, args.launch, args.steps, args.warmup, 1, res["ms"], "configs[0]")
            if cpu:
                line["cpu_baseline"] = safe_cpu_baseline(UNIT, budget_s=8.0, envs=4096)
        elif args.workload.startswith("swarm"):
            line = bench_swarm(int(args.workload[5:]), min(args.steps, 1100) if args.steps != 2000 else 1100, args.warmup, dev,
                               math=args.math, cpu=cpu, e2e=not args.no_e2e)
        else:
            line = bench_rollout(max(1, min(8, args.steps // 128)) if args.steps != 2000 else 4, dev, impl=args.rollout_impl, cpu=cpu)
        print(json.dumps(line), flush=True)
        return

    affinity = "unchanged (single rank)"
    if world > 1:
        affinity = bind_to_gpu_numa_node(physical_gpu_index(local_rank))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from drone_b200.drone_race import DroneRace

    n = args.envs_per_gpu
    res = time_race(n, args.math, args.launch, args.steps, args.warmup, dev, rank, world, dist)
    tms = torch.tensor([res["ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    total_envs = n * world

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- e2e: public host-buffer API (numpy in / numpy out), PCIe inside the timed region
    e2e = None
    if not args.no_e2e:
        env = DroneRace(num_envs=n, report_interval=1 << 30, seed=0, buffers="host", device=local_rank,
                        math=args.math, env_id_base=rank * n, write_clamped_actions=args.e2e_clamp)
        env.reset(0)
        rng = np.random.default_rng(TAPE_SEED + rank)
        # the step's inputs come from pinned host memory (the bench contract): four page-locked action arrays
        htape = [env.pinned_actions() for _ in range(4)]
        for h in htape:
            h[:] = rng.uniform(-1, 1, size=(n, 4)).astype(np.float32)
        e2e_steps = max(3, min(args.steps, 200))
        for k in range(max(args.warmup, 3)):
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
        checksum = float(np.abs(obs).sum(dtype=np.float64))
        env.close()
        barrier()
        floor = pcie_floor_ms(n, dev)  # all ranks copy at once, like the step does
        tf = torch.tensor([floor], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        floor = float(tf.item())
        e2e = {"value": total_envs * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": total_envs * 16,
               "d2h_bytes_per_step": total_envs * (116 + 4 + 1), "steps": e2e_steps,
               "ms_per_step": dt / e2e_steps * 1e3,
               "pcie_floor_ms": floor, "ms_per_step_over_pcie_floor": dt / e2e_steps * 1e3 / floor,
               "pcie_floor_note": f"bare pinned copies of the same bytes per rank ({n * 16 / 1e6:.1f} MB up beside {n * 121 / 1e6:.1f} MB down), "
                                  f"all {world} rank(s) copying at once, best of 5, measured in this run"
                                  + ("; with several ranks the box's host DMA is the limiter, not the GPUs" if world > 1 else ""),
               "api": "drone_b200.drone_race.DroneRace(buffers='host').step(np.ndarray) -> binding.vec_step_actions -> b2d_vec_step_host_from",
               "inputs": "four page-locked NumPy action arrays (env.pinned_actions()), cycled; H2D from where they are, the env's own clamped "
                         "action buffer filled meanwhile; a pageable array takes the copy path (about +0.1 ms per step)",
               "host_cpu_affinity": affinity, "checksum": checksum}

    if rank == 0:
        line = race_line(res, n, args.math, args.launch, args.steps, args.warmup, world, ms_max)
        if e2e is not None:
            line["e2e"] = e2e
        par = parity_summary()
        if par:
            line["parity"] = par
        if world == 1 and cpu:
            line["cpu_baseline"] = safe_cpu_baseline(UNIT, budget_s=15.0)
        if world == 1 and not args.no_extra:
            extra = {}

            def sub(name, fn):
                try:
                    extra[name] = fn()
                except Exception as e:  # noqa: BLE001 - a failing sub-result must not cost the headline line
                    extra[name] = {"error": repr(e)}

            def race_variant(nn, math, launch, which, steps):
                r = time_race(nn, math, launch, steps, 20, dev)
                ln = race_line(r, nn, math, launch, steps, 20, 1, r["ms"], which)
                return ln

            def c0():
                ln = race_variant(4096, "fast", "tape", "configs[0]", 2000)
                if cpu:
                    ln["cpu_baseline"] = safe_cpu_baseline(UNIT, budget_s=6.0, envs=4096)
                return ln

            sub("race4096_configs0", c0)
            sub("swarm16_configs2", lambda: bench_swarm(16, 1100, 20, dev, cpu=cpu))
            sub("swarm64_configs2", lambda: bench_swarm(64, 1100, 20, dev, cpu=cpu))
            sub("rollout_configs3", lambda: bench_rollout(4, dev, impl=args.rollout_impl, cpu=cpu))
            sub("race_strict_math", lambda: race_variant(n, "strict", "tape", "configs[1]", 500))
            sub("race_single_launches", lambda: race_variant(n, "fast", "single", "configs[1]", 1000))
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
