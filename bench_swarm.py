#!/usr/bin/env python
"""bench_swarm.py -- secondary measurement: drone-steps/sec of the multi-drone swarm env
(BASELINE.json configs[2]: 65,536 envs x A drones, max_rings=10, warp/CTA-per-env kernel).
Same timing rules and JSON keys as bench.py (which stays the headline: configs[1]).

    python bench_swarm.py [--drones 16|32|64] [--envs 65536] [--steps K] [--warmup W] [--math fast|strict]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (ClockSampler, measured_peaks, cpu_sample)

ALGO_BYTES_PER_DRONE_STEP = 521


def swarm_traffic(args):
    """DRAM bytes per launch from the committed ncu --set full capture (only for the captured shape)."""
    if args.drones != 64 or args.envs != 65536 or args.math != "fast":
        return None
    try:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "roofline_traffic.json")
        return json.load(open(p)).get("swarm_kernel_fast_A64_bytes_per_launch")
    except Exception:  # noqa: BLE001
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--drones", type=int, default=64)
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=1100)  # crosses the 1023-tick env-wide reset
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import torch
    from drone_b200.vec import SwarmVec
    if not torch.cuda.is_available():
        raise SystemExit("bench_swarm.py needs a CUDA device (the product path has no CPU fallback)")
    dev = torch.device("cuda", 0)
    rows = args.envs * args.drones
    vec = SwarmVec(args.envs, args.drones, 10, seed=0, device=dev, math=args.math)
    g = torch.Generator(device="cpu").manual_seed(1234)
    tape = (torch.rand((4, rows, 4), generator=g) * 2.0 - 1.0).to(dev)
    vec.reset(0)
    for t in range(args.warmup):
        vec.step(tape[t % 4])
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(bench.physical_gpu_index(0))
    sampler.start()
    l0 = vec.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for t in range(args.steps):
        vec.step(tape[t % 4])
    ev1.record()
    torch.cuda.synchronize()
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = vec.kernel_launches - l0
    stats = vec.log()
    value = rows * args.steps / (ms * 1e-3)
    peak, peak_src = bench.measured_peaks()
    gbs = ALGO_BYTES_PER_DRONE_STEP * value / 1e9
    line = {"metric": "drone_steps_per_sec", "value": value, "unit": "drone-steps/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "math": args.math,
            "config": {"workload": f"drone_swarm {args.envs} envs x {args.drones} drones, max_rings=10, fixed random action tape "
                                   f"(4 x U(-1,1)), step-only incl. respawns and the 1023-tick env-wide reset (BASELINE.json configs[2])",
                       "rows": rows, "l2": f"working set {rows * 521 / 1e6:.0f} MB per step > 126 MB L2"},
            "episode_stats": stats, "clocks": sampler.summary(), "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": swarm_traffic(args),
                         "kernel": f"swarm_kernel<{args.math}>", "algorithmic_bytes_per_launch": ALGO_BYTES_PER_DRONE_STEP * rows,
                         "avg_launch_us": ms / args.steps * 1e3, "peak_source": peak_src}}
    vec.close()
    if not args.no_cpu_baseline:
        res = bench.cpu_sample(10.0, "reference", drones=args.drones)
        line["cpu_baseline"] = {"value": res["env_steps_per_s"], "unit": "drone-steps/s", "cores": res["procs"], "kind": res["kind"],
                                "sample": f"{res['envs']} envs x {args.drones} drones x {res['steps']} steps, {res['procs']} processes, {res['wall_s']:.1f} s"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
