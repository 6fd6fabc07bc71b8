#!/usr/bin/env python
"""bench_rollout.py -- secondary measurement (BASELINE.json configs[3]): full on-device rollout,
torch MLP policy (29 -> 128 GELU -> 4 + value) + CUDA-graph K=128 step loop, 1M race envs.
Reports env-steps/s including the policy forward, sampling and experience stores."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--horizon", type=int, default=128)
    ap.add_argument("--replays", type=int, default=8)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"],
                    help="fused policy GEMM precision: tf32 = tensor cores, the reference's torch.set_float32_matmul_precision('high')")
    ap.add_argument("--policy", default="fused", choices=["fused", "torch"],
                    help="fused: one CUDA kernel per policy step (drone_b200.policy); torch: library GEMMs + elementwise ops")
    args = ap.parse_args()
    import torch
    from drone_b200.rollout import DeviceRollout, DronePolicy
    from drone_b200.vec import RaceVec
    torch.manual_seed(0)
    torch.set_float32_matmul_precision("high")  # the reference's setting (pufferl.py:55); affects --policy torch
    vec = RaceVec(args.envs, seed=0)
    vec.reset(0)
    policy = DronePolicy().cuda()
    ro = DeviceRollout(vec, policy, horizon=args.horizon, use_graph=not args.no_graph,
                       autocast=torch.bfloat16 if args.bf16 else None,
                       policy_impl="torch" if args.bf16 else args.policy, precision=args.precision)
    ro.collect()
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(bench.physical_gpu_index(0))
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.replays):
        ro.collect()
    ev1.record()
    torch.cuda.synchronize()
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    steps = args.replays * args.horizon
    line = {"metric": "rollout_env_steps_per_sec", "value": args.envs * steps / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": 1,
            "steps": steps, "ms_per_step": ms / steps, "higher_is_better": True, "dtype": "f32 env / " + ("bf16" if args.bf16 else "f32") + " policy",
            "policy_impl": ro.policy_impl, "policy_precision": args.precision if ro.policy_impl == "fused" else None, "gpu_launches": (2 if ro.policy_impl == "fused" else None) and 2 * steps,
            "data": "synthetic",
            "config": {"workload": f"on-device rollout: DronePolicy MLP ({ro.policy_impl}) + race env step, {args.envs} envs, K={args.horizon} per "
                                   f"{'CUDA-graph replay' if not args.no_graph else 'eager loop'} (BASELINE.json configs[3])"},
            "episode_stats": vec.log(), "clocks": sampler.summary()}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
