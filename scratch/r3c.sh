(time timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_rollout_gpu.py -m gpu -x -q) 2>&1 | tail -25
for pr in tf32 fp32; do python bench_rollout.py --precision $pr 2>&1 | tail -1 | tee gpurun_out/r3c_rollout_$pr.json | cut -c1-400; done
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from drone_b200.rollout import DronePolicy
from drone_b200.policy import FusedPolicyStep
for D in (29, 41):
  for pr in ("tf32", "fp32"):
    n = 1 << 20
    p = DronePolicy(obs_dim=D).cuda()
    obs = torch.randn((n, D), device='cuda'); rew = torch.randn(n, device='cuda'); term = torch.zeros(n, dtype=torch.uint8, device='cuda'); act = torch.zeros((n, 4), device='cuda')
    K = 8
    so = torch.zeros((K, n, D), device='cuda'); sa = torch.zeros((K, n, 4), device='cuda'); s1 = torch.zeros((4, K, n), device='cuda')
    f = FusedPolicyStep(p, obs, rew, term, act, precision=pr)
    for k in range(3): f.act(so[k], sa[k], s1[0, k], s1[1, k], s1[2, k], s1[3, k])
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    R = 40
    for r in range(R): k = r % K; f.act(so[k], sa[k], s1[0, k], s1[1, k], s1[2, k], s1[3, k])
    e1.record(); torch.cuda.synchronize(); us = e0.elapsed_time(e1) / R * 1e3
    B = n * (2 * D * 4 + 16 + 16 + 16 + 5)
    print(f"policy_act D={D} {pr}: {us:.1f} us per 1M rows, {B/us/1e3:.0f} GB/s algorithmic, {n/us*1e6:.3e} rows/s")
PY
ncu --set full --clock-control none --import-source on -k regex:policy_act -s 3 -c 1 -o gpurun_out/r3c_policy python bench_rollout.py --horizon 8 --replays 2 --no-graph > gpurun_out/r3c_ncu.log 2>&1; tail -2 gpurun_out/r3c_ncu.log | cut -c1-300
