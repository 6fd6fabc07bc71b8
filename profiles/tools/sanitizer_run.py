#!/usr/bin/env python
"""Small runs of every hot kernel for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python profiles/tools/sanitizer_run.py
Covers: race step (tape launches with overlap, stand-alone launches, resets), swarm step at A = 16 and 64 (overlap,
respawns, an env-wide reset at tick 1023 via a forced tick), the rollout kernel (K = 6), the fused policy step,
the advantage kernel, state hooks, the host-buffer pipeline."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from drone_b200.vec import RaceVec, SwarmVec
from drone_b200.rollout import DeviceRollout, DronePolicy
from drone_b200.advantage import compute_puff_advantage

torch.manual_seed(0)
n = 4133
v = RaceVec(n, seed=3, max_moves=25)
v.reset(3)
tape = torch.rand((8, n, 4), device='cuda') * 2 - 1
v.step_tape(tape, 0, 40)
for t in range(10):
    v.step(tape[t % 8])
blob = v.get_state(range(0, 64)); v.put_state(blob, range(0, 64)); v.observe()
print('race', v.log()['n'], v.step_count)
ro = DeviceRollout(v, DronePolicy().cuda(), horizon=6)
ro.collect(); ro.collect()
print('rollout', ro.policy_impl, float(ro.values.abs().mean()))
v.close()

for A, envs in ((16, 130), (64, 35)):
    s = SwarmVec(envs, A, 10, math='fast', seed=4)
    s.reset(4)
    st = torch.rand((4, envs * A, 4), device='cuda') * 2 - 1
    for t in range(30):
        s.step(st[t % 4])
    b = s.get_state(range(envs)); s.put_state(b, range(envs)); s.observe()
    print('swarm', A, s.log()['n'])
    s.close()

K, N = 16, 1000
val = torch.randn(K, N, device='cuda'); rew = torch.randn(K, N, device='cuda'); done = (torch.rand(K, N, device='cuda') < 0.1).float()
adv = compute_puff_advantage(val, rew, done, torch.ones_like(val), torch.zeros_like(val), 0.99, 0.95, 1.0, 1.0, time_major=True)
print('advantage', float(adv.abs().mean()))

from drone_b200.drone_race import DroneRace
e = DroneRace(num_envs=140_003, seed=1, report_interval=1 << 30)
e.reset(1)
a = np.random.default_rng(0).uniform(-1.5, 1.5, size=(140_003, 4)).astype(np.float32)
for t in range(3):
    e.step(a)
print('host', float(np.abs(e.observations).sum()))
e.close()
torch.cuda.synchronize()
print('done')
