"""e2e (host-buffer) step time vs number of pipeline chunks; run under gpurun."""
import os, sys, time, subprocess, json
code = r'''
import os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
from drone_b200.drone_race import DroneRace
n = 1 << 20
env = DroneRace(num_envs=n, report_interval=1 << 30, seed=0, buffers="host", device=0, math="fast")
env.reset(0)
rng = np.random.default_rng(1)
ht = rng.uniform(-1, 1, size=(4, n, 4)).astype(np.float32)
for k in range(5): env.step(ht[k % 4])
torch.cuda.synchronize(); t0 = time.perf_counter()
K = 100
for k in range(K): env.step(ht[k % 4])
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
print(os.environ.get("B2D_HOST_CHUNKS", "default"), "ms/step %.3f" % (dt * 1e3), "env-steps/s %.3e" % (n / dt))
env.close()
'''
for c in ("", "6", "8", "10", "12"):
    env = dict(os.environ)
    if c: env["B2D_HOST_CHUNKS"] = c
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print((r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
