import os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
from drone_b200.drone_race import DroneRace
n = 1 << 20
env = DroneRace(num_envs=n, report_interval=1 << 30, seed=0, buffers="host", device=0, math="fast")
env.reset(0)
rng = np.random.default_rng(1)
ht = rng.uniform(-1, 1, size=(4, n, 4)).astype(np.float32)
for k in range(5): env.step(ht[k % 4])
res = {"split": [], "nosplit": []}
for rnd in range(6):
    for mode in ("split", "nosplit"):
        if mode == "nosplit": os.environ["B2D_HOST_NO_SPLIT"] = "1"
        else: os.environ.pop("B2D_HOST_NO_SPLIT", None)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        K = 40
        for k in range(K): env.step(ht[k % 4])
        torch.cuda.synchronize(); res[mode].append((time.perf_counter() - t0) / K * 1e3)
for m, v in res.items(): print(m, ["%.3f" % x for x in v], "min %.3f" % min(v))
env.close()
