import time, numpy as np, torch, ctypes as C
from drone_b200.drone_race import DroneRace, binding
from drone_b200 import capi
n=1<<20
env=DroneRace(num_envs=n, report_interval=1<<30, seed=0, buffers="host")
env.reset(0)
rng=np.random.default_rng(0); tape=rng.uniform(-1,1,(4,n,4)).astype(np.float32)
def T(f, k=30):
    f(); torch.cuda.synchronize(); t=time.perf_counter()
    for i in range(k): f(i)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/k*1e3
print('numpy copy ms', T(lambda i=0: env.actions.__setitem__(slice(None), tape[i%4])))
print('env.step(tape) ms', T(lambda i=0: env.step(tape[i%4])))
print('env.step(self.actions) ms (no copy)', T(lambda i=0: binding.vec_step(env.c_envs)))
h=binding.vec_handle(env.c_envs); L=capi.lib()
print('b2d_vec_step_host ms', T(lambda i=0: L.b2d_vec_step_host(C.c_void_p(h), None)))
# raw copies
obs_h=env.observations; d=torch.empty((n,29),device='cuda')
ht=torch.from_numpy(obs_h)
print('pinned?', ht.is_pinned())
print('raw D2H obs ms', T(lambda i=0: ht.copy_(d, non_blocking=True)))
a_h=torch.from_numpy(env.actions); da=torch.empty((n,4),device='cuda')
print('raw H2D act ms', T(lambda i=0: da.copy_(a_h, non_blocking=True)))
