import sys, time; sys.path.insert(0,'.')
import torch
from drone_b200.vec import RaceVec
n=1<<20
vec=RaceVec(n, seed=0)
g=torch.Generator().manual_seed(1234)
tape=(torch.rand((16,n,4),generator=g)*2-1).cuda()
vec.reset(0)
for t in range(100): vec.step(tape[t%16])
torch.cuda.synchronize()
def timeit(fn, reps):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
def loop16():
    for t in range(16): vec.step(tape[t])
print('eager us/step', timeit(loop16, 100)/16*1e3)
s=torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    loop16()
    gr=torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        loop16()
torch.cuda.synchronize()
print('graph us/step', timeit(gr.replay, 100)/16*1e3)
# kernel-only timing per launch using events around single steps
ts=[]
for t in range(50):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); vec.step(tape[t%16]); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
ts.sort(); print('single-step sync us median', ts[len(ts)//2], 'min', ts[0])
vec.profile_kernels(True)
for t in range(200): vec.step(tape[t%16])
print(vec.profile_kernels(False))
