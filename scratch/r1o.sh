timeout 900 python -m pytest tests/test_advantage_gpu.py tests/test_wrappers_gpu.py -m gpu -x -q 2>&1 | tail -12
python - <<'PY'
import torch, time
from drone_b200.advantage import compute_puff_advantage
N,K=1<<20,128
x=[torch.rand((K,N),device='cuda') for _ in range(4)]; adv=torch.zeros((K,N),device='cuda'); pr=torch.zeros(N,device='cuda')
for tm,name in [(True,'time-major [128, 1M]')]:
    for _ in range(3): compute_puff_advantage(*x,adv,0.99,0.95,1.0,1.0,time_major=tm,priority=pr)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(10): compute_puff_advantage(*x,adv,0.99,0.95,1.0,1.0,time_major=tm,priority=pr)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/10
    print(name, ms,'ms', 5*4*N*K/ms/1e6,'GB/s algorithmic')
y=[t.T.contiguous() for t in x]; adv2=torch.zeros((N,K),device='cuda')
for _ in range(3): compute_puff_advantage(*y,adv2,0.99,0.95,1.0,1.0)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
for _ in range(10): compute_puff_advantage(*y,adv2,0.99,0.95,1.0,1.0)
e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/10
print('row-major [1M, 128] (reference layout)', ms,'ms', 5*4*N*K/ms/1e6,'GB/s algorithmic')
PY
