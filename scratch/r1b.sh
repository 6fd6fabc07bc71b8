mkdir -p gpurun_out
run() { python bench.py --no-cpu-baseline --no-e2e "$@" 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['reset_fraction_per_step'], d['clocks']['sm_mhz'])"; }
for f in scratch/libs/lib_v*.so; do echo "== $f"; B2D_LIBRARY=$f run; B2D_LIBRARY=$f run --steps 6000; done
echo "== parity v3"; B2D_LIBRARY=scratch/libs/lib_v3.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== parity v4"; B2D_LIBRARY=scratch/libs/lib_v4.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import torch, time
n=127*1024*1024
h=torch.empty(n,dtype=torch.uint8,pin_memory=True); d=torch.empty(n,dtype=torch.uint8,device='cuda')
for name,fn in [('d2h',lambda: h.copy_(d,non_blocking=True)),('h2d',lambda: d.copy_(h,non_blocking=True))]:
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
    print(name, n/dt/1e9,'GB/s')
# both directions at once
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
h2=torch.empty(n,dtype=torch.uint8,pin_memory=True); d2=torch.empty(n,dtype=torch.uint8,device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print('bidir each', n/dt/1e9,'GB/s')
PY
nvidia-smi -q | grep -i -A3 "pcie\|link" | head -40
