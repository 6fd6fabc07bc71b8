for n in skipmath timing; do
  B2D_LIBRARY=/root/repo/scratch/libs/lib_$n.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 2>gpurun_out/err_$n.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n', d['ms_per_step']*1e3, d['roofline']['frac'])"
  grep "b2d timing" gpurun_out/err_$n.txt
  B2D_LIBRARY=/root/repo/scratch/libs/lib_$n.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 --launch single 2>gpurun_out/err_$n.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n single', d['ms_per_step']*1e3, d['roofline']['frac'])"
  grep "b2d timing" gpurun_out/err_$n.txt
done
