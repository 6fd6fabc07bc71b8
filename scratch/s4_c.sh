run() { B2D_LIBRARY=/root/repo/scratch/libs/lib_$1.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 1500 --warmup 50 $2 2>gpurun_out/err_$1.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $2', d['ms_per_step']*1e3, d['roofline']['frac'])"; grep "b2d timing" gpurun_out/err_$1.txt; }
run fenceblk
run fenceblk "--launch single"
B2D_TRACE_FILE=gpurun_out/trace_tape.csv run timing
B2D_TRACE_FILE=gpurun_out/trace_single.csv run timing "--launch single"
