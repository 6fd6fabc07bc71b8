nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active,temperature.gpu --format=csv,noheader -lms 50 > gpurun_out/clk_tape.csv &
NS=$!
sleep 1
timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 30000 --warmup 50 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tape 30000', d['ms_per_step']*1e3, d['roofline']['frac'], d['clocks'])"
sleep 1
echo MARK >> gpurun_out/clk_tape.csv
B2D_LIBRARY=/root/repo/scratch/libs/lib_skipmath.so timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 30000 --warmup 50 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skipmath 30000', d['ms_per_step']*1e3, d['roofline']['frac'], d['clocks'])"
kill $NS
for n in 65536 131072 262144 524288; do
timeout 120 python bench.py --no-e2e --no-cpu-baseline --steps 4000 --warmup 50 --envs-per-gpu $n 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('envs $n', d['ms_per_step']*1e3, d['ms_per_step']*1e3*1048576/$n)"
done
