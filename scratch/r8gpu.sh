mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2000 --warmup 100 > gpurun_out/bench_r01_8gpu.json 2> gpurun_out/bench_r01_8gpu.err
tail -1 gpurun_out/bench_r01_8gpu.json | cut -c1-400; tail -3 gpurun_out/bench_r01_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 2000 --warmup 100 --no-e2e > gpurun_out/bench_r01_4gpu.json 2>/dev/null
tail -1 gpurun_out/bench_r01_4gpu.json | cut -c1-200
