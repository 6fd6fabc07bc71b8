mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2000 --warmup 100 2> gpurun_out/r01g_bench_8gpu.err | tail -1 > gpurun_out/r01g_bench_8gpu.json
cut -c1-300 gpurun_out/r01g_bench_8gpu.json; tail -2 gpurun_out/r01g_bench_8gpu.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 2000 --warmup 100 2>/dev/null | tail -1 > gpurun_out/r01g_bench_4gpu.json
cut -c1-200 gpurun_out/r01g_bench_4gpu.json
