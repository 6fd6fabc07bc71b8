mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2000 --warmup 100 2>&1 | tail -1 > gpurun_out/r01e_bench_2gpu.json; cut -c1-400 gpurun_out/r01e_bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r01e_bench_reference_2gpu.json; cut -c1-300 gpurun_out/r01e_bench_reference_2gpu.json
timeout 300 python -m pytest tests/test_shard_gloo.py -q 2>&1 | tail -2
