#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "bench 8gpu exit $?"; tail -3 gpurun_out/bench_8gpu.err; cut -c1-400 gpurun_out/bench_8gpu.json
