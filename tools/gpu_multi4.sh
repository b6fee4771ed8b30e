#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 3 --warmup 3 --no-sweep > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
echo "bench 4gpu exit $?"; tail -2 gpurun_out/bench_4gpu.err
