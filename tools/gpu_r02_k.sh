#!/bin/sh
# round 2, GPU call K: complete GPU suite + default bench on the state with balanced radix-256 digits
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 -rs --durations=12 > gpurun_out/r02k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
tail -5 gpurun_out/r02k_pytest.log
timeout 1800 python bench.py --steps 3 --warmup 3 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
echo "bench rc=$?" >> gpurun_out/r02k_bench.err
tail -2 gpurun_out/r02k_bench.err
OZ_CONFIGS="6:3:4096,7:3:4096" timeout 600 python tools/oz_route_bench.py 16384 > gpurun_out/r02k_route.jsonl 2> gpurun_out/r02k_route.err
cut -c1-500 gpurun_out/r02k_route.jsonl
