#!/bin/sh
# round 2, GPU call J: balanced radix-256 digits (8 NS - 1 bits for NS (NS + 1) / 2 pairs), K blocking: tests + timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ozaki.py -m gpu -q --timeout 900 -rs --durations=8 > gpurun_out/r02j_pytest_ozaki.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02j_pytest_ozaki.log
grep -E "full size|passed|failed|^FAILED|^E  " gpurun_out/r02j_pytest_ozaki.log | head -30
timeout 600 python tools/ozaki_bench.py 4096 8192 16384 > gpurun_out/r02j_ozaki_bench.jsonl 2> gpurun_out/r02j_ozaki_bench.err
cat gpurun_out/r02j_ozaki_bench.jsonl; tail -3 gpurun_out/r02j_ozaki_bench.err
OZ_CONFIGS="0:0:0,5:3:4096,6:3:4096,7:3:4096,8:3:4096" timeout 1200 python tools/oz_route_bench.py 4096 8192 16384 > gpurun_out/r02j_route.jsonl 2> gpurun_out/r02j_route.err
cut -c1-700 gpurun_out/r02j_route.jsonl; tail -3 gpurun_out/r02j_route.err
