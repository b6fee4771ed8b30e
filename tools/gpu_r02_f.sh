#!/bin/sh
# round 2, GPU call F: INT8 route (merged-N MMAs, general sliced products, factor + inverse recursion): tests, GEMM and route timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ozaki.py -m gpu -q --timeout 900 -x -rs --durations=10 > gpurun_out/r02f_pytest_ozaki.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_pytest_ozaki.log
tail -25 gpurun_out/r02f_pytest_ozaki.log
timeout 600 python tools/ozaki_bench.py 2048 4096 8192 16384 > gpurun_out/r02f_ozaki_bench.jsonl 2> gpurun_out/r02f_ozaki_bench.err
cat gpurun_out/r02f_ozaki_bench.jsonl; tail -3 gpurun_out/r02f_ozaki_bench.err
timeout 1200 python tools/oz_route_bench.py 4096 8192 16384 > gpurun_out/r02f_route.jsonl 2> gpurun_out/r02f_route.err
cut -c1-420 gpurun_out/r02f_route.jsonl; tail -3 gpurun_out/r02f_route.err
