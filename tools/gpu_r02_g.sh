#!/bin/sh
# round 2, GPU call G: INT8 GEMM with decoupled roles (producer runs ahead across tiles, per-slice ring slots): tests + timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ozaki.py -m gpu -q --timeout 900 -rs --durations=8 > gpurun_out/r02g_pytest_ozaki.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_pytest_ozaki.log
tail -25 gpurun_out/r02g_pytest_ozaki.log
timeout 600 python tools/ozaki_bench.py 2048 4096 8192 16384 > gpurun_out/r02g_ozaki_bench.jsonl 2> gpurun_out/r02g_ozaki_bench.err
cat gpurun_out/r02g_ozaki_bench.jsonl; tail -3 gpurun_out/r02g_ozaki_bench.err
OZ_CONFIGS="0:0:0,7:3:4096,6:3:4096,8:3:4096" timeout 1200 python tools/oz_route_bench.py 2048 4096 8192 16384 > gpurun_out/r02g_route.jsonl 2> gpurun_out/r02g_route.err
cut -c1-420 gpurun_out/r02g_route.jsonl; tail -3 gpurun_out/r02g_route.err
