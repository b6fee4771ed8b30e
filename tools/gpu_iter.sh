#!/bin/bash
# Iteration call: parity tests, potrf trace at N=4096, small-size sweep, bench with map leg.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/ll_trace.py 4096 > gpurun_out/ll_trace_4096.log 2>&1; tail -30 gpurun_out/ll_trace_4096.log
SWEEP_MAX_N=${SWEEP_MAX_N:-8192} SWEEP_NO_MAP=1 timeout 600 python tools/sweep.py > gpurun_out/sweep_small.jsonl 2> gpurun_out/sweep_small.err; cat gpurun_out/sweep_small.jsonl; tail -3 gpurun_out/sweep_small.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
