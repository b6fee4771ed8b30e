#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "spd_factor or nll or fit or positive" > gpurun_out/pytest_ll.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_ll.log
timeout 300 python tools/ll_trace.py 4096 > gpurun_out/ll_trace_4096.log 2>&1; tail -32 gpurun_out/ll_trace_4096.log | cut -c1-150
SWEEP_MAX_N=${SWEEP_MAX_N:-8192} SWEEP_NO_MAP=1 timeout 600 python tools/sweep.py > gpurun_out/sweep_small.jsonl 2> gpurun_out/sweep_small.err; cut -c1-330 gpurun_out/sweep_small.jsonl; tail -3 gpurun_out/sweep_small.err
