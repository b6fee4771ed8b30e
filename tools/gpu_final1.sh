#!/bin/bash
# 1-GPU: parity tests (incl. full-size golden), smoke, bench, full sweep (config 5)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
echo "bench exit $?"; cat gpurun_out/bench_f.json; tail -5 gpurun_out/bench_f.err
timeout 900 python tools/sweep.py > gpurun_out/sweep_f.jsonl 2> gpurun_out/sweep_f.err; cut -c1-400 gpurun_out/sweep_f.jsonl; tail -3 gpurun_out/sweep_f.err
