#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "dmma or spd_factor or nll_grad_medium" > gpurun_out/pytest_gemm.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gemm.log
timeout 300 python tools/gemm_perf.py > gpurun_out/gemm_perf_ws.log 2>&1
echo "gemm_perf ws exit $?"; cat gpurun_out/gemm_perf_ws.log
