#!/bin/sh
# round 2, GPU call D: + streamed sub-diagonal tile, DMMA ILP, L2 prefetch; correctness subset, timings, trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x \
  -k "spd_factor or nll or fit or positive or dof2_nll or gemm" > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -3 gpurun_out/r02d_pytest.log
timeout 300 python tools/potrf_perf.py 200 1024 2048 4096 8192 16384 > gpurun_out/r02d_perf.jsonl 2> gpurun_out/r02d_perf.err
cat gpurun_out/r02d_perf.jsonl
SGP_LL_TRACE_FILE=gpurun_out/r02d_trace_4096.txt timeout 300 python tools/ll_trace.py 4096 > gpurun_out/r02d_ll_trace_4096.log 2>&1
tail -14 gpurun_out/r02d_ll_trace_4096.log
