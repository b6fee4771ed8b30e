#!/bin/sh
# round 2, GPU call B: new diagonal-tile path of potrf_ll (pipelined block inverses, DMMA trailing update + tile inverse)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x \
  -k "spd_factor or nll or fit or positive or dof2_nll or gemm" > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -3 gpurun_out/r02b_pytest.log
timeout 300 python tools/potrf_perf.py > gpurun_out/r02b_perf.jsonl 2> gpurun_out/r02b_perf.err
for v in unr2 unr4; do
  SYMPGPR_B200_LIB=$PWD/sympgpr_b200/_variants/libsympgpr_b200_$v.so timeout 300 python tools/potrf_perf.py 200 2048 4096 >> gpurun_out/r02b_perf.jsonl 2>> gpurun_out/r02b_perf.err
done
cat gpurun_out/r02b_perf.jsonl
SGP_LL_TRACE_FILE=gpurun_out/r02b_trace_4096.txt timeout 300 python tools/ll_trace.py 4096 > gpurun_out/r02b_ll_trace_4096.log 2>&1
timeout 900 python -m pytest tests/test_reference_scripts.py -m gpu -q --timeout 600 -k "03_henon" > gpurun_out/r02b_pytest03.log 2>&1
tail -3 gpurun_out/r02b_pytest03.log
