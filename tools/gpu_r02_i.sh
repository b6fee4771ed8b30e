#!/bin/sh
# round 2, GPU call I: factor-only INT8 recursion (value-only evaluations), launch list of one evaluation on the INT8 route
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ozaki.py -m gpu -q --timeout 900 -rs --durations=8 > gpurun_out/r02i_pytest_ozaki.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_pytest_ozaki.log
tail -12 gpurun_out/r02i_pytest_ozaki.log
OZ_CONFIGS="0:0:0,7:3:4096" timeout 1200 python tools/oz_route_bench.py 4096 8192 16384 > gpurun_out/r02i_route.jsonl 2> gpurun_out/r02i_route.err
cut -c1-600 gpurun_out/r02i_route.jsonl; tail -3 gpurun_out/r02i_route.err
SGP_OZAKI=7:3:4096 python tools/prof_nll.py 16384 > gpurun_out/r02i_nll16384_int8.log 2>&1 &&
SGP_OZAKI=7:3:4096 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_ncu_traffic_nll_n32768_int8.csv python tools/prof_nll.py 16384 > gpurun_out/r02i_ncu_traffic.log 2>&1
echo "ncu traffic exit $?"; wc -l gpurun_out/r02_ncu_traffic_nll_n32768_int8.csv
