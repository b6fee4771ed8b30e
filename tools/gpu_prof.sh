#!/bin/bash
# GEMM kernel sweep + ncu launch list of one NLL+grad evaluation + full capture of one big GEMM launch.
mkdir -p gpurun_out
python tools/gemm_perf.py > gpurun_out/gemm_perf.log 2>&1
echo "gemm_perf exit $?"; cat gpurun_out/gemm_perf.log
python tools/prof_nll.py 16384 > gpurun_out/prof_nll_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_nll16384.csv python tools/prof_nll.py 16384 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
python tools/gemm_perf.py --one > gpurun_out/gemm_one_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -s 1 -c 1 -o gpurun_out/prof_gemm_big python tools/gemm_perf.py --one > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full.log
