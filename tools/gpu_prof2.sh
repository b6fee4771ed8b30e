#!/bin/bash
# ncu captures for profiles/: WS GEMM (MN/MN and K/K), potrf_ll, and the launch list of the default bench command
mkdir -p gpurun_out
python tools/gemm_one.py 0 0 0 74 64 8192 > gpurun_out/p_gemm_mn.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_ws -s 1 -c 1 -o gpurun_out/prof_gemm_ws_mn python tools/gemm_one.py 0 0 0 74 64 8192 > gpurun_out/ncu_gemm_mn.log 2>&1
echo "ncu gemm mn exit $?"
python tools/gemm_one.py 1 1 2 64 64 8192 > gpurun_out/p_gemm_kk.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_ws -s 1 -c 1 -o gpurun_out/prof_gemm_ws_kk python tools/gemm_one.py 1 1 2 64 64 8192 > gpurun_out/ncu_gemm_kk.log 2>&1
echo "ncu gemm kk exit $?"
python tools/prof_nll.py 8192 > gpurun_out/p_nll8192.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:potrf_ll -c 1 -o gpurun_out/prof_potrf_ll python tools/prof_nll.py 8192 > gpurun_out/ncu_potrf_ll.log 2>&1
echo "ncu potrf_ll exit $?"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
echo "ncu launch list exit $?"
wc -l gpurun_out/launches_bench.csv
