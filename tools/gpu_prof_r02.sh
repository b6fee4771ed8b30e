#!/bin/bash
# Round-2 evidence for profiles/: DRAM traffic of one evaluation at the headline size, ncu --set full of potrf_ll at the mid
# size VERDICT named (n = 8192), of the grouped trtri GEMM and of the opt-in INT8 (tcgen05) GEMM, launch list of the bench.
mkdir -p gpurun_out
python tools/prof_nll.py 16384 > gpurun_out/r02p_nll16384.log 2>&1 &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_ncu_traffic_nll_n32768.csv python tools/prof_nll.py 16384 > gpurun_out/r02p_ncu_traffic.log 2>&1
echo "ncu traffic exit $?"; wc -l gpurun_out/r02_ncu_traffic_nll_n32768.csv
python tools/prof_nll.py 4096 > gpurun_out/r02p_nll4096.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_ll -c 1 -f -o gpurun_out/r02_prof_potrf_ll_n8192 python tools/prof_nll.py 4096 > gpurun_out/r02p_ncu_potrf.log 2>&1
echo "ncu potrf_ll exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_ws -s 10 -c 2 -f -o gpurun_out/r02_prof_trtri_grouped_n8192 python tools/prof_nll.py 4096 > gpurun_out/r02p_ncu_trtri.log 2>&1
echo "ncu trtri exit $?"
python tools/prof_oz.py 4096 7 > gpurun_out/r02p_oz_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm -s 1 -c 1 -f -o gpurun_out/r02_prof_oz_gemm_n4096 python tools/prof_oz.py 4096 7 > gpurun_out/r02p_ncu_oz.log 2>&1
echo "ncu oz exit $?"
for f in r02_prof_potrf_ll_n8192 r02_prof_trtri_grouped_n8192 r02_prof_oz_gemm_n4096; do
  ncu -i gpurun_out/$f.ncu-rep --page details > gpurun_out/$f.txt 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
# launch list of the default bench command (capped: the DGEMM/memset probes, warm-up and timed evaluations come first)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 400 --log-file gpurun_out/r02_ncu_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-sweep --no-configs --no-map --no-cpu-baseline > gpurun_out/r02p_bench_under_ncu.json 2> gpurun_out/r02p_bench_under_ncu.err
echo "ncu launch list exit $?"; wc -l gpurun_out/r02_ncu_launches_bench.csv
