#!/bin/bash
# Round-1 evidence for profiles/: plain bench, ncu launch list of the same command (capped), DRAM traffic of the
# DMMA kernels at the headline size, ncu --set full of potrf_ll / map / fill / gradient kernels.
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
echo "bench exit $?"
# 1. traffic + duration of every kernel of one NLL+gradient evaluation at n = 32768 (single-pass metrics: no replay)
python tools/prof_nll.py 16384 > gpurun_out/p_nll16384.log 2>&1 &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/traffic_nll16384.csv python tools/prof_nll.py 16384 > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit $?"; wc -l gpurun_out/traffic_nll16384.csv
# 2. full captures
python tools/prof_nll.py 8192 > gpurun_out/p_nll8192.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_ll -c 1 -f -o gpurun_out/prof_potrf_ll_e python tools/prof_nll.py 8192 > gpurun_out/ncu_potrf_ll.log 2>&1
echo "ncu potrf_ll exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fill_hess|grad_hess" -c 2 -f -o gpurun_out/prof_fill_grad python tools/prof_nll.py 8192 > gpurun_out/ncu_fill_grad.log 2>&1
echo "ncu fill/grad exit $?"
python tools/prof_map.py 4096 100000 4 3 > gpurun_out/prof_map_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_kernel -s 1 -c 1 -f -o gpurun_out/prof_map_e python tools/prof_map.py 4096 100000 4 3 > gpurun_out/ncu_map.log 2>&1
echo "ncu map exit $?"; cat gpurun_out/prof_map_plain.log
for f in prof_potrf_ll_e prof_fill_grad prof_map_e; do
  ncu -i gpurun_out/$f.ncu-rep --page details > gpurun_out/$f.txt 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,sm__cycles_active.avg > gpurun_out/$f.raw.csv 2>/dev/null
done
# 3. launch list of the default bench command (capped: probe + warm-up evaluations are identical to the timed ones)
timeout 1000 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 1700 --log-file gpurun_out/launches_bench_e.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
echo "ncu launch list exit $?"; wc -l gpurun_out/launches_bench_e.csv
