#!/bin/bash
mkdir -p gpurun_out
python tools/prof_map.py 4096 100000 4 1 > gpurun_out/prof_map_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:map_kernel -s 1 -c 1 -o gpurun_out/prof_map2 python tools/prof_map.py 4096 100000 4 1 > gpurun_out/ncu_map.log 2>&1
echo "ncu exit $?"; cat gpurun_out/prof_map_plain.log; tail -3 gpurun_out/ncu_map.log
