#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "applymap or thousand or nan_and_empty or literal" > gpurun_out/pytest_map.log 2>&1
echo "pytest map exit $?"; tail -15 gpurun_out/pytest_map.log
for E in 10000 100000 1000000; do timeout 300 python tools/prof_map.py 4096 $E 16 1 | tail -1; done
timeout 300 python tools/prof_map.py 4096 100000 64 1 | tail -1
timeout 300 python tools/prof_map.py 4096 100000 16 0 | tail -1
timeout 300 python tools/prof_map.py 1024 100000 16 1 | tail -1
timeout 300 python tools/prof_map.py 200 100000 50 1 | tail -1
