#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "applymap or thousand or nan_and_empty or literal" > gpurun_out/pytest_map.log 2>&1
echo "pytest map exit $?"; tail -15 gpurun_out/pytest_map.log
timeout 600 python bench.py --steps 1 --warmup 1 --n-train 2048 --no-cpu-baseline --map-steps 50 > gpurun_out/bench_map.json 2> gpurun_out/bench_map.err
echo "bench exit $?"; tail -3 gpurun_out/bench_map.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_map.json'))
    print(d['map'])
except Exception as e: print(e)
PY
