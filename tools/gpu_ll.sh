#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "spd_factor or positive_definite" > gpurun_out/pytest_ll.log 2>&1
echo "pytest spd exit $?"; tail -15 gpurun_out/pytest_ll.log
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest all exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-map --no-cpu-baseline > gpurun_out/bench_ll.json 2> gpurun_out/bench_ll.err
echo "bench exit $?"; tail -3 gpurun_out/bench_ll.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_ll.json'))
    print(d['ms_per_step'], {k:(round(v['ms'],2)) for k,v in d['stages'].items()}, d['result'])
except Exception as e: print(e)
PY
