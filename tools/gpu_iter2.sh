#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
for b in 0 1; do
  echo "SGP_LL_BLOCKED=$b"
  SGP_LL_BLOCKED=$b timeout 300 python tools/quick_perf.py 16384 2 2>&1 | tail -3
  SGP_LL_BLOCKED=$b timeout 300 python tools/quick_perf.py 8192 2 2>&1 | tail -2
done
