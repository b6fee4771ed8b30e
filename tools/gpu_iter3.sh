#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "quality or iterate" > gpurun_out/pytest_q.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_q.log
for b in 0 1; do
  echo "SGP_LL_BLOCKED=$b"
  SGP_LL_BLOCKED=$b SWEEP_MIN_N=8192 SWEEP_MAX_N=16384 SWEEP_NO_MAP=1 timeout 600 python tools/sweep.py 2>&1 | cut -c1-330
done
