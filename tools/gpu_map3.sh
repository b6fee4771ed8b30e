#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --n-train 1024 --steps 1 --warmup 3 --no-cpu-baseline --map-steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())['map']
print({k:d[k] for k in ('value','sweeps_per_orbit_step','lane_utilisation','newton_refstart_value','hybrd_value','unconverged')}, d['roofline']['frac'])"; }
run SGP_MAP_GTHR=12 SGP_MAP_SLICE=16
run SGP_MAP_GTHR=12 SGP_MAP_SLICE=64
run SGP_MAP_GTHR=33 SGP_MAP_SLICE=16
run SGP_MAP_GTHR=24 SGP_MAP_SLICE=64
run SGP_MAP_GTHR=4 SGP_MAP_SLICE=64
