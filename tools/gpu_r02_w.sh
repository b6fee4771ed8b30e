#!/bin/sh
# round 2, GPU call W: complete GPU suite + default bench on the final kernels (state "w")
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 -rs --durations=12 > gpurun_out/r02w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02w_pytest.log
tail -5 gpurun_out/r02w_pytest.log
timeout 1800 python bench.py --steps 3 --warmup 3 > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
echo "bench rc=$?" >> gpurun_out/r02w_bench.err
tail -2 gpurun_out/r02w_bench.err
