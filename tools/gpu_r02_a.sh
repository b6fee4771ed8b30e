#!/bin/sh
# round 2, GPU call A: full GPU test suite (minus the config-4 golden, generated later) + default bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/r02a_smi.txt 2>&1
nproc > gpurun_out/r02a_nproc.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 -rs --durations=30 \
  --deselect tests/test_gpu_scale.py::test_newton_delta_config4_golden \
  --deselect tests/test_gpu_scale.py::test_hybrd_config4_first_steps > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?" >> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_pytest.log
