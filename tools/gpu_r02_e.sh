#!/bin/sh
# round 2, GPU call E: the complete GPU suite + ozaki bench + default bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 -rs --durations=15 > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -4 gpurun_out/r02e_pytest.log
timeout 600 python tools/ozaki_bench.py 2048 4096 8192 > gpurun_out/r02e_ozaki_bench.jsonl 2> gpurun_out/r02e_ozaki_bench.err
cat gpurun_out/r02e_ozaki_bench.jsonl
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
echo "bench rc=$?" >> gpurun_out/r02e_bench.err
tail -2 gpurun_out/r02e_bench.err
