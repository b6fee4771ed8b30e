#!/bin/sh
# round 2, GPU call H: the complete GPU suite, the default bench (with the all-stages INT8 leg), ncu --set full of the INT8 GEMM
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 -rs --durations=12 > gpurun_out/r02h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -5 gpurun_out/r02h_pytest.log
timeout 1800 python bench.py --steps 3 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
echo "bench rc=$?" >> gpurun_out/r02h_bench.err
tail -2 gpurun_out/r02h_bench.err
python tools/prof_oz.py 8192 7 > gpurun_out/r02h_oz_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm -s 1 -c 1 -f -o gpurun_out/r02_prof_oz_gemm_n8192 python tools/prof_oz.py 8192 7 > gpurun_out/r02h_ncu_oz.log 2>&1
echo "ncu oz exit $?"
ncu -i gpurun_out/r02_prof_oz_gemm_n8192.ncu-rep --page details > gpurun_out/r02_prof_oz_gemm_n8192.txt 2>/dev/null
ncu -i gpurun_out/r02_prof_oz_gemm_n8192.ncu-rep --page raw --csv > gpurun_out/r02_prof_oz_gemm_n8192.raw.csv 2>/dev/null
