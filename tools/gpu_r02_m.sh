#!/bin/sh
# round 2, GPU call M: ncu --set full of the INT8 GEMM with 6 digits and of the two slicing kernels; quick ozaki tests; DFMA probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py -m gpu -q --timeout 600 > gpurun_out/r02m_pytest_ozaki.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02m_pytest_ozaki.log; tail -3 gpurun_out/r02m_pytest_ozaki.log
python tools/prof_oz.py 8192 6 > gpurun_out/r02m_oz_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"oz_gemm|oz_slice_mn|oz_rowmax" -s 3 -c 3 -f -o gpurun_out/r02_prof_oz6_n8192 python tools/prof_oz.py 8192 6 > gpurun_out/r02m_ncu_oz.log 2>&1
echo "ncu oz exit $?"
ncu -i gpurun_out/r02_prof_oz6_n8192.ncu-rep --page details > gpurun_out/r02_prof_oz6_n8192.txt 2>/dev/null
ncu -i gpurun_out/r02_prof_oz6_n8192.ncu-rep --page raw --csv > gpurun_out/r02_prof_oz6_n8192.raw.csv 2>/dev/null
python -c "
import ctypes,sys
sys.path.insert(0,'.')
from sympgpr_b200 import _lib
r=ctypes.c_double(0.0)
_lib.check(_lib.lib().sgp_bench_dfma(_lib.context(0).handle,3,ctypes.byref(r)),'dfma')
print('dfma dp instr/s', r.value, 'nominal', 148*64*1.965e9)
"
