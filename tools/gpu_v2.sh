#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/diag_gemm.py > gpurun_out/diag.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/quick.log
for o in "gemm_version=1" "gemm_version=2,pair=0" "gemm_version=2,pair=1" "gemm_version=2,pair=1,chunk_kb=2" "gemm_version=2,pair=1,chunk_kb=0"; do
  timeout 600 python tools/quick_time.py C3 $o >> gpurun_out/quick.log 2>&1
done
timeout 1200 python tools/diag_precision.py bias C3 C5 > gpurun_out/precision.log 2>&1
grep -h '"cfg"' gpurun_out/diag.log | cut -c1-400; grep -E "passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | tail -12; grep QUICK gpurun_out/quick.log | cut -c1-300
grep -E "BIAS|^C[235]" gpurun_out/precision.log
