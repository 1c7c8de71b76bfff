#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/quick.log
for o in "gemm_version=2,pair=1,chunk_kb=1" "gemm_version=2,pair=1,chunk_kb=2" "gemm_version=2,pair=1,chunk_kb=4" "gemm_version=2,pair=1,chunk_kb=0" "gemm_version=2,pair=0,chunk_kb=2"; do
  timeout 600 python tools/quick_time.py C3 $o >> gpurun_out/quick.log 2>&1
done
EFTS_OPTS="chunk_kb=2" timeout 1800 python tools/diag_precision.py bias C3 C5 fp64 > gpurun_out/precision_c2.log 2>&1
grep -E "passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | tail -12; grep QUICK gpurun_out/quick.log | cut -c1-160
grep -E "^BIAS|^C[235]" gpurun_out/precision_c2.log
