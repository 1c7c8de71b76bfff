#!/bin/bash
# First GPU session: GEMM variants in isolated processes, the parity suite, quick timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python tools/diag_gemm.py > gpurun_out/diag.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for cfg in "C2 0 1" "C3 0 1" "C3 0 0" "C3 1 1"; do
  timeout 600 python tools/quick_time.py $cfg >> gpurun_out/quick.log 2>&1
done
tail -5 gpurun_out/diag.log; tail -40 gpurun_out/pytest_gpu.log; grep QUICK gpurun_out/quick.log
