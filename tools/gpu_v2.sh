#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/exp.log
for o in "" "fp32_master=0"; do
  echo "== opts: $o" >> gpurun_out/exp.log
  EFTS_BENCH_OPTS="$o" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['kernel_ms_per_step'])" >> gpurun_out/exp.log 2>&1
done
timeout 1200 python tools/diag_precision.py C3 C5 fp64 > gpurun_out/precision.log 2>&1
grep -E "passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | tail -8; cat gpurun_out/exp.log; grep -E "^C[235]" gpurun_out/precision.log | cut -c1-700
