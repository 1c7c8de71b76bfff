#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/b3.log
for rep in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print(round(d['ms_per_step'],3), round(sum(k.values()),3), d['clocks']['sm_mhz'], round(d['value']/1e6,3), round(d['e2e']['value']/1e6,3), d['gpu_launches'])" >> gpurun_out/b3.log 2>&1
done
cat gpurun_out/b3.log
