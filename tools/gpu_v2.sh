#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
grep -E "passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | tail -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['kernel_ms_per_step']); print(d['roofline']['executed_frac'], d['e2e']['value'], d['rtf_batch1'].get('ms'), d['cpu_baseline']['value'])
PY
