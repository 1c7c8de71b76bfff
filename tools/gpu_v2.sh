#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/quick.log
for o in "chunk_kb=1" "chunk_kb=2" "chunk_kb=0" "pair=0,chunk_kb=1"; do
  timeout 600 python tools/quick_time.py C3 $o >> gpurun_out/quick.log 2>&1
done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
grep -E "passed|failed|FAILED|pytest exit" gpurun_out/pytest_gpu.log | tail -12; grep QUICK gpurun_out/quick.log | cut -c1-160
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','clocks','kernel_ms_per_step')}); print(d['roofline']['executed_frac'], d['e2e']['value'], d['rtf_batch1'].get('ms'))
PY
