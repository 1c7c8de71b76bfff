#!/bin/bash
# Fused-B conv kernel: parity suite, A/B timing against the three-MMA kernel, precision at full size.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|pytest exit" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python tools/exp_loads.py "fuse_b=0" "fuse_b=1" "fuse_b=0" "fuse_b=1" 2>&1 | grep -E "EXP|Error|error" | tee gpurun_out/exp_fuse.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_fuse.json 2> gpurun_out/bench_fuse.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_fuse.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['kernel_ms_per_step']); print(d['roofline']['executed_frac'], d['e2e']['value'], d['rtf_batch1'].get('ms'))
PY
timeout 1200 python tools/diag_precision.py bias C3 C5 fp64 > gpurun_out/precision.log 2>&1
grep -E "^BIAS|^C[235]" gpurun_out/precision.log | cut -c1-700
