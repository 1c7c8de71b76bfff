#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp.log
for o in "" "debug_mask=2" "debug_mask=3" "chunk_kb=0" "chunk_kb=0,debug_mask=3"; do
  echo "== opts: $o" >> gpurun_out/exp.log
  EFTS_BENCH_OPTS="$o" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['clocks'], {k:d['kernel_ms_per_step'][k] for k in ('dec_conv','linear','expand_gemm')})" >> gpurun_out/exp.log 2>&1
done
cat gpurun_out/exp.log
