#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
for rep in 1 2 3; do
for lib in "" "efficient_tts_b200/libefts_b200_A.so"; do
  echo "== rep $rep lib: ${lib:-current}" >> gpurun_out/ab.log
  EFTS_B200_LIB="$lib" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print(round(d['ms_per_step'],3), d['clocks']['sm_mhz'], k['dec_conv'], k['mel_conv'], k['text_conv'], k['linear'])" >> gpurun_out/ab.log 2>&1
done; done
cat gpurun_out/ab.log
