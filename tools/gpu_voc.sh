#!/bin/bash
# Vocoder evidence: smoke, launch list and ncu --set full of the dominant generator layers, bench with the vocoder block.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log; tail -4 gpurun_out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_voc.csv python tools/voc_bench.py > gpurun_out/voc_under_ncu.log 2>&1
# B = 16 forward: skip the 23 B = 1 forwards + 3 warm-ups (77 GEMM launches each); capture one stage-2 and one stage-4 layer
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 2030 -c 1 -f -o gpurun_out/prof_voc_stage2 python tools/voc_bench.py > gpurun_out/ncu_voc2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 2075 -c 1 -f -o gpurun_out/prof_voc_stage4 python tools/voc_bench.py > gpurun_out/ncu_voc4.log 2>&1
tail -1 gpurun_out/ncu_voc2.log; tail -1 gpurun_out/ncu_voc4.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['rtf_batch1']['ms'], d.get('vocoder'))
PY
