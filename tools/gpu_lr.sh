#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "length_regulator" 2>&1 | tail -2
timeout 300 python tools/lr_bench.py 2>&1 | grep -E "^LR|Error" | tee gpurun_out/lr_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"length_regulator" -s 6 -c 2 -f -o gpurun_out/prof_lr python tools/lr_bench.py > gpurun_out/ncu_lr.log 2>&1
tail -1 gpurun_out/ncu_lr.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c1.csv python tools/c1_breakdown.py > gpurun_out/c1_under_ncu.log 2>&1
