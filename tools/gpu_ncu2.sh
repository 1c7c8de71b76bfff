#!/bin/bash
mkdir -p gpurun_out
# prenet = 8th gemm2_kernel launch of a forward (index 7); 22 gemm launches per forward; skip 2 forwards
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 51 -c 1 -f -o gpurun_out/prof_prenet \
   python tools/quick_time.py C3 > gpurun_out/ncu_prenet.log 2>&1
tail -n 2 gpurun_out/ncu_prenet.log
