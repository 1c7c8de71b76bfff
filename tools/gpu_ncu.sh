#!/bin/bash
# ncu --set full of the decoder conv launches (gemm2 kernel) under two option sets
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 37 -c 2 -f -o gpurun_out/prof_g2_c1 \
   python tools/quick_time.py C3 "gemm_version=2,pair=1,chunk_kb=1" > gpurun_out/ncu_c1.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 37 -c 2 -f -o gpurun_out/prof_g2_c0 \
   python tools/quick_time.py C3 "gemm_version=2,pair=1,chunk_kb=0" > gpurun_out/ncu_c0.log 2>&1
tail -2 gpurun_out/ncu_c1.log gpurun_out/ncu_c0.log
