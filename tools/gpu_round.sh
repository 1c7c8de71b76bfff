#!/bin/bash
# Round GPU session: parity suite, smoke, bench (both arms), ncu launch list + full captures, precision.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 37 -c 2 -f -o gpurun_out/prof_dec_conv \
   python tools/quick_time.py C3 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"reconstruct_alignment_rows|imv_scan_block|aligned_positions_block" -s 6 -c 3 -f -o gpurun_out/prof_imv_chain \
   python tools/quick_time.py C3 > gpurun_out/ncu_imv.log 2>&1
timeout 1200 python tools/diag_precision.py bias C3 C5 fp64 > gpurun_out/precision.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_ref.json | cut -c1-200; cat gpurun_out/bench.json | cut -c1-300; tail -2 gpurun_out/bench.err
grep -E "BIAS|^C[235]" gpurun_out/precision.log | cut -c1-400
