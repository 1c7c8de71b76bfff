#!/bin/bash
# IMV-chain kernels: parity suite, A/B timing of the kernel generations, ncu captures of the HBM-bound kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|pytest exit" gpurun_out/pytest_gpu.log | tail -8
EFTS_BENCH_OPTS="imv_version=1,reconstruct_version=2" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_imv_old.json 2> gpurun_out/bench_imv_old.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_imv_new.json 2> gpurun_out/bench_imv_new.err
python - <<'PY'
import json
for n in ('old','new'):
    try:
        d=json.load(open('gpurun_out/bench_imv_%s.json'%n))
        print(n, d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k:d['kernel_ms_per_step'][k] for k in ('imv_scan','aligned_pos','reconstruct','energy_gemm','expand_gemm')}, d['roofline_hbm_imv']['frac'], d['roofline_hbm_imv']['reconstruct']['frac'])
    except Exception as ex:
        print(n, 'failed', ex)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"reconstruct_alignment_rows|imv_scan_block|aligned_positions_block" -s 6 -c 3 -f -o gpurun_out/prof_imv_chain \
   python tools/quick_time.py C3 > gpurun_out/ncu_imv.log 2>&1
tail -2 gpurun_out/ncu_imv.log
timeout 1200 python tools/diag_precision.py C3 C5 fp64 > gpurun_out/precision.log 2>&1
grep -E "^C[235]" gpurun_out/precision.log | cut -c1-600
