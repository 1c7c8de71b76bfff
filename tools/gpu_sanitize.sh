#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
  -k "forward_matches_oracle or batched_inference or length_regulator_bit_exact or helper_methods or duration_predictor or inference_matches_golden or without_loss_masking" > gpurun_out/memcheck_tests.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck_tests.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/synccheck.log 2>&1
echo "synccheck exit $?" >> gpurun_out/synccheck.log
grep -E "passed|failed|ERROR SUMMARY|exit" gpurun_out/memcheck_tests.log | tail -5; grep -E "ERROR SUMMARY|exit|smoke" gpurun_out/synccheck.log | tail -5
