#!/bin/bash
# One GPU session: the -m gpu suite, smoke(), and one bench line.  Usage (from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh [pytest-args...]'
# Everything lands under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 "$@" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
