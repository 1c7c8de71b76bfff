#!/bin/bash
# Last validation of a round: full GPU suite, smoke, one bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print(d["value"], d["ms_per_step"], d["timing"]["reported"], d["clocks"]["sm_mhz"], d["rtf_batch1"]["ms"],
      d["vocoder"]["samples_per_s"], d["vocoder"]["text_to_wave_c1"]["ms"],
      d["cpu_baseline"].get("one_thread", {}).get("value"), d["cpu_baseline"].get("c1_inference"))
PY
