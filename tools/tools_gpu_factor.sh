#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --no-cpu-baseline > gpurun_out/bench_f1.json 2> gpurun_out/bench_f1.err; echo "rc=$?"; tail -3 gpurun_out/bench_f1.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_f1.json"))
print("value %.4g ms %.3f frac %.3f kernel_ms %.3f e2e_ms %.3f setup %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["setup"]))
PY
