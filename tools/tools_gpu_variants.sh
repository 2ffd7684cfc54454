#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0 1 2; do
  python bench.py --apply-variant $v --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "variant $v rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_v$v.json"))
print("variant $v value %.4g ms %.3f frac %.3f kernel_ms %.3f e2e_ms %.3f setup %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["setup"]))
PY
done
