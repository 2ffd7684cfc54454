#!/bin/bash
# end-of-round capture: default bench line, reference arm, ncu launch list of the same (shortened) command
mkdir -p gpurun_out
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "reference rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-spd-packed > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-spd-packed > gpurun_out/ncu1.log 2>&1
echo "ncu rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("ms %.3f value %.4g frac %.3f e2e_ms %.3f setup %.2f factor %.1f extract %.1f lists %.1f cpu %.4g parity %s" % (d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["setup"]["setup_wall_s"], d["setup"]["factor_ms"], d["setup"]["extract_ms"], d["setup"]["lists_ms"], d["cpu_baseline"]["value"], d["cpu_baseline"]["parity"]))
r = json.loads(open("gpurun_out/final_reference.json").read().strip().splitlines()[-1])
print("reference arm: value %.4g %s cores %s" % (r["value"], r["unit"], r["cpu_baseline"]["cores"]))
PY
