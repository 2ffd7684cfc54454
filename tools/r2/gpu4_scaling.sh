mkdir -p gpurun_out/r2
N=${1:-4}
timeout 600 python bench.py --gpus $N --cells 16 --steps 5 --warmup 3 --partition brick --scaling strong > gpurun_out/r2/bench_${N}gpu_brick16.json 2> gpurun_out/r2/bench_${N}gpu_brick16.err; echo "brick16 rc=$?"; tail -2 gpurun_out/r2/bench_${N}gpu_brick16.err
for mode in "slab strong" "brick strong" "slab weak" "brick weak"; do
  set -- $mode
  timeout 900 python bench.py --gpus $N --steps 20 --warmup 3 --partition $1 --scaling $2 > gpurun_out/r2/bench_${N}gpu_$1_$2.json 2> gpurun_out/r2/bench_${N}gpu_$1_$2.err; echo "$1 $2 rc=$?"
  tail -2 gpurun_out/r2/bench_${N}gpu_$1_$2.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/bench_${N}gpu_$1_$2.json"))
    print("$1 $2", "ms", round(d["ms_per_step"],3), "Gdof/s", round(d["value"]/1e9,3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "pageable", round(d["e2e_pageable"]["ms_per_step"],3), "exch", round(d.get("exchange_ms_per_apply",0),3), "frac", round(d["roofline"]["frac"],3), "parity", d["parity"]["ok"], d["parity"]["max_rel_err_vs_oracle"], d["parity"]["dofs_checked"])
except Exception as e: print("ERR", e)
PY
done
