mkdir -p gpurun_out/r2
N=8
run() {  # name, extra args, env
  name=$1; shift
  timeout 900 python bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/r2/bench_${N}gpu_$name.json 2> gpurun_out/r2/bench_${N}gpu_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/bench_${N}gpu_$name.json"))
    e=d.get("e2e") or {}; p=d.get("e2e_pageable") or {}
    print("$name", "ms", round(d["ms_per_step"],3), "Gdof/s", round(d["value"]/1e9,3), "e2e ms", round(e.get("ms_per_step",0),3), "pageable", round(p.get("ms_per_step",0),3), "exch", round(d.get("exchange_ms_per_apply",0),3), "frac", round(d["roofline"]["frac"],3), "parity", d["parity"]["ok"], d["parity"]["max_rel_err_vs_oracle"], d["parity"]["dofs_checked"], "setup", round(d["setup"]["setup_wall_s"],2))
except Exception as ex: print("ERR", ex)
PY
  grep "ssc pipe rank" gpurun_out/r2/bench_${N}gpu_$name.err | tail -16 > gpurun_out/r2/pipe_trace_${N}gpu_$name.txt
}
SSC_PIPE_TRACE=1 run slab_weak
run slab_strong --scaling strong --no-e2e
run brick_strong --scaling strong --partition brick --no-e2e
run brick_weak --partition brick --no-e2e
timeout 300 python tools/mgpu_pcie.py > gpurun_out/r2/mgpu_pcie_8.txt 2>&1; tail -12 gpurun_out/r2/mgpu_pcie_8.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
