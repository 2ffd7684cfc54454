#!/bin/bash
# weak-scaling line at 1, 2, 4, 8 GPUs with the final round-1 code (run with gpurun --gpus 8)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 8 --master-port 29561 tests/mgpu_check.py 4 3 1 1 2>&1 | grep -E "mgpu_check|rror" | tail -3
python bench.py --gpus 1 --no-cpu-baseline --steps 50 > gpurun_out/final_scale_1.json 2> gpurun_out/final_scale_1.err
for N in 2 4 8; do
timeout 900 $T --nproc-per-node $N --master-port 2957$N bench.py --gpus $N --steps 50 > gpurun_out/final_scale_$N.json 2> gpurun_out/final_scale_$N.err; echo "N=$N rc=$?"
done
python - <<PY
import json
base = None
for N in (1, 2, 4, 8):
    try:
        d = json.load(open("gpurun_out/final_scale_%d.json" % N))
        if N == 1: base = d["ms_per_step"]
        print(N, "value %.4g ms %.3f eff %.3f kernel_ms %.3f e2e_ms %.3f exch %s sym %.1e setup_wall %.2f" % (d["value"], d["ms_per_step"], base / d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["config"].get("exchange"), d["symmetry_check"]["rel_diff"], d["setup"]["setup_wall_s"]))
    except Exception as e:
        print(N, "failed", e)
PY
