#!/bin/bash
# e2e (host vectors) and device-resident numbers at N = 8 and 2 after the multi-rank host pipeline (gpurun --gpus 8)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 2; do
timeout 600 $T --nproc-per-node $N --master-port 2958$N bench.py --gpus $N --steps 30 > gpurun_out/pipe_scale_$N.json 2> gpurun_out/pipe_scale_$N.err; echo "N=$N rc=$?"
done
python - <<PY
import json
for N in (2, 8):
    try:
        d = json.loads(open("gpurun_out/pipe_scale_%d.json" % N).read().strip().splitlines()[-1])
        print(N, "value %.4g ms %.3f kernel_ms %.3f e2e_ms %.3f e2e %.4g exch %s sym %.1e setup_wall %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["config"].get("exchange"), d["symmetry_check"]["rel_diff"], d["setup"]["setup_wall_s"]))
    except Exception as e:
        print(N, "failed", e)
PY
