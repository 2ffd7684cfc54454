#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 4 --master-port 29521 tests/mgpu_check.py 4 3 1 1 2>&1 | grep -E "mgpu_check|rror" | tail -5
timeout 300 $T --nproc-per-node 8 --master-port 29522 tests/mgpu_check.py 4 2 0 1 2>&1 | grep -E "mgpu_check|rror" | tail -5
timeout 300 $T --nproc-per-node 4 --master-port 29523 tests/mgpu_check.py 4 2 0 0 2>&1 | grep -E "mgpu_check|rror" | tail -5
for N in 8 4; do
timeout 900 $T --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; echo "N=$N rc=$?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/scale_$N.err | tail -3
done
python - <<PY
import json
for N in (4, 8):
    try:
        d = json.load(open("gpurun_out/scale_%d.json" % N))
        print(N, "value %.4g ms %.3f kernel_ms %.3f e2e_ms %.3f exchange %s setup %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["config"].get("exchange"), d["setup"]))
    except Exception as e:
        print(N, "failed", e)
PY
