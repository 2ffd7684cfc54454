#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29511 tests/mgpu_check.py 4 3 0 1 2>&1 | grep -E "mgpu_check|rror" | tail -5
timeout 300 $T --master-port 29512 tests/mgpu_check.py 6 2 1 1 2>&1 | grep -E "mgpu_check|rror" | tail -5
timeout 300 $T --master-port 29513 tests/mgpu_check.py 5 3 1 0 2>&1 | grep -E "mgpu_check|rror" | tail -5
timeout 600 $T --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 --exchange nccl > gpurun_out/bench2_nccl.json 2> gpurun_out/bench2_nccl.err; tail -2 gpurun_out/bench2_nccl.err
timeout 600 $T --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 3 --exchange p2p > gpurun_out/bench2_p2p.json 2> gpurun_out/bench2_p2p.err; tail -2 gpurun_out/bench2_p2p.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench1.json 2> gpurun_out/bench1.err
python - <<PY
import json
for f in ("bench1", "bench2_nccl", "bench2_p2p"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.4g ms %.3f kernel_ms %.3f e2e_ms %.3f exchange %s launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["config"].get("exchange"), d["gpu_launches"]))
    except Exception as e:
        print(f, "failed", e)
PY
