#!/bin/bash
# Q4 (n = 343 interior) weak scaling over the GPUs of one box from element matrices: bash tools/gpu_q4_8gpu.sh NGPU CELLS
N=${1:-8}; C=${2:-40}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --degree 4 --cells $C --operator elements > gpurun_out/q4_${C}_1gpu.json 2> gpurun_out/q4_${C}_1gpu.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --degree 4 --cells $C --operator elements > gpurun_out/q4_${C}_${N}gpu.json 2> gpurun_out/q4_${C}_${N}gpu.err
fi
echo "rc=$?"; tail -c 1500 gpurun_out/q4_${C}_${N}gpu.json; tail -5 gpurun_out/q4_${C}_${N}gpu.err
