#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-spd-packed > gpurun_out/plain_ex.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_extract_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/extract_n125 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-spd-packed > gpurun_out/ncu_ex.log 2>&1
echo "rc=$?"
