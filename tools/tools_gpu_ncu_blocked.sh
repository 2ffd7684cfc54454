#!/bin/bash
mkdir -p gpurun_out
python bench.py --degree 4 --cells 12 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_invert_blocked --launch-skip 1 --launch-count 1 -o gpurun_out/invert_blocked python bench.py --degree 4 --cells 12 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu4.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu4.log | cut -c1-200
