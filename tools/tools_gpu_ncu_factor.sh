#!/bin/bash
mkdir -p gpurun_out
python bench.py --cells 24 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_invert_reg --launch-skip 3 --launch-count 1 -o gpurun_out/invert_reg python bench.py --cells 24 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu3.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu3.log
