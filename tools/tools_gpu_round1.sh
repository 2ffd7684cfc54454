#!/bin/bash
# first measurement pass on a B200: sanity, full bench, ncu launch list + full capture of the apply kernel
set -x
mkdir -p gpurun_out
python bench.py --cells 16 --steps 3 --warmup 3 --cpu-sample 512 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_apply_kernel --launch-skip 15 --launch-count 1 -o gpurun_out/apply_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
ls -la gpurun_out
