#!/bin/bash
# Round-2 evidence on one B200: GPU tests, the default bench line, the reference arm, the ncu launch list of the same command,
# and ncu --set full captures of the apply, extract and factor kernels at the benchmark size.
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/final_gputests.txt 2>&1; tail -2 gpurun_out/r2/final_gputests.txt
python bench.py > gpurun_out/r2/final_bench.json 2> gpurun_out/r2/final_bench.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r2/final_reference.json 2> gpurun_out/r2/final_reference.err; echo "reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2/final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spd-packed --no-other-configs > gpurun_out/r2/final_ncu_launches.log 2>&1; echo "launch list rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spd-packed --no-other-configs --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:ssc_apply_pipe --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2/final_apply $B > gpurun_out/r2/final_ncu_apply.log 2>&1; echo "apply rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ssc_invert_l2 --launch-skip 0 --launch-count 1 -f -o gpurun_out/r2/final_factor $B > gpurun_out/r2/final_ncu_factor.log 2>&1; echo "factor rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ssc_extract --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2/final_extract $B > gpurun_out/r2/final_ncu_extract.log 2>&1; echo "extract rc=$?"
ls -la gpurun_out/r2/*.ncu-rep
