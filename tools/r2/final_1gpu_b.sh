#!/bin/bash
# Round-2 closing evidence on one B200 after the set-up overlap and the multigrid coarse solver: GPU tests, the default bench line,
# the reference arm, the ncu launch list of the same command (kernels are those of final_1gpu.sh; their --set full captures stand).
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/final_gputests.txt 2>&1; tail -2 gpurun_out/r2/final_gputests.txt
timeout 600 python bench.py > gpurun_out/r2/final_bench.json 2> gpurun_out/r2/final_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/r2/final_reference.json 2> gpurun_out/r2/final_reference.err; echo "reference rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2/final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spd-packed --no-other-configs > gpurun_out/r2/final_ncu_launches.log 2>&1; echo "launch list rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
