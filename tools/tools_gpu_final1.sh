#!/bin/bash
mkdir -p gpurun_out
for S in 4 16 32; do
python bench.py --no-cpu-baseline --host-pipeline $S > gpurun_out/bench_p$S.json 2>gpurun_out/bench_p$S.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_p$S.json"))
print("pipeline $S: ms %.3f e2e_ms %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_apply_pipe_kernel --launch-skip 15 --launch-count 1 -o gpurun_out/apply_pipe_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
