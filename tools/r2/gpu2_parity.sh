mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2/gpu_multi_2.txt 2>&1; tail -3 gpurun_out/r2/gpu_multi_2.txt
timeout 900 python bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/bench_2gpu.json 2> gpurun_out/r2/bench_2gpu.err; echo rc=$?
tail -c 1500 gpurun_out/r2/bench_2gpu.json; tail -5 gpurun_out/r2/bench_2gpu.err
