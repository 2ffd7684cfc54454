set -x
mkdir -p gpurun_out/r2
python tools/r2/err_probe.py > gpurun_out/r2/err_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gputests0.txt 2>&1
tail -5 gpurun_out/r2/gputests0.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench0.json 2> gpurun_out/r2/bench0.err
tail -c 3000 gpurun_out/r2/bench0.json
