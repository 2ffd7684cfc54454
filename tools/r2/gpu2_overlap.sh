mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
for ov in true false; do
python - <<PY
import subprocess, json, sys
out = subprocess.run([sys.executable, "bench.py", "--gpus", "2", "--steps", "30", "--warmup", "3", "--scaling", "strong", "--no-e2e"] , capture_output=True, text=True, env=dict(__import__("os").environ, SSC_OVERLAP_EXCHANGE="$ov"))
try:
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    print("overlap_exchange=$ov strong 2 GPUs: ms %.4f exch %.4f parity %s" % (d["ms_per_step"], d["exchange_ms_per_apply"], d["parity"]["ok"]))
except Exception as e:
    print("ERR", e, out.stderr[-1500:])
PY
done
