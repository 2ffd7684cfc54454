mkdir -p gpurun_out/r2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/mgpu_pcie.py bind > gpurun_out/r2/mgpu_pcie_8.txt 2>&1; grep "GB/s" gpurun_out/r2/mgpu_pcie_8.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29612 tools/mgpu_pcie.py bind > gpurun_out/r2/mgpu_pcie_1.txt 2>&1; grep "GB/s" gpurun_out/r2/mgpu_pcie_1.txt
SSC_PIPE_TRACE=1 timeout 600 python bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2/final_bench_8gpu.json 2> gpurun_out/r2/final_bench_8gpu.err; echo rc=$?
grep "ssc pipe rank" gpurun_out/r2/final_bench_8gpu.err | tail -8 > gpurun_out/r2/pipe_trace_8gpu_final.txt; cat gpurun_out/r2/pipe_trace_8gpu_final.txt
python -c "
import json; d=json.load(open('gpurun_out/r2/final_bench_8gpu.json'))
print('ms', d['ms_per_step'], 'Gdof/s', d['value']/1e9, 'e2e', d['e2e']['ms_per_step'], 'pageable', d['e2e_pageable']['ms_per_step'], 'parity', d['parity']['ok'], d['parity']['dofs_checked'])
"
