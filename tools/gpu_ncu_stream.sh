#!/bin/bash
# ncu capture of the persistent small-patch apply kernel (3D Q2 96^3, n = 27)
mkdir -p gpurun_out
cat > /tmp/run_stream.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, DeviceVec, fem, plex
from ssc_b200.patch import setup_patch_pc
mesh = plex.TensorPlex((96, 96, 96)); V = fem.FunctionSpace(mesh, 2); disc = fem.Discretisation([V])
pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
pc.setElementMatrices(fem.element_matrix_tensor(V))
pc.setUp()
x = DeviceVec(pc, disc.total_dofs); y = DeviceVec(pc, disc.total_dofs); x.set(np.random.default_rng(0).uniform(-1, 1, disc.total_dofs))
for _ in range(3): pc.apply(x, y)
pc.synchronize(); print("ok", pc.size("bytes_apply"))
PY
python /tmp/run_stream.py > gpurun_out/plain_stream.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_apply_stream_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/apply_stream python /tmp/run_stream.py > gpurun_out/ncu_stream.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_stream.log
