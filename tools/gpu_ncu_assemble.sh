#!/bin/bash
# ncu capture of the element-matrix assemble kernel (n = 125 class at 40^3 Q3)
mkdir -p gpurun_out
cat > /tmp/run_asm.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc
mesh = plex.TensorPlex((40, 40, 40)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
pc.setElementMatrices(fem.element_matrix_tensor(V))
pc.setUp()
print("ok assemble ms", pc.eventTime("PCPATCHComputeOp"))
PY
python /tmp/run_asm.py > gpurun_out/plain_asm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_assemble_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/assemble python /tmp/run_asm.py > gpurun_out/ncu_asm.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_asm.log
