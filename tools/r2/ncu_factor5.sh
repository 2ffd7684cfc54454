#!/bin/bash
# ncu --set full of the register-resident DMMA factor kernel (n = 125 launch) at 24^3 Q3
mkdir -p gpurun_out/r2
cat > /tmp/run_f5.py <<'PY'
import sys, os, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc
mesh = plex.TensorPlex((24, 24, 24)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
pc.setElementMatrices(fem.element_matrix_tensor(V))
pc.setOption("ssc_factor_variant", int(os.environ.get("FVAR", "5")))
pc.setOption("sub_pc_type", os.environ.get("SUB", "lu"))
pc.setUp()
print("ok factor ms", pc.eventTime("PCPATCHSolve"))
PY
python /tmp/run_f5.py > gpurun_out/r2/plain_f5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_invert_dmma --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2/invert_dmma python /tmp/run_f5.py > gpurun_out/r2/ncu_f5.log 2>&1
echo "rc=$?"; cat gpurun_out/r2/plain_f5.log; tail -3 gpurun_out/r2/ncu_f5.log
