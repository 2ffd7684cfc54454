#!/bin/bash
# ncu capture of the look-ahead factor kernel (n = 125 class) at 24^3 Q3
mkdir -p gpurun_out
cat > /tmp/run_la.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex
from ssc_b200.patch import setup_patch_pc
mesh = plex.TensorPlex((24, 24, 24)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
pc.setElementMatrices(fem.element_matrix_tensor(V))
pc.setOption("ssc_factor_variant", 4)
pc.setUp()
print("ok factor ms", pc.eventTime("PCPATCHSolve"))
PY
python /tmp/run_la.py > gpurun_out/plain_la.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_invert_panel2 --launch-skip 3 --launch-count 1 -o gpurun_out/invert_panel python /tmp/run_la.py > gpurun_out/ncu_la.log 2>&1
echo "rc=$?"
