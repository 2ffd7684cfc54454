#!/bin/bash
# ncu --set full of the extract kernel (n = 125 launch) at 32^3 Q3 from the assembled CSR
mkdir -p gpurun_out/r2
cat > /tmp/run_ex.py <<'PY'
import sys, os, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, fem, plex, synth
from ssc_b200.patch import setup_patch_pc
N = int(os.environ.get("CELLS", "32"))
mesh = plex.TensorPlex((N, N, N)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
csr = synth.tensor_poisson_csr(V, disc.ghost_bc_nodes)
pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
pc.setOperatorCSR(*csr)
pc.setUp()
print("ok extract ms", pc.eventTime("PCPATCHComputeOp"), "factor ms", pc.eventTime("PCPATCHSolve"))
PY
python /tmp/run_ex.py > gpurun_out/r2/plain_ex.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_extract --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2/extract python /tmp/run_ex.py > gpurun_out/r2/ncu_ex.log 2>&1
echo "rc=$?"; cat gpurun_out/r2/plain_ex.log; tail -3 gpurun_out/r2/ncu_ex.log
