#!/bin/bash
# ncu capture of the persistent packed-symmetric apply kernel (n = 125 class) at 40^3
mkdir -p gpurun_out
cat > /tmp/run_sym.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from ssc_b200 import PC, DeviceVec, fem, plex, synth
from ssc_b200.patch import setup_patch_pc
mesh = plex.TensorPlex((40, 40, 40)); V = fem.FunctionSpace(mesh, 3); disc = fem.Discretisation([V])
pc = PC(0); setup_patch_pc(pc, disc, None); pc._compute_op = None
pc.setOperatorCSR(*synth.tensor_poisson_csr(V, disc.ghost_bc_nodes))
pc.setOption("sub_pc_type", "cholesky"); pc.setOption("ssc_symmetric_storage", True)
pc.setOption("ssc_symmetric_persistent", sys.argv[1]); pc.setOption("ssc_symmetric_bulk_copy", sys.argv[2])
pc.setUp()
x = DeviceVec(pc, disc.total_dofs); y = DeviceVec(pc, disc.total_dofs); x.set(np.random.default_rng(0).uniform(-1, 1, disc.total_dofs))
for _ in range(4): pc.apply(x, y)
pc.synchronize(); print("ok")
PY
python /tmp/run_sym.py true true > gpurun_out/plain_sym.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_apply_sym_persistent --launch-skip 2 --launch-count 1 -o gpurun_out/sym_persistent python /tmp/run_sym.py true true > gpurun_out/ncu_sym.log 2>&1
echo "rc=$?"
python /tmp/run_sym.py false true > gpurun_out/plain_sym2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssc_apply_sym2_kernel --launch-skip 11 --launch-count 1 -o gpurun_out/sym_simple python /tmp/run_sym.py false true > gpurun_out/ncu_sym2.log 2>&1
echo "rc=$?"
