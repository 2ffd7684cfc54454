// sm_100a kernels of the patch preconditioner, by theme.  Layouts are described in kernels_common.cuh.
#pragma once
#include "kernels_common.cuh"
#include "kernels_operators.cuh"   // K1 extract, K1e assemble
#include "kernels_factor.cuh"      // K2 Gauss-Jordan family
#include "kernels_factor_l2.cuh"   // K2, 33 <= n <= 128: blocked Gauss-Jordan on the fp64 tensor cores, block in L2, four patches in flight per SM
#include "kernels_apply.cuh"       // K3+K4+K5 fused apply
#include "kernels_vector.cuh"      // BC rows, weights, slot sum, SpMV, calibration read
#include "kernels_peer.cuh"        // NVLink peer exchange
#include "kernels_packed.cuh"      // packed-symmetric storage
