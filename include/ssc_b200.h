/*
 * ssc_b200.h -- C ABI of the B200-native patch subspace-correction preconditioner.
 *
 * This is the boundary a maintainer of wence-/ssc binds instead of libssc.c's per-patch loops
 * (INTEGRATION.md shows the Cython stub).  Plain C: opaque handle, pointers and sizes, int return
 * (0 = success, otherwise a PETSc-style error code; SSCGetLastError() gives the text), mirroring
 * `PetscErrorCode` + CHKERR of ssc/PatchPC.pyx:36-47.  No PETSc, torch or CUDA types appear here.
 *
 * SSCInt is PetscInt as the reference builds it (`ctypedef long PetscInt`, ssc/PatchPC.pyx:7);
 * scalars are fp64 (`ctypedef double PetscScalar`, ssc/PatchPC.pyx:8).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference root).
 */
#ifndef SSC_B200_H
#define SSC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t SSCInt;
typedef struct SSCPatchPC_s *SSCPatchPC;

/* memory space of the vector / matrix pointers passed to SetOperator / Apply */
enum { SSC_MEM_HOST = 0, SSC_MEM_DEVICE = 1 };
/* patch construction, PCPatchConstructType ssc/libssc.h:11 */
enum { SSC_PATCH_STAR = 0, SSC_PATCH_VANKA = 1, SSC_PATCH_USER = 2, SSC_PATCH_PYTHON = 3 };
/* patch sub-solver: -sub_pc_type under -sub_ksp_type preonly (ssc/libssc.c:1240-1241, :1295) */
enum { SSC_SUB_LU = 0, SSC_SUB_CHOLESKY = 1 };

/* ---- library / device -------------------------------------------------------------------- */
/* text of the last error raised on this thread (SETERRQ messages of ssc/libssc.c) */
const char *SSCGetLastError(void);
/* number of CUDA devices visible; 0 is an error for every compute entry point (there is no CPU fallback) */
int SSCDeviceCount(int *count);
/* PCPatchInitializePackage ssc/libssc.c:11-27: registers the six log-event names used by SSCPatchGetEventTime */
int SSCPatchInitializePackage(void);

/* ---- lifecycle: PC vtable slots of ssc/libssc.c:1714-1722 -------------------------------- */
/* device = CUDA ordinal of this rank's GPU.  device = -1 creates a host-only handle on which only the patch
 * construction entry points work (SetPlex .. BuildPatches, GetPatches/GetCells/GetDofs); SetUp/Apply on it fail. */
int SSCPatchCreate(int device, SSCPatchPC *pc);           /* PCCreate_PATCH   ssc/libssc.c:1687-1725 (defaults :1697-1711) */
int SSCPatchReset(SSCPatchPC pc);                         /* PCReset_PATCH    ssc/libssc.c:902-985  */
int SSCPatchDestroy(SSCPatchPC *pc);                      /* PCDestroy_PATCH  ssc/libssc.c:987-1004 */

/* ---- options: PCSetFromOptions_PATCH ssc/libssc.c:1551-1620 ------------------------------- */
/* name is the option without the prefix and leading dash, e.g. "pc_patch_save_operators", "sub_pc_type";
 * value is its string form ("true", "1", "lu", ...).  Unknown names and unsupported values return an error
 * (no silent fallback).  "pc_patch_multiplicative true": the reference raises an error (mode removed, ssc/libssc.c:1570-1572);
 * this build runs a coloured multiplicative sweep instead (BASELINE.json north_star).
 * Names of the reference (ssc/libssc.c:1551-1620): pc_patch_save_operators, pc_patch_partition_of_unity,
 * pc_patch_multiplicative, pc_patch_construction_dim | _codim | _type, pc_patch_vanka_dim, pc_patch_exclude_subspace,
 * pc_patch_sub_mat_type, pc_patch_print_patches, pc_patch_symmetrise_sweep, sub_ksp_type, sub_pc_type.
 * Extensions of this build all start with "ssc_" (kernel selection, launch shapes, storage modes, host pipeline);
 * they are listed with their defaults in DESIGN.md section 2 and never change results beyond rounding. */
int SSCPatchSetOption(SSCPatchPC pc, const char *name, const char *value);
int SSCPatchSetSaveOperators(SSCPatchPC pc, int flg);     /* PCPatchSetSaveOperators     ssc/libssc.c:84-91   */
int SSCPatchSetPartitionOfUnity(SSCPatchPC pc, int flg);  /* PCPatchSetPartitionOfUnity  ssc/libssc.c:93-100  */
int SSCPatchSetSubMatType(SSCPatchPC pc, const char *t);  /* PCPatchSetSubMatType        ssc/libssc.c:281-291 */

/* ---- topology + discretisation (host pointers) ------------------------------------------- */
/* Replaces PCSetDM(mesh._plex) (ssc/patch.py:214) and the DMPlex queries of ssc/libssc.c:472-504:
 * chart [0,pEnd), cone table, depth strata [depthStart[d], depthEnd[d]) for d = 0..dim, and the
 * "pyop2_ghost" label as one byte per point (may be NULL = nothing is ghost).  Arrays are copied. */
int SSCPatchSetPlex(SSCPatchPC pc, SSCInt dim, SSCInt pEnd, const SSCInt *coneOff, const SSCInt *cones,
                    const SSCInt *depthStart, const SSCInt *depthEnd, const unsigned char *ghost);
/* PCPatchSetCellNumbering ssc/libssc.c:225-234: the section is passed as its (dof, offset) arrays over the chart. Copied. */
int SSCPatchSetCellNumbering(SSCPatchPC pc, const SSCInt *dof, const SSCInt *off);
/* PCPatchSetDiscretisationInfo ssc/libssc.c:237-279 (declared ssc/libssc.h:7).  `DM *dms` becomes the default
 * section of each DM as (secDof[k], secOff[k]) arrays over the chart.  Ownership follows the reference:
 * cellNodeMap[k] is BORROWED for the lifetime of the PC (:268); everything else is copied (:265-277). */
int SSCPatchSetDiscretisationInfo(SSCPatchPC pc, SSCInt nsubspaces, const SSCInt *const *secDof,
                                  const SSCInt *const *secOff, const SSCInt *bs, const SSCInt *nodesPerCell,
                                  const SSCInt *const *cellNodeMap, const SSCInt *subspaceOffsets,
                                  SSCInt numGhostBcs, const SSCInt *ghostBcNodes,
                                  SSCInt numGlobalBcs, const SSCInt *globalBcNodes);
/* Result of PCPatchSetUserPatchConstructionOperator's callback (ssc/libssc.c:310, :487; ssc/PatchPC.pyx:74-98):
 * the user's index sets flattened as offsets/points, and the iteration set.  Copied. */
int SSCPatchSetUserPatches(SSCPatchPC pc, SSCInt nuserIS, const SSCInt *userOff, const SSCInt *userPoints,
                           SSCInt niter, const SSCInt *iterationSet);
/* Alternative to the three calls above for a caller that already holds the dof lists (e.g. a live PETSc
 * PCPATCH): gtolCounts/gtol of ssc/libssc.c:779-837 as offsets[npatch+1] + gtol[offsets[npatch]], the size of
 * the local (owned+ghost) vector, the size of the owned vector, and globalBcNodes (:277).  Copied. */
int SSCPatchSetPatches(SSCPatchPC pc, SSCInt npatch, const SSCInt *offsets, const SSCInt *gtol,
                       SSCInt nlocal, SSCInt nowned, SSCInt numGlobalBcs, const SSCInt *globalBcNodes);

/* ---- star forest (owner <-> ghost map), replaces PCPatchCreateDefaultSF_Private ssc/libssc.c:102-223 -- */
/* nowned = roots on this rank (size of x, y); the local vector has nlocal = subspaceOffsets[nsubspaces] leaves.
 * selfLeaf[i] (i < nself) is a local slot whose root is selfRoot[i] on this rank.  For neighbour k of nneigh:
 * sendRoot[sendOff[k]..sendOff[k+1]) are owned indices whose values fill ghosts on rank neigh[k] (SFBcast, :1323)
 * and that receive-and-add the neighbour's contributions (SFReduce, :1399); recvLeaf[recvOff[k]..) are the local
 * ghost slots owned by rank neigh[k], in the order that rank lists them in its sendRoot.  Copied.
 * Without this call the SF is the identity on [0, nowned) (one process). */
int SSCPatchSetStarForest(SSCPatchPC pc, SSCInt nowned, SSCInt nself, const SSCInt *selfLeaf, const SSCInt *selfRoot,
                          SSCInt nneigh, const int *neigh, const SSCInt *sendOff, const SSCInt *sendRoot,
                          const SSCInt *recvOff, const SSCInt *recvLeaf);
/* NCCL communicator over the ranks of one NVSwitch box (the MPI communicator of the reference's PC).
 * uniqueId is the 128-byte ncclUniqueId produced by SSCCommGetUniqueId on rank 0 and passed around by the host. */
int SSCCommGetUniqueId(char uniqueId[128]);
int SSCPatchCommInit(SSCPatchPC pc, int nranks, int rank, const char uniqueId[128]);

/* NVLink peer exchange, an alternative to the NCCL path for ranks on one NVSwitch box.  Export gives 192 bytes (three
 * CUDA IPC handles: this rank's ghost-value array of x, its "incoming" array with one entry per owned dof, and its
 * barrier flags) to be handed to every neighbour.  Connect takes the neighbours' exports in the order of the star
 * forest's neigh[], this rank's position in each neighbour's neigh[] (slotAtNeigh), and for every entry of sendRoot /
 * recvLeaf the matching index on the neighbour: its ghost slot number (local leaf index minus its owned count) /
 * its owned index.  After Connect the halo fill is direct peer stores, the overlap sum is issued by the patch kernel
 * itself as fp64 reductions into the owners' incoming arrays, and x / y are used in place (no local copies);
 * SSCPatchCommInit is then not needed. */
int SSCPatchPeerExport(SSCPatchPC pc, char *handles192);
int SSCPatchPeerConnect(SSCPatchPC pc, const char *neighHandles, const int *slotAtNeigh, const SSCInt *remoteLeaf,
                        const SSCInt *remoteRoot);

/* ---- operator: replaces PCPatchSetComputeOperator ssc/libssc.c:293-308 (declared ssc/libssc.h:8) ------ */
/* The reference re-assembles every patch matrix through a callback (:1150).  Here the patch blocks are cut out
 * of the assembled local operator in CSR form: rows for every local dof that lies in a patch, columns in local
 * numbering.  memspace says where the three arrays live.  The arrays are BORROWED until the next SetUp returns. */
int SSCPatchSetOperatorCSR(SSCPatchPC pc, SSCInt nrows, const SSCInt *rowptr, const int32_t *colidx,
                           const double *vals, int memspace);
/* New values on the nonzero pattern of the previous SSCPatchSetOperatorCSR (the matrix was re-assembled with
 * SAME_NONZERO_PATTERN, as before every PatchPC.update of a nonlinear solve, ssc/patch.py:269-270): the next SetUp moves
 * only `vals` to the device (two thirds of the bytes of the operator).  Same memspace as the pattern; the rowptr / colidx
 * arrays of the earlier call are not read again (the device copy of the pattern is kept). */
int SSCPatchSetOperatorValues(SSCPatchPC pc, const double *vals);

/* The reference's own source of the patch operators: element matrices added cell by cell through the per-cell
 * dof table with negative (BC) indices dropped -- PCPatchSetComputeOperator / PCPatchComputeOperator
 * (ssc/libssc.h:8, ssc/libssc.c:1120-1156), the callback trampoline ssc/PatchPC.pyx:54-72 and the element kernel
 * call ssc/patch.py:28-62, 210-213.  Ke holds one dofs_per_cell x dofs_per_cell row-major matrix per cell, rows of
 * Ke indexed by the cell numbering (SSCPatchSetCellNumbering), cell dofs ordered subspace by subspace, node by node,
 * component by component (the order of the dofs table, ssc/libssc.c:846-889); cell_stride is the distance in
 * doubles between two cells' matrices, 0 = one matrix shared by all cells (congruent cells, constant coefficients).
 * Replaces a previous SSCPatchSetOperatorCSR (and vice versa).  Borrowed until SSCPatchSetUp returns (until
 * SSCPatchReset when save_operators is off).  The blocks are summed in the reference's order, cell after cell, so
 * they are bit-identical to the oracle's element-wise assembly.  Needs patches built from the mesh (not
 * SSCPatchSetPatches); multiplicative sweeps need the assembled operator instead. */
int SSCPatchSetElementMatrices(SSCPatchPC pc, SSCInt ncells, SSCInt dofs_per_cell, const double *Ke, SSCInt cell_stride, int memspace);

/* The arrays handed to SSCPatchSetOperatorCSR / SSCPatchSetElementMatrices are borrowed, like the data behind the
 * reference's compute-operator callback (ssc/PatchPC.pyx:158-163 keeps the Python objects alive, nothing is copied).
 * This call tells the library they are gone: device copies of the last SSCPatchSetUp stay usable, a further
 * SSCPatchSetUp fails with a wrong-state error until an operator is set again. */
int SSCPatchReleaseOperator(SSCPatchPC pc);

/* ---- setup and apply ----------------------------------------------------------------------- */
int SSCPatchSetUp(SSCPatchPC pc);                         /* PCSetUp_PATCH ssc/libssc.c:1201-1299 (+ the lazy factorisation inside :1374) */
/* PCApply_PATCH ssc/libssc.c:1301-1426: y = sum_p R_p^T B_p^{-1} R_p x on free dofs, y = x on owned Dirichlet dofs,
 * optionally weighted.  x and y have nowned entries and live in `memspace`. Stream-ordered; returns after y is complete
 * for SSC_MEM_HOST, after enqueue for SSC_MEM_DEVICE (use SSCPatchSynchronize). */
int SSCPatchApply(SSCPatchPC pc, const double *x, double *y, int memspace);
/* slot pc->ops->applytranspose is 0 in the reference (ssc/libssc.c:1715): always an error (PETSC_ERR_SUP) */
int SSCPatchApplyTranspose(SSCPatchPC pc, const double *x, double *y, int memspace);
int SSCPatchSynchronize(SSCPatchPC pc);
/* PCView_PATCH ssc/libssc.c:1622-1685: writes the same lines into buf */
int SSCPatchView(SSCPatchPC pc, char *buf, SSCInt len);

/* ---- introspection -------------------------------------------------------------------------- */
/* sizes: "npatch", "nlocal", "nowned", "sum_n", "sum_n2", "max_n", "num_cells", "num_classes", "block_bytes",
 * "bytes_apply", "bytes_extract", "flops_factor", "failed_patches", "gpu_launches" */
int SSCPatchGetSize(SSCPatchPC pc, const char *what, SSCInt *value);
/* run the host-side patch construction now (ssc/libssc.c:1219-1220); needs no device.  SetUp calls it if needed. */
int SSCPatchBuildPatches(SSCPatchPC pc);
/* -pc_patch_print_patches (ssc/libssc.c:674-730): owned / seen / global-BC / artificial-BC dofs of every patch, ascending;
 * printed to stdout when the patches are built and available here as text (needed = length incl. the final 0) */
int SSCPatchGetPrintout(SSCPatchPC pc, char *buf, SSCInt len, SSCInt *needed);
/* copies of the host-side construction results (ssc/libssc.c:452-900); any pointer may be NULL */
int SSCPatchGetPatches(SSCPatchPC pc, SSCInt *gtolOffsets, SSCInt *gtol);
int SSCPatchGetCells(SSCPatchPC pc, SSCInt *cellOffsets, SSCInt *cells);
int SSCPatchGetDofs(SSCPatchPC pc, SSCInt *dofs);         /* only if option "ssc_keep_dofs_table" is on */
/* dense operator (which = 0) or stored inverse (which = 1) of patch p, row-major n x n, copied to host */
int SSCPatchGetPatchMatrix(SSCPatchPC pc, SSCInt p, int which, double *out);
/* colour of every patch (-1 for empty ones) after SetUp with -pc_patch_multiplicative (N4: no reference counterpart) */
int SSCPatchGetColours(SSCPatchPC pc, SSCInt *colours);
/* patches whose operator could not be factorised at the last SetUp (the reference would surface
 * KSP_DIVERGED_PCSETUP_FAILED per sub-KSP, dead code at ssc/libssc.c:1437-1443) */
int SSCPatchGetFailedPatches(SSCPatchPC pc, SSCInt *list, SSCInt maxn, SSCInt *nfailed);
/* dof weights of ssc/libssc.c:1254-1284 (nowned values, host) */
int SSCPatchGetDofWeights(SSCPatchPC pc, double *w);
/* device time in ms of the last completed stage, named as the reference's log events (ssc/libssc.c:19-24):
 * "PCPATCHCreate", "PCPATCHComputeOp" (extract), "PCPATCHSolve" (factor at setup / fused apply kernel),
 * "PCPATCHApply", "PCPATCHScatter", "PCPATCHPrealloc". */
int SSCPatchGetEventTime(SSCPatchPC pc, const char *event, double *ms);
/* accumulate device time of the fused apply kernel(s) over applies (CUDA events on the apply stream) */
int SSCPatchKernelTimerReset(SSCPatchPC pc, int enable);
int SSCPatchKernelTimerRead(SSCPatchPC pc, double *total_ms, SSCInt *napply);

/* ---- device plumbing for callers that keep vectors resident (no torch needed) -------------- */
int SSCDeviceMalloc(SSCPatchPC pc, SSCInt bytes, void **ptr);
int SSCDeviceFree(SSCPatchPC pc, void *ptr);
int SSCHostAlloc(SSCPatchPC pc, SSCInt bytes, void **ptr);  /* pinned */
int SSCHostFree(SSCPatchPC pc, void *ptr);
int SSCMemcpy(SSCPatchPC pc, void *dst, const void *src, SSCInt bytes, int dstSpace, int srcSpace);
int SSCDeviceFlushL2(SSCPatchPC pc);
/* streaming-read bandwidth (GB/s) of this GPU over the resident blocks: roofline calibration, best of `reps` passes */
int SSCDeviceMeasureReadBandwidth(SSCPatchPC pc, int reps, double *gbs);                        /* overwrite a buffer larger than L2 (bench hygiene) */
/* CUDA events on the PC's stream */
int SSCEventCreate(SSCPatchPC pc, void **ev);
int SSCEventRecord(SSCPatchPC pc, void *ev);
int SSCEventElapsed(SSCPatchPC pc, void *start, void *stop, double *ms);  /* synchronises on stop */
int SSCEventDestroy(SSCPatchPC pc, void *ev);

/* ---- low-order (P1) coarse correction, SURVEY.md 8(f) N1: class P1PC of ssc/lo.py:105-198 ------------------ */
/* y = R^T A1^{-1} R x with Dirichlet values carried over.  The host assembles R (ssc/lo.py:83-102) and the coarse
 * operator A1 (ssc/lo.py:142-144); the device does the two sparse products and the coarse solve.  The reference hands A1 to
 * the KSP/PC its options name (ssc/lo.py:164-166); here, under the same option names:
 *   lo_ksp_type preonly + lo_pc_type lu | cholesky : dense factorisation on the device for coarse spaces up to 4096 dofs; beyond,
 *       the exact solve these options ask for is multigrid-preconditioned CG driven to 1e-12 (below);
 *   lo_ksp_type cg + lo_pc_type jacobi [lo_ksp_rtol 1e-10, lo_ksp_max_it 10000] : device-resident preconditioned conjugate
 *       gradients on the coarse CSR operator, any size (the Q1 space of a 64^3 mesh has 274 625 dofs);
 *   lo_pc_type gamg : smoothed-aggregation multigrid (hierarchy built on the host at SetUp, V(1,1) cycle on the device), under
 *       lo_ksp_type cg as the preconditioner of that CG, under lo_ksp_type preonly as ONE cycle per apply (a fixed symmetric
 *       positive operator).  lo_amg_block_size bs (default 1): the coarse space is vector-valued, dof = bs * node + component; the
 *       hierarchy then aggregates nodes and carries the bs componentwise constants, or the vectors given to
 *       SSCLowOrderSetNearNullSpace (rigid-body modes for elasticity), orthonormalised aggregate by aggregate.
 * lo_mat_type is accepted and ignored. */
typedef struct SSCLowOrderPC_s *SSCLowOrderPC;
int SSCLowOrderCreate(int device, SSCLowOrderPC *lo);                       /* P1PC.__init__            ssc/lo.py:107-110 */
int SSCLowOrderDestroy(SSCLowOrderPC *lo);
int SSCLowOrderSetOption(SSCLowOrderPC lo, const char *name, const char *value);   /* lo.setFromOptions ssc/lo.py:164-167 */
int SSCLowOrderSetRestriction(SSCLowOrderPC lo, SSCInt ncoarse, SSCInt nfine, const SSCInt *rowptr,
                              const int32_t *colidx, const double *vals);   /* restriction_matrix       ssc/lo.py:83-102 */
int SSCLowOrderSetOperator(SSCLowOrderPC lo, SSCInt n, const SSCInt *rowptr, const int32_t *colidx,
                           const double *vals);                             /* lo_op                    ssc/lo.py:142-144 */
int SSCLowOrderSetBcs(SSCLowOrderPC lo, SSCInt nbc, const SSCInt *bcIndices);      /* bc_indices        ssc/lo.py:175-183 */
int SSCLowOrderSetNearNullSpace(SSCLowOrderPC lo, SSCInt n, int k, const double *vecs);   /* setNearNullSpace of the coarse matrix, ssc/lo.py:146-159: k vectors of n entries, used by lo_pc_type gamg */
int SSCLowOrderSetUp(SSCLowOrderPC lo);                                     /* initialize / update      ssc/lo.py:112-186 */
int SSCLowOrderApply(SSCLowOrderPC lo, const double *x, double *y, int memspace);  /* apply             ssc/lo.py:187-198 */
int SSCLowOrderGetIterations(SSCLowOrderPC lo, SSCInt *last, SSCInt *total, SSCInt *solves);   /* coarse CG iterations (KSPGetIterationNumber of the lo KSP) */
int SSCLowOrderAXPY(SSCLowOrderPC lo, SSCInt n, double a, const double *x, double *y);  /* device y += a x on the coarse PC's stream: the additive PCCOMPOSITE of ssc/ssc.py:38-49 */
int SSCLowOrderSynchronize(SSCLowOrderPC lo);

#ifdef __cplusplus
}
#endif
#endif
