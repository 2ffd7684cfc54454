/*
 * ssc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the patch preconditioner of wence-/ssc
 * (ssc/libssc.c), used only as the checker by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Nothing under
 * ssc_b200/ may load it.
 *
 * Parity status: the reference cannot be compiled here (it needs PETSc private
 * headers, ssc/libssc.c:1-4) and its tests hold no golden vectors.  The known
 * answers the reference's own tests hold for this path are both reproduced:
 *  (i) "vertex-star patches at degree 1 are point Jacobi", scalar / vector /
 *      tensor / mixed on its three meshes (tests/test_jacobi_equivalent.py:85)
 *      -> tests/test_oracle_known_answers.py;
 * (ii) the twelve CG iteration counts of ssc.SSC for degrees 1..4
 *      (tests/test_p_independence.py:17-24: [5,5,5,5], [10,12,12,12], [4,21,22,22])
 *      -> tests/test_p_independence.py, with this oracle as the patch PC.
 * (ii) pins PCApply for degrees 2-4 through integer outcomes, not through
 * vectors: no reference-produced PCApply vector exists to compare with, and the
 * order inside gtol remains defined by this oracle.
 *
 * What is restated, with the lines it follows:
 *   transitive closure / star      DMPlexGetTransitiveClosure call sites :344 :348 :1460 (PETSc, external)
 *   PCPatchConstruct_Star          ssc/libssc.c:1447-1469
 *   PCPatchConstruct_Vanka         ssc/libssc.c:1471-1519
 *   PCPatchConstruct_User          ssc/libssc.c:1522-1547
 *   PCPatchCompleteCellPatch       ssc/libssc.c:325-363
 *   PCPatchGetPointDofs            ssc/libssc.c:370-416
 *   PCPatchComputeSetDifference    ssc/libssc.c:420-440
 *   PCPatchCreateCellPatches       ssc/libssc.c:452-580
 *   PCPatchCreateCellPatchDiscretisationInfo   ssc/libssc.c:598-900
 *   patch operator by assembly     ssc/libssc.c:1120-1156 + ssc/patch.py:28-62 (MatSetValues, -1 ignored)
 *   PCPatch_ScatterLocal_Private   ssc/libssc.c:1158-1198
 *   partition-of-unity weights     ssc/libssc.c:1254-1284
 *   PCApply_PATCH                  ssc/libssc.c:1301-1426
 *   sub-KSP preonly + LU           KSPSolve at ssc/libssc.c:1374 (PETSc, external, un-pinned):
 *                                  restated as dense LU with partial pivoting (LAPACK dgetrf/dgetrs algorithm).
 *
 * Documented deviation: the reference iterates PETSc khash tables in bucket order
 * (:336-341, :532-539, :562-570, :828-836); that order is a PETSc implementation detail that is
 * not available here.  Every such iteration is done in ASCENDING KEY ORDER instead.
 * The dof *sets* do not depend on it; the order inside gtol does.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

/* ------------------------------------------------------------------ */
/* integer hash map with PetscHMapI semantics (Get returns -1 if absent) */
/* ------------------------------------------------------------------ */
typedef struct { i64 *keys, *vals; unsigned char *used; i64 cap, size; } hmap;

static void hmap_init(hmap *h) { h->cap = 64; h->size = 0; h->keys = malloc(h->cap * sizeof(i64)); h->vals = malloc(h->cap * sizeof(i64)); h->used = calloc(h->cap, 1); }
static void hmap_free(hmap *h) { free(h->keys); free(h->vals); free(h->used); }
static void hmap_clear(hmap *h) { memset(h->used, 0, h->cap); h->size = 0; }
static i64 hmap_slot(const hmap *h, i64 key) {
    uint64_t x = (uint64_t)key * 0x9E3779B97F4A7C15ull;
    i64 s = (i64)(x >> 20) & (h->cap - 1);
    while (h->used[s] && h->keys[s] != key) s = (s + 1) & (h->cap - 1);
    return s;
}
static void hmap_set(hmap *h, i64 key, i64 val);
static void hmap_grow(hmap *h) {
    hmap o = *h; h->cap = o.cap * 2; h->size = 0;
    h->keys = malloc(h->cap * sizeof(i64)); h->vals = malloc(h->cap * sizeof(i64)); h->used = calloc(h->cap, 1);
    for (i64 i = 0; i < o.cap; i++) if (o.used[i]) hmap_set(h, o.keys[i], o.vals[i]);
    hmap_free(&o);
}
static void hmap_set(hmap *h, i64 key, i64 val) {
    if (2 * (h->size + 1) > h->cap) hmap_grow(h);
    i64 s = hmap_slot(h, key);
    if (!h->used[s]) { h->used[s] = 1; h->keys[s] = key; h->size++; }
    h->vals[s] = val;
}
static i64 hmap_get(const hmap *h, i64 key) { i64 s = hmap_slot(h, key); return h->used[s] ? h->vals[s] : -1; }
static int hmap_has(const hmap *h, i64 key) { return h->used[hmap_slot(h, key)]; }
static int cmp_i64(const void *a, const void *b) { i64 x = *(const i64 *)a, y = *(const i64 *)b; return (x > y) - (x < y); }
/* ascending-key iteration: caller frees */
static i64 *hmap_sorted_keys(const hmap *h) {
    i64 *k = malloc((h->size + 1) * sizeof(i64)), n = 0;
    for (i64 i = 0; i < h->cap; i++) if (h->used[i]) k[n++] = h->keys[i];
    qsort(k, n, sizeof(i64), cmp_i64);
    return k;
}

/* ------------------------------------------------------------------ */
typedef struct {
    /* topology (borrowed) */
    i64 dim, pEnd;
    const i64 *coneOff, *cones, *suppOff, *supports, *depthStart, *depthEnd;
    const unsigned char *ghost;
    /* cell numbering section (borrowed) */
    const i64 *cnDof, *cnOff;
    /* discretisation info: PCPatchSetDiscretisationInfo ssc/libssc.c:237-279 */
    i64 nsubspaces, totalDofsPerCell;
    const i64 **secDof, **secOff;      /* borrowed */
    const i64 **cellNodeMap;           /* borrowed (:268) */
    i64 *bs, *nodesPerCell, *subspaceOffsets;          /* copied (:265-269, :275) */
    i64 numGhostBcs, *ghostBcNodes, numGlobalBcs, *globalBcNodes; /* copied (:276-277) */
    /* options, defaults of PCCreate_PATCH ssc/libssc.c:1697-1711 */
    int construct_type;                /* 0 star, 1 vanka, 2 user */
    i64 codim, cdim, exclude_subspace, vankadim;
    int partition_of_unity, symmetrise_sweep;
    /* user patches */
    i64 nuserIS; const i64 *userOff, *userPts; i64 niter; const i64 *iterationSet;
    /* outputs of setup */
    i64 vStart, vEnd, npatch;
    i64 *cellDof, *cellOff, *cells, numCells;
    i64 *gtolDof, *gtolOff, *gtol, numGtol;
    i64 *dofs, numDofs;
    /* operator */
    i64 nrows; const i64 *rowptr, *colidx; const int32_t *colidx32; const double *vals;
    double **mat; double **lu; i64 **piv;
    double *dof_weights; i64 nowned;
    char err[256];
} OraclePC;

OraclePC *oracle_create(void) {
    OraclePC *pc = calloc(1, sizeof(OraclePC));
    pc->codim = -1; pc->cdim = -1; pc->exclude_subspace = -1; pc->vankadim = -1;
    return pc;
}
const char *oracle_error(OraclePC *pc) { return pc->err; }

static void free_operators(OraclePC *pc) {
    if (pc->mat) { for (i64 i = 0; i < pc->npatch; i++) { free(pc->mat[i]); free(pc->lu[i]); free(pc->piv[i]); } }
    free(pc->mat); free(pc->lu); free(pc->piv); pc->mat = pc->lu = NULL; pc->piv = NULL;
}
void oracle_destroy(OraclePC *pc) {
    free_operators(pc);
    free(pc->secDof); free(pc->secOff); free(pc->cellNodeMap); free(pc->bs); free(pc->nodesPerCell);
    free(pc->subspaceOffsets); free(pc->ghostBcNodes); free(pc->globalBcNodes);
    free(pc->cellDof); free(pc->cellOff); free(pc->cells); free(pc->gtolDof); free(pc->gtolOff);
    free(pc->gtol); free(pc->dofs); free(pc->dof_weights);
    free(pc);
}

int oracle_set_plex(OraclePC *pc, i64 dim, i64 pEnd, const i64 *coneOff, const i64 *cones, const i64 *suppOff,
                    const i64 *supports, const i64 *depthStart, const i64 *depthEnd, const unsigned char *ghost) {
    pc->dim = dim; pc->pEnd = pEnd; pc->coneOff = coneOff; pc->cones = cones; pc->suppOff = suppOff;
    pc->supports = supports; pc->depthStart = depthStart; pc->depthEnd = depthEnd; pc->ghost = ghost;
    return 0;
}
/* PCPatchSetCellNumbering ssc/libssc.c:225-234 */
int oracle_set_cell_numbering(OraclePC *pc, const i64 *dof, const i64 *off) { pc->cnDof = dof; pc->cnOff = off; return 0; }

/* PCPatchSetDiscretisationInfo ssc/libssc.c:237-279 (the DMs are replaced by their default sections) */
int oracle_set_discretisation_info(OraclePC *pc, i64 nsubspaces, const i64 **secDof, const i64 **secOff, const i64 *bs,
                                   const i64 *nodesPerCell, const i64 **cellNodeMap, const i64 *subspaceOffsets,
                                   i64 numGhostBcs, const i64 *ghostBcNodes, i64 numGlobalBcs, const i64 *globalBcNodes) {
    pc->nsubspaces = nsubspaces; pc->totalDofsPerCell = 0;
    pc->secDof = malloc(nsubspaces * sizeof(i64 *)); pc->secOff = malloc(nsubspaces * sizeof(i64 *));
    pc->cellNodeMap = malloc(nsubspaces * sizeof(i64 *));
    pc->bs = malloc(nsubspaces * sizeof(i64)); pc->nodesPerCell = malloc(nsubspaces * sizeof(i64));
    pc->subspaceOffsets = malloc((nsubspaces + 1) * sizeof(i64));
    for (i64 i = 0; i < nsubspaces; i++) {
        pc->secDof[i] = secDof[i]; pc->secOff[i] = secOff[i]; pc->bs[i] = bs[i]; pc->nodesPerCell[i] = nodesPerCell[i];
        pc->totalDofsPerCell += nodesPerCell[i] * bs[i];
        pc->cellNodeMap[i] = cellNodeMap[i]; pc->subspaceOffsets[i] = subspaceOffsets[i];
    }
    pc->subspaceOffsets[nsubspaces] = subspaceOffsets[nsubspaces];
    pc->numGhostBcs = numGhostBcs; pc->ghostBcNodes = malloc((numGhostBcs + 1) * sizeof(i64));
    memcpy(pc->ghostBcNodes, ghostBcNodes, numGhostBcs * sizeof(i64));
    pc->numGlobalBcs = numGlobalBcs; pc->globalBcNodes = malloc((numGlobalBcs + 1) * sizeof(i64));
    memcpy(pc->globalBcNodes, globalBcNodes, numGlobalBcs * sizeof(i64));
    return 0;
}

/* options of PCSetFromOptions_PATCH ssc/libssc.c:1551-1620 that change results */
int oracle_set_options(OraclePC *pc, int construct_type, i64 cdim, i64 codim, i64 exclude_subspace, i64 vankadim,
                       int partition_of_unity, int symmetrise_sweep) {
    if (cdim >= 0 && codim >= 0) { snprintf(pc->err, 256, "Can only set one of dimension or co-dimension"); return 83; } /* :1576 */
    pc->construct_type = construct_type; pc->cdim = cdim; pc->codim = codim; pc->exclude_subspace = exclude_subspace;
    pc->vankadim = vankadim; pc->partition_of_unity = partition_of_unity; pc->symmetrise_sweep = symmetrise_sweep;
    return 0;
}
int oracle_set_user_patches(OraclePC *pc, i64 n, const i64 *off, const i64 *pts, i64 niter, const i64 *iter) {
    pc->nuserIS = n; pc->userOff = off; pc->userPts = pts; pc->niter = niter; pc->iterationSet = iter; return 0;
}

/* breadth-first transitive closure without repeats; first entry is the point itself */
static i64 closure(const OraclePC *pc, i64 p, int useCone, i64 **buf, i64 *cap, hmap *mark) {
    const i64 *off = useCone ? pc->coneOff : pc->suppOff, *adj = useCone ? pc->cones : pc->supports;
    i64 n = 0, head = 0;
    hmap_clear(mark);
    if (*cap < 16) { *cap = 16; *buf = realloc(*buf, *cap * sizeof(i64)); }
    (*buf)[n++] = p; hmap_set(mark, p, 0);
    while (head < n) {
        i64 q = (*buf)[head++];
        for (i64 k = off[q]; k < off[q + 1]; k++) {
            i64 r = adj[k];
            if (hmap_has(mark, r)) continue;
            hmap_set(mark, r, 0);
            if (n == *cap) { *cap *= 2; *buf = realloc(*buf, *cap * sizeof(i64)); }
            (*buf)[n++] = r;
        }
    }
    return n;
}

typedef struct { i64 *star, *clos; i64 scap, ccap; hmap mark; } work;
static void work_init(work *w) { w->star = w->clos = NULL; w->scap = w->ccap = 0; hmap_init(&w->mark); }
static void work_free(work *w) { free(w->star); free(w->clos); hmap_free(&w->mark); }

/* ssc/libssc.c:1447-1469 */
static void construct_star(const OraclePC *pc, i64 entity, hmap *ht, work *w) {
    hmap_clear(ht);
    hmap_set(ht, entity, 0);
    i64 n = closure(pc, entity, 0, &w->star, &w->scap, &w->mark);
    for (i64 si = 0; si < n; si++) hmap_set(ht, w->star[si], 0);
}
/* ssc/libssc.c:1471-1519 */
static void construct_vanka(const OraclePC *pc, i64 entity, hmap *ht, work *w) {
    i64 iStart = 0, iEnd = 0, cStart = pc->depthStart[pc->dim], cEnd = pc->depthEnd[pc->dim];
    int shouldIgnore = pc->vankadim >= 0;
    hmap_clear(ht);
    hmap_set(ht, entity, 0);
    if (shouldIgnore) { iStart = pc->depthStart[pc->vankadim]; iEnd = pc->depthEnd[pc->vankadim]; }
    i64 n = closure(pc, entity, 0, &w->star, &w->scap, &w->mark);
    for (i64 si = 0; si < n; si++) {
        i64 cell = w->star[si];
        if (cell < cStart || cell >= cEnd) continue;
        i64 m = closure(pc, cell, 1, &w->clos, &w->ccap, &w->mark);
        for (i64 ci = 0; ci < m; ci++) {
            i64 e = w->clos[ci];
            if (shouldIgnore && iStart <= e && e < iEnd) continue;
            hmap_set(ht, e, 0);
        }
    }
}
/* ssc/libssc.c:1522-1547 */
static int construct_user(OraclePC *pc, i64 entity, hmap *ht) {
    hmap_clear(ht);
    for (i64 i = pc->userOff[entity]; i < pc->userOff[entity + 1]; i++) {
        i64 e = pc->userPts[i];
        if (e < 0 || e >= pc->pEnd) { snprintf(pc->err, 256, "Entities need to be between the bounds of DMPlexGetChart()"); return 83; }
        hmap_set(ht, e, 0);
    }
    return 0;
}
static int construct(OraclePC *pc, i64 entity, hmap *ht, work *w) {
    if (pc->construct_type == 0) construct_star(pc, entity, ht, w);
    else if (pc->construct_type == 1) construct_vanka(pc, entity, ht, w);
    else return construct_user(pc, entity, ht);
    return 0;
}

/* ssc/libssc.c:325-363 */
static void complete_cell_patch(const OraclePC *pc, const hmap *ht, hmap *cht, work *w) {
    hmap_clear(cht);
    i64 *keys = hmap_sorted_keys(ht);
    for (i64 k = 0; k < ht->size; k++) {
        i64 n = closure(pc, keys[k], 0, &w->star, &w->scap, &w->mark);
        for (i64 si = 0; si < n; si++) {
            i64 m = closure(pc, w->star[si], 1, &w->clos, &w->ccap, &w->mark);
            for (i64 ci = 0; ci < m; ci++) hmap_set(cht, w->clos[ci], 0);
        }
    }
    free(keys);
}

/* ssc/libssc.c:370-416 */
static void get_point_dofs(const OraclePC *pc, const hmap *pts, hmap *dofs, i64 base, i64 exclude_subspace) {
    hmap_clear(dofs);
    i64 *keys = hmap_sorted_keys(pts);
    for (i64 k = 0; k < pc->nsubspaces; k++) {
        const i64 *sd = pc->secDof[k], *so = pc->secOff[k];
        i64 bs = pc->bs[k], subspaceOffset = pc->subspaceOffsets[k];
        if (k == exclude_subspace) {
            i64 ldof = sd[base], loff = so[base];
            for (i64 j = loff; j < ldof + loff; j++) for (i64 l = 0; l < bs; l++) hmap_set(dofs, bs * j + l + subspaceOffset, 0);
            continue;
        }
        for (i64 q = 0; q < pts->size; q++) {
            i64 p = keys[q], ldof = sd[p], loff = so[p];
            for (i64 j = loff; j < ldof + loff; j++) for (i64 l = 0; l < bs; l++) hmap_set(dofs, bs * j + l + subspaceOffset, 0);
        }
    }
    free(keys);
}
/* ssc/libssc.c:420-440: keys of B not in A */
static void set_difference(const hmap *A, const hmap *B, hmap *C) {
    hmap_clear(C);
    for (i64 i = 0; i < B->cap; i++) if (B->used[i] && !hmap_has(A, B->keys[i])) hmap_set(C, B->keys[i], 0);
}

/* ssc/libssc.c:452-580 */
static int create_cell_patches(OraclePC *pc) {
    hmap ht, cht; work w;
    i64 cStart = pc->depthStart[pc->dim], cEnd = pc->depthEnd[pc->dim], vStart, vEnd;
    int ierr = 0;
    if (pc->construct_type == 2) { vStart = 0; vEnd = pc->nuserIS; }
    else if (pc->codim < 0) {
        i64 d = pc->cdim < 0 ? 0 : pc->cdim;
        vStart = pc->depthStart[d]; vEnd = pc->depthEnd[d];
    } else { vStart = pc->depthStart[pc->dim - pc->codim]; vEnd = pc->depthEnd[pc->dim - pc->codim]; }
    hmap_init(&ht); hmap_init(&cht); work_init(&w);
    pc->vStart = vStart; pc->vEnd = vEnd; pc->npatch = vEnd - vStart;
    pc->cellDof = calloc(pc->npatch + 1, sizeof(i64)); pc->cellOff = calloc(pc->npatch + 1, sizeof(i64));
    for (i64 v = vStart; v < vEnd; v++) {                                   /* count pass :511-541 */
        if (pc->construct_type != 2 && pc->ghost && pc->ghost[v]) continue; /* :515-521 */
        if ((ierr = construct(pc, v, &ht, &w))) goto done;
        complete_cell_patch(pc, &ht, &cht, &w);
        if (cht.size == 0) continue;
        for (i64 i = 0; i < cht.cap; i++) if (cht.used[i] && cStart <= cht.keys[i] && cht.keys[i] < cEnd) pc->cellDof[v - vStart]++;
    }
    i64 tot = 0;
    for (i64 i = 0; i < pc->npatch; i++) { pc->cellOff[i] = tot; tot += pc->cellDof[i]; }
    pc->numCells = tot; pc->cells = malloc((tot + 1) * sizeof(i64));
    for (i64 v = vStart; v < vEnd; v++) {                                   /* fill pass :550-572 */
        i64 ndof = pc->cellDof[v - vStart], off = pc->cellOff[v - vStart];
        if (ndof <= 0) continue;
        if ((ierr = construct(pc, v, &ht, &w))) goto done;
        complete_cell_patch(pc, &ht, &cht, &w);
        i64 *keys = hmap_sorted_keys(&cht);
        ndof = 0;
        for (i64 i = 0; i < cht.size; i++) if (cStart <= keys[i] && keys[i] < cEnd) pc->cells[off + ndof++] = keys[i];
        free(keys);
    }
done:
    hmap_free(&ht); hmap_free(&cht); work_free(&w);
    return ierr;
}

/* ssc/libssc.c:598-900 */
static int create_discretisation_info(OraclePC *pc) {
    i64 totalDofsPerCell = pc->totalDofsPerCell, numDofs = pc->numCells * totalDofsPerCell, globalIndex = 0, vStart = pc->vStart;
    i64 *dofsArray = malloc((numDofs + 1) * sizeof(i64)), *asmArray = malloc((numDofs + 1) * sizeof(i64));
    i64 *newCells = malloc((pc->numCells + 1) * sizeof(i64));
    hmap ht, globalBcs, ownedpts, seenpts, owneddofs, seendofs, artificialbcs; work w;
    int ierr = 0;
    pc->gtolDof = calloc(pc->npatch + 1, sizeof(i64)); pc->gtolOff = calloc(pc->npatch + 1, sizeof(i64));
    hmap_init(&ht); hmap_init(&globalBcs); hmap_init(&ownedpts); hmap_init(&seenpts); hmap_init(&owneddofs);
    hmap_init(&seendofs); hmap_init(&artificialbcs); work_init(&w);
    for (i64 i = 0; i < pc->numGhostBcs; i++) hmap_set(&globalBcs, pc->ghostBcNodes[i], 0);    /* :642-648 */
    for (i64 v = pc->vStart; v < pc->vEnd; v++) {
        i64 dof = pc->cellDof[v - vStart], off = pc->cellOff[v - vStart], localIndex = 0;
        hmap_clear(&ht);
        if (dof <= 0) continue;
        if ((ierr = construct(pc, v, &ownedpts, &w))) goto done;                               /* :669 */
        complete_cell_patch(pc, &ownedpts, &seenpts, &w);                                      /* :670 */
        get_point_dofs(pc, &ownedpts, &owneddofs, v, pc->exclude_subspace);                    /* :671 */
        get_point_dofs(pc, &seenpts, &seendofs, v, -1);                                        /* :672 */
        set_difference(&owneddofs, &seendofs, &artificialbcs);                                 /* :673 */
        for (i64 k = 0; k < pc->nsubspaces; k++) {                                             /* :731-778 */
            i64 nodesPerCell = pc->nodesPerCell[k], subspaceOffset = pc->subspaceOffsets[k], bs = pc->bs[k];
            const i64 *cellNodeMap = pc->cellNodeMap[k];
            for (i64 i = off; i < off + dof; i++) {
                i64 c = pc->cells[i];
                if (pc->cnDof[c] <= 0) { snprintf(pc->err, 256, "Cell doesn't appear in cell numbering map"); ierr = 63; goto done; }
                i64 cell = pc->cnOff[c];
                newCells[i] = cell;
                for (i64 j = 0; j < nodesPerCell; j++) {
                    i64 globalDof = cellNodeMap[cell * nodesPerCell + j] * bs + subspaceOffset;
                    for (i64 l = 0; l < bs; l++) {
                        if (hmap_get(&globalBcs, globalDof + l) >= 0 || hmap_get(&artificialbcs, globalDof + l) >= 0) {
                            dofsArray[globalIndex++] = -1;
                        } else {
                            i64 localDof = hmap_get(&ht, globalDof + l);
                            if (localDof == -1) { localDof = localIndex++; hmap_set(&ht, globalDof + l, localDof); }
                            if (globalIndex >= numDofs) { snprintf(pc->err, 256, "Found more dofs than expected"); ierr = 63; goto done; }
                            dofsArray[globalIndex++] = localDof;
                        }
                    }
                }
            }
        }
        pc->gtolDof[v - vStart] = ht.size;                                                     /* :779-781 */
    }
    if (globalIndex != numDofs) { snprintf(pc->err, 256, "Expected number of dofs (%lld) doesn't match found number (%lld)", (long long)numDofs, (long long)globalIndex); ierr = 63; goto done; }
    i64 tot = 0;
    for (i64 i = 0; i < pc->npatch; i++) { pc->gtolOff[i] = tot; tot += pc->gtolDof[i]; }
    pc->numGtol = tot; pc->gtol = malloc((tot + 1) * sizeof(i64));
    i64 key = 0, asmKey = 0;
    for (i64 v = pc->vStart; v < pc->vEnd; v++) {                                              /* :796-873 */
        i64 dof = pc->cellDof[v - vStart], off = pc->cellOff[v - vStart];
        hmap_clear(&ht);
        if (dof <= 0) continue;
        for (i64 k = 0; k < pc->nsubspaces; k++) {
            i64 nodesPerCell = pc->nodesPerCell[k], subspaceOffset = pc->subspaceOffsets[k], bs = pc->bs[k];
            const i64 *cellNodeMap = pc->cellNodeMap[k];
            for (i64 i = off; i < off + dof; i++) {
                i64 cell = pc->cnOff[pc->cells[i]];
                for (i64 j = 0; j < nodesPerCell; j++) for (i64 l = 0; l < bs; l++) {
                    i64 globalDof = cellNodeMap[cell * nodesPerCell + j] * bs + subspaceOffset + l;
                    i64 localDof = dofsArray[key++];
                    if (localDof >= 0) hmap_set(&ht, globalDof, localDof);
                }
            }
            i64 goff = pc->gtolOff[v - vStart];
            for (i64 s = 0; s < ht.cap; s++) if (ht.used[s] && ht.keys[s] >= 0) pc->gtol[goff + ht.vals[s]] = ht.keys[s];
        }
        if (pc->nsubspaces > 1) {                                                              /* :846-872 cell-major table */
            for (i64 i = off; i < off + dof; i++) {
                i64 cell = pc->cnOff[pc->cells[i]];
                for (i64 k = 0; k < pc->nsubspaces; k++) {
                    i64 nodesPerCell = pc->nodesPerCell[k], subspaceOffset = pc->subspaceOffsets[k], bs = pc->bs[k];
                    const i64 *cellNodeMap = pc->cellNodeMap[k];
                    for (i64 j = 0; j < nodesPerCell; j++) for (i64 l = 0; l < bs; l++)
                        asmArray[asmKey++] = hmap_get(&ht, cellNodeMap[cell * nodesPerCell + j] * bs + subspaceOffset + l);
                }
            }
        }
    }
    if (pc->nsubspaces == 1) memcpy(asmArray, dofsArray, numDofs * sizeof(i64));               /* :875-879 */
    memcpy(pc->cells, newCells, pc->numCells * sizeof(i64));                                   /* :887 */
    pc->dofs = asmArray; asmArray = NULL; pc->numDofs = numDofs;
done:
    free(dofsArray); free(asmArray); free(newCells);
    hmap_free(&ht); hmap_free(&globalBcs); hmap_free(&ownedpts); hmap_free(&seenpts); hmap_free(&owneddofs);
    hmap_free(&seendofs); hmap_free(&artificialbcs); work_free(&w);
    return ierr;
}

/* first-time part of PCSetUp_PATCH ssc/libssc.c:1210-1250 */
int oracle_setup_patches(OraclePC *pc) {
    int ierr;
    free(pc->cellDof); free(pc->cellOff); free(pc->cells); free(pc->gtolDof); free(pc->gtolOff); free(pc->gtol); free(pc->dofs);
    pc->cellDof = pc->cellOff = pc->cells = pc->gtolDof = pc->gtolOff = pc->gtol = pc->dofs = NULL;
    if ((ierr = create_cell_patches(pc))) return ierr;
    return create_discretisation_info(pc);
}

i64 oracle_npatch(OraclePC *pc) { return pc->npatch; }
i64 oracle_vstart(OraclePC *pc) { return pc->vStart; }
i64 oracle_num_cells(OraclePC *pc) { return pc->numCells; }
i64 oracle_num_gtol(OraclePC *pc) { return pc->numGtol; }
i64 oracle_num_dofs(OraclePC *pc) { return pc->numDofs; }
const i64 *oracle_cell_dof(OraclePC *pc) { return pc->cellDof; }
const i64 *oracle_cell_off(OraclePC *pc) { return pc->cellOff; }
const i64 *oracle_cells(OraclePC *pc) { return pc->cells; }
const i64 *oracle_gtol_dof(OraclePC *pc) { return pc->gtolDof; }
const i64 *oracle_gtol_off(OraclePC *pc) { return pc->gtolOff; }
const i64 *oracle_gtol(OraclePC *pc) { return pc->gtol; }
const i64 *oracle_dofs(OraclePC *pc) { return pc->dofs; }

/* a caller that already has the dof lists (e.g. a golden fixture) can install them directly */
int oracle_set_patches(OraclePC *pc, i64 npatch, const i64 *gtolOff, const i64 *gtol) {
    free(pc->gtolDof); free(pc->gtolOff); free(pc->gtol);
    pc->npatch = npatch; pc->vStart = 0; pc->vEnd = npatch;
    pc->gtolDof = malloc((npatch + 1) * sizeof(i64)); pc->gtolOff = malloc((npatch + 1) * sizeof(i64));
    for (i64 i = 0; i < npatch; i++) { pc->gtolOff[i] = gtolOff[i]; pc->gtolDof[i] = gtolOff[i + 1] - gtolOff[i]; }
    pc->numGtol = gtolOff[npatch]; pc->gtol = malloc((pc->numGtol + 1) * sizeof(i64));
    memcpy(pc->gtol, gtol, pc->numGtol * sizeof(i64));
    return 0;
}

/* ------------------------------------------------------------------ */
/* patch operators                                                     */
/* ------------------------------------------------------------------ */
int oracle_set_operator_csr(OraclePC *pc, i64 nrows, const i64 *rowptr, const i64 *colidx, const double *vals) {
    pc->nrows = nrows; pc->rowptr = rowptr; pc->colidx = colidx; pc->colidx32 = NULL; pc->vals = vals; return 0;
}
/* same with 32-bit column indices (benchmark-size operators) */
int oracle_set_operator_csr32(OraclePC *pc, i64 nrows, const i64 *rowptr, const int32_t *colidx, const double *vals) {
    pc->nrows = nrows; pc->rowptr = rowptr; pc->colidx = NULL; pc->colidx32 = colidx; pc->vals = vals; return 0;
}

/* B_p = A[gtol_p, gtol_p]: equal to the element-wise assembly of ssc/libssc.c:1120-1156 for star patches */
static void extract_patch(const OraclePC *pc, i64 p, double *B, hmap *g2l) {
    i64 n = pc->gtolDof[p]; const i64 *g = pc->gtol + pc->gtolOff[p];
    hmap_clear(g2l);
    for (i64 i = 0; i < n; i++) hmap_set(g2l, g[i], i);
    memset(B, 0, n * n * sizeof(double));
    for (i64 i = 0; i < n; i++)
        for (i64 k = pc->rowptr[g[i]]; k < pc->rowptr[g[i] + 1]; k++) {
            i64 j = hmap_get(g2l, pc->colidx ? pc->colidx[k] : (i64)pc->colidx32[k]);
            if (j >= 0) B[i * n + j] += pc->vals[k];
        }
}

/* the reference's own route: MatSetValues(ADD_VALUES) of element matrices through the `dofs` table, negative
 * indices ignored (ssc/libssc.c:1150, ssc/patch.py:28-62).  Ke holds one tdpc x tdpc matrix per Firedrake cell. */
int oracle_patch_matrix_from_elements(OraclePC *pc, i64 p, const double *Ke, double *B) {
    i64 n = pc->gtolDof[p], t = pc->totalDofsPerCell;
    memset(B, 0, n * n * sizeof(double));
    for (i64 c = 0; c < pc->cellDof[p]; c++) {
        i64 cell = pc->cells[pc->cellOff[p] + c];
        const i64 *idx = pc->dofs + (pc->cellOff[p] + c) * t;
        const double *K = Ke + cell * t * t;
        for (i64 i = 0; i < t; i++) { if (idx[i] < 0) continue;
            for (i64 j = 0; j < t; j++) { if (idx[j] < 0) continue; B[idx[i] * n + idx[j]] += K[i * t + j]; } }
    }
    return 0;
}

/* dense LU with partial pivoting, row-major, in place (dgetrf algorithm, unblocked) */
static int lu_factor(double *A, i64 n, i64 *piv) {
    for (i64 k = 0; k < n; k++) {
        i64 p = k; double mx = fabs(A[k * n + k]);
        for (i64 i = k + 1; i < n; i++) if (fabs(A[i * n + k]) > mx) { mx = fabs(A[i * n + k]); p = i; }
        piv[k] = p;
        if (mx == 0.0) return 1;
        if (p != k) for (i64 j = 0; j < n; j++) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
        double inv = 1.0 / A[k * n + k];
        for (i64 i = k + 1; i < n; i++) {
            double l = A[i * n + k] * inv; A[i * n + k] = l;
            if (l != 0.0) for (i64 j = k + 1; j < n; j++) A[i * n + j] -= l * A[k * n + j];
        }
    }
    return 0;
}
static void lu_solve(const double *LU, const i64 *piv, i64 n, const double *b, double *x) {
    for (i64 i = 0; i < n; i++) x[i] = b[i];
    for (i64 k = 0; k < n; k++) { i64 p = piv[k]; if (p != k) { double t = x[k]; x[k] = x[p]; x[p] = t; } }
    for (i64 i = 0; i < n; i++) { double s = x[i]; const double *r = LU + i * n; for (i64 j = 0; j < i; j++) s -= r[j] * x[j]; x[i] = s; }
    for (i64 i = n - 1; i >= 0; i--) { double s = x[i]; const double *r = LU + i * n; for (i64 j = i + 1; j < n; j++) s -= r[j] * x[j]; x[i] = s / r[i]; }
}

/* per-patch operator + factorisation: ssc/libssc.c:1286-1292 plus the lazy PCSetUp of the sub-KSP inside :1374.
 * keep_mat != 0 keeps the unfactored blocks for inspection. */
int oracle_setup_operators(OraclePC *pc, int keep_mat, int nthreads) {
    free_operators(pc);
    pc->mat = calloc(pc->npatch + 1, sizeof(double *)); pc->lu = calloc(pc->npatch + 1, sizeof(double *)); pc->piv = calloc(pc->npatch + 1, sizeof(i64 *));
    int fail = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        hmap g2l; hmap_init(&g2l);
#pragma omp for schedule(dynamic, 64)
        for (i64 p = 0; p < pc->npatch; p++) {
            i64 n = pc->gtolDof[p];
            if (n <= 0) continue;
            pc->lu[p] = malloc(n * n * sizeof(double)); pc->piv[p] = malloc(n * sizeof(i64));
            extract_patch(pc, p, pc->lu[p], &g2l);
            if (keep_mat) { pc->mat[p] = malloc(n * n * sizeof(double)); memcpy(pc->mat[p], pc->lu[p], n * n * sizeof(double)); }
            if (lu_factor(pc->lu[p], n, pc->piv[p])) {
#pragma omp atomic write
                fail = 1;
            }
        }
        hmap_free(&g2l);
    }
    if (fail) { snprintf(pc->err, 256, "zero pivot in patch LU"); return 71; }
    return 0;
}
int oracle_get_patch_matrix(OraclePC *pc, i64 p, double *out) {
    if (!pc->mat || !pc->mat[p]) { snprintf(pc->err, 256, "patch matrix not kept"); return 73; }
    memcpy(out, pc->mat[p], pc->gtolDof[p] * pc->gtolDof[p] * sizeof(double)); return 0;
}

/* partition-of-unity weights on the LOCAL vector, before the SF reduce: ssc/libssc.c:1263-1273 */
int oracle_multiplicity_local(OraclePC *pc, i64 nlocal, double *local) {
    memset(local, 0, nlocal * sizeof(double));
    for (i64 i = 0; i < pc->npatch; i++) {
        if (pc->gtolDof[i] <= 0) continue;
        for (i64 l = 0; l < pc->gtolDof[i]; l++) local[pc->gtol[pc->gtolOff[i] + l]] += 1.0;    /* :1190 */
    }
    return 0;
}

/* the patch loop of PCApply_PATCH, ssc/libssc.c:1334-1388: localY = sum_p R_p^T B_p^{-1} R_p localX.
 * With nthreads > 1 the patches are split statically and each thread adds into a private copy (the
 * analogue of mpiexec -n nthreads); with nthreads == 1 the loop runs in the reference's order. */
int oracle_apply_local(OraclePC *pc, i64 nlocal, const double *localX, double *localY, int nthreads) {
    i64 np = pc->npatch, nsweep = pc->symmetrise_sweep ? 2 : 1;
    i64 niter = pc->construct_type == 2 ? pc->niter : np;
    memset(localY, 0, nlocal * sizeof(double));                                                 /* :1334 */
    if (!pc->lu) { snprintf(pc->err, 256, "operators not set up"); return 73; }
    if (nthreads <= 1) {
        i64 maxn = 0; for (i64 i = 0; i < np; i++) if (pc->gtolDof[i] > maxn) maxn = pc->gtolDof[i];
        double *px = malloc((maxn + 1) * sizeof(double)), *py = malloc((maxn + 1) * sizeof(double));
        for (i64 sweep = 0; sweep < nsweep; sweep++) {
            for (i64 jj = 0; jj < niter; jj++) {
                i64 j = sweep == 0 ? jj : niter - 1 - jj;                                       /* :1343 */
                i64 i = pc->construct_type == 2 ? pc->iterationSet[j] : j;                      /* :1347-1351 */
                i64 len = pc->gtolDof[i]; const i64 *g = pc->gtol + pc->gtolOff[i];
                if (len <= 0) continue;                                                         /* :1355 */
                for (i64 l = 0; l < len; l++) px[l] = localX[g[l]];                             /* :1188 */
                lu_solve(pc->lu[i], pc->piv[i], len, px, py);                                   /* :1374 */
                for (i64 l = 0; l < len; l++) localY[g[l]] += py[l];                            /* :1190 */
            }
        }
        free(px); free(py);
        return 0;
    }
#ifdef _OPENMP
    omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        i64 maxn = 0; for (i64 i = 0; i < np; i++) if (pc->gtolDof[i] > maxn) maxn = pc->gtolDof[i];
        double *px = malloc((maxn + 1) * sizeof(double)), *py = malloc((maxn + 1) * sizeof(double));
        double *mine = calloc(nlocal, sizeof(double));
#pragma omp for schedule(static)
        for (i64 jj = 0; jj < niter; jj++) {
            i64 i = pc->construct_type == 2 ? pc->iterationSet[jj] : jj;
            i64 len = pc->gtolDof[i]; const i64 *g = pc->gtol + pc->gtolOff[i];
            if (len <= 0) continue;
            for (i64 l = 0; l < len; l++) px[l] = localX[g[l]];
            lu_solve(pc->lu[i], pc->piv[i], len, px, py);
            for (i64 l = 0; l < len; l++) mine[g[l]] += nsweep * py[l];
        }
#pragma omp critical
        for (i64 k = 0; k < nlocal; k++) localY[k] += mine[k];
        free(px); free(py); free(mine);
    }
    return 0;
}

/* Test scaffolding for the per-entry parity check (tests/helpers.entry_err), not a reference function:
 * scale[i] = sum over the patches p containing local dof i of max_l |(B_p^{-1} R_p x)_l| -- the magnitude the
 * patch solves of ssc/libssc.c:1374 that are added into entry i (:1190) work at.  Sweeps count twice (:1336-1343). */
int oracle_apply_scale(OraclePC *pc, i64 nlocal, const double *localX, double *scale) {
    i64 np = pc->npatch, nsweep = pc->symmetrise_sweep ? 2 : 1;
    i64 niter = pc->construct_type == 2 ? pc->niter : np;
    memset(scale, 0, nlocal * sizeof(double));
    if (!pc->lu) { snprintf(pc->err, 256, "operators not set up"); return 73; }
    i64 maxn = 0; for (i64 i = 0; i < np; i++) if (pc->gtolDof[i] > maxn) maxn = pc->gtolDof[i];
    double *px = malloc((maxn + 1) * sizeof(double)), *py = malloc((maxn + 1) * sizeof(double));
    for (i64 jj = 0; jj < niter; jj++) {
        i64 i = pc->construct_type == 2 ? pc->iterationSet[jj] : jj;
        i64 len = pc->gtolDof[i]; const i64 *g = pc->gtol + pc->gtolOff[i];
        if (len <= 0) continue;
        for (i64 l = 0; l < len; l++) px[l] = localX[g[l]];
        lu_solve(pc->lu[i], pc->piv[i], len, px, py);
        double m = 0.0;
        for (i64 l = 0; l < len; l++) if (fabs(py[l]) > m) m = fabs(py[l]);
        for (i64 l = 0; l < len; l++) scale[g[l]] += nsweep * m;
    }
    free(px); free(py);
    return 0;
}

/* one-process PCApply_PATCH (identity star forest): steps 2-7 of ssc/libssc.c:1321-1421.
 * VecReciprocal (:1282, PETSc) leaves zero entries at zero; restated here. */
int oracle_setup_weights(OraclePC *pc, i64 n) {
    free(pc->dof_weights); pc->dof_weights = malloc((n + 1) * sizeof(double)); pc->nowned = n;
    oracle_multiplicity_local(pc, n, pc->dof_weights);
    for (i64 i = 0; i < n; i++) if (pc->dof_weights[i] != 0.0) pc->dof_weights[i] = 1.0 / pc->dof_weights[i];
    return 0;
}
int oracle_apply(OraclePC *pc, i64 n, const double *x, double *y, int nthreads) {
    int ierr = oracle_apply_local(pc, n, x, y, nthreads);                                       /* identity SF: localX == x, y == localY */
    if (ierr) return ierr;
    for (i64 i = 0; i < pc->numGlobalBcs; i++) { i64 idx = pc->globalBcNodes[i]; if (idx < n) y[idx] = x[idx]; }   /* :1408-1413 */
    if (pc->partition_of_unity) {                                                               /* :1418-1421 */
        if (!pc->dof_weights || pc->nowned != n) oracle_setup_weights(pc, n);
        for (i64 i = 0; i < n; i++) y[i] *= pc->dof_weights[i];
    }
    return 0;
}
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Sequential multiplicative Schwarz over the patches in the given order (N4; the reference has no such code path any
 * more, ssc/libssc.c:1568-1572, so this defines the semantics the GPU's coloured sweep is checked against):
 *   y = 0;  for p in order:  r_p = (x - A y)[gtol_p];  y[gtol_p] += B_p^{-1} r_p;   then Dirichlet pass-through. */
int oracle_apply_multiplicative(OraclePC *pc, i64 n, const double *x, double *y, i64 norder, const i64 *order) {
    if (!pc->lu) { snprintf(pc->err, 256, "operators not set up"); return 73; }
    i64 maxn = 0; for (i64 i = 0; i < pc->npatch; i++) if (pc->gtolDof[i] > maxn) maxn = pc->gtolDof[i];
    double *px = malloc((maxn + 1) * sizeof(double)), *py = malloc((maxn + 1) * sizeof(double));
    memset(y, 0, n * sizeof(double));
    for (i64 o = 0; o < norder; o++) {
        i64 i = order[o], len = pc->gtolDof[i]; const i64 *g = pc->gtol + pc->gtolOff[i];
        if (len <= 0) continue;
        for (i64 l = 0; l < len; l++) {
            double s = 0.0;
            for (i64 k = pc->rowptr[g[l]]; k < pc->rowptr[g[l] + 1]; k++) s += pc->vals[k] * y[pc->colidx ? pc->colidx[k] : (i64)pc->colidx32[k]];
            px[l] = x[g[l]] - s;
        }
        lu_solve(pc->lu[i], pc->piv[i], len, px, py);
        for (i64 l = 0; l < len; l++) y[g[l]] += py[l];
    }
    for (i64 i = 0; i < pc->numGlobalBcs; i++) { i64 idx = pc->globalBcNodes[i]; if (idx < n) y[idx] = x[idx]; }
    free(px); free(py);
    return 0;
}
