/* Plain-C use of the C ABI (include/ssc_b200.h): the language of the code this library replaces (ssc/libssc.c).
 *
 * 1-D Laplacian tridiag(-1, 2, -1) on N dofs, overlapping patches of three consecutive dofs handed over as index sets
 * (SSCPatchSetPatches, what PCPatchCreateCellPatchDiscretisationInfo ssc/libssc.c:598-900 would have produced).  Every
 * patch operator is the 3x3 matrix [[2,-1,0],[-1,2,-1],[0,-1,2]] (2x2 at the ends), so PCApply has a closed form the
 * program checks itself against.
 *
 *   gcc -std=c99 -Iinclude examples/c_api_demo.c -Lssc_b200 -lssc_b200 -Wl,-rpath,$PWD/ssc_b200 -o c_api_demo
 *
 * Exit code 0: result verified on the GPU.  Exit code 3: no CUDA device (the library has no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "ssc_b200.h"

#define N 64

static int check(int ierr, const char *what)
{
    if (ierr) fprintf(stderr, "%s failed with code %d: %s\n", what, ierr, SSCGetLastError());
    return ierr;
}

int main(void)
{
    SSCPatchPC pc = NULL;
    int ndev = 0;
    if (SSCDeviceCount(&ndev) || ndev < 1) {
        fprintf(stderr, "no CUDA device: %s\n", SSCGetLastError());
        return 3;
    }
    if (check(SSCPatchCreate(0, &pc), "SSCPatchCreate")) return 1;
    /* CSR of the Laplacian */
    SSCInt rowptr[N + 1];
    int32_t colidx[3 * N];
    double vals[3 * N];
    SSCInt nnz = 0;
    for (int i = 0; i < N; ++i) {
        rowptr[i] = nnz;
        if (i > 0) { colidx[nnz] = i - 1; vals[nnz++] = -1.0; }
        colidx[nnz] = i; vals[nnz++] = 2.0;
        if (i < N - 1) { colidx[nnz] = i + 1; vals[nnz++] = -1.0; }
    }
    rowptr[N] = nnz;
    /* patch p = {p-1, p, p+1} clipped to [0, N) */
    SSCInt off[N + 1], gtol[3 * N], ng = 0;
    for (int p = 0; p < N; ++p) {
        off[p] = ng;
        for (int d = p - 1; d <= p + 1; ++d) if (d >= 0 && d < N) gtol[ng++] = d;
    }
    off[N] = ng;
    if (check(SSCPatchSetPatches(pc, N, off, gtol, N, N, 0, NULL), "SSCPatchSetPatches")) return 1;
    if (check(SSCPatchSetOption(pc, "sub_pc_type", "cholesky"), "SSCPatchSetOption")) return 1;
    if (check(SSCPatchSetOperatorCSR(pc, N, rowptr, colidx, vals, SSC_MEM_HOST), "SSCPatchSetOperatorCSR")) return 1;
    if (check(SSCPatchSetUp(pc), "SSCPatchSetUp")) return 1;
    double x[N], y[N], ref[N];
    for (int i = 0; i < N; ++i) { x[i] = sin(0.37 * i) + 0.1 * i; ref[i] = 0.0; }
    if (check(SSCPatchApply(pc, x, y, SSC_MEM_HOST), "SSCPatchApply")) return 1;
    /* closed form: inverse of the 3x3 block = 1/4 [[3,2,1],[2,4,2],[1,2,3]], of the 2x2 end blocks = 1/3 [[2,1],[1,2]] */
    for (int p = 0; p < N; ++p) {
        if (p == 0 || p == N - 1) {
            const int a = p == 0 ? 0 : N - 2, b = a + 1;
            ref[a] += (2.0 * x[a] + x[b]) / 3.0;
            ref[b] += (x[a] + 2.0 * x[b]) / 3.0;
        } else {
            const double u = x[p - 1], v = x[p], w = x[p + 1];
            ref[p - 1] += (3.0 * u + 2.0 * v + w) / 4.0;
            ref[p] += (2.0 * u + 4.0 * v + 2.0 * w) / 4.0;
            ref[p + 1] += (u + 2.0 * v + 3.0 * w) / 4.0;
        }
    }
    double err = 0.0, nrm = 0.0;
    for (int i = 0; i < N; ++i) { err += (y[i] - ref[i]) * (y[i] - ref[i]); nrm += ref[i] * ref[i]; }
    err = sqrt(err / nrm);
    printf("c_api_demo: %d dofs, %d patches, relative error vs closed form %.2e\n", N, N, err);
    SSCPatchDestroy(&pc);
    return err < 1e-13 ? 0 : 2;
}
