// Synthetic operators for tests and benchmarks (NOT part of the preconditioner): the assembled stiffness
// matrix of inner(grad u, grad v)*dx for tensor-product Q_p Lagrange elements on a structured box, written
// straight to CSR in the caller's node numbering.  On a uniform box the matrix is the Kronecker sum
//   A = sum_a  M_0 x .. x K_a x .. x M_{d-1}
// of the assembled 1-D stiffness (K_a) and mass (M_a) matrices, so a row is the tensor product of the 1-D
// rows' supports.  This stands in for Firedrake's assembly (absent here) at sizes numpy cannot hold
// (3-D Q3 on 64^3 cells: 8.9e8 nonzeros).  Rows and columns of Dirichlet nodes are replaced by the identity.
#include <cstdint>
#include <vector>
#include <algorithm>
#include <utility>
#ifdef _OPENMP
#include <omp.h>
#endif
typedef int64_t i64;

namespace {
struct Grid {
    int d, p;
    i64 N[3];
    inline void range(int a, i64 L, i64 &lo, i64 &hi) const {
        if (L % p == 0) { lo = std::max<i64>(L - p, 0); hi = std::min<i64>(L + p, N[a] - 1); }
        else { lo = (L / p) * p; hi = lo + p; }
    }
    inline void unravel(i64 e, i64 *L) const { for (int a = d - 1; a >= 0; --a) { L[a] = e % N[a]; e /= N[a]; } }
    inline i64 ravel(const i64 *L) const { i64 e = 0; for (int a = 0; a < d; ++a) e = e * N[a] + L[a]; return e; }
};
}

extern "C" int ssc_synth_tensor_rowptr(int d, const i64 *N, int p, const i64 *perm, i64 *rowptr)
{
    Grid g; g.d = d; g.p = p; i64 tot = 1;
    for (int a = 0; a < 3; ++a) g.N[a] = a < d ? N[a] : 1;
    for (int a = 0; a < d; ++a) tot *= N[a];
    rowptr[0] = 0;
#pragma omp parallel for schedule(static)
    for (i64 e = 0; e < tot; ++e) {
        i64 L[3] = {0, 0, 0}; g.unravel(e, L);
        i64 w = 1;
        for (int a = 0; a < d; ++a) { i64 lo, hi; g.range(a, L[a], lo, hi); w *= hi - lo + 1; }
        rowptr[perm[e] + 1] = w;
    }
    for (i64 r = 0; r < tot; ++r) rowptr[r + 1] += rowptr[r];
    return 0;
}

// K[a], M[a]: dense N[a] x N[a] row-major.  isbc: one byte per lattice node (lexicographic), may be null.
extern "C" int ssc_synth_tensor_fill(int d, const i64 *N, int p, const double *const *K, const double *const *M, const i64 *perm,
                                     const unsigned char *isbc, const i64 *rowptr, int32_t *colidx, double *vals)
{
    Grid g; g.d = d; g.p = p; i64 tot = 1;
    for (int a = 0; a < 3; ++a) g.N[a] = a < d ? N[a] : 1;
    for (int a = 0; a < d; ++a) tot *= N[a];
#pragma omp parallel
    {
        std::vector<std::pair<int32_t, double>> row;
#pragma omp for schedule(dynamic, 1024)
        for (i64 e = 0; e < tot; ++e) {
            i64 L[3] = {0, 0, 0}, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}; g.unravel(e, L);
            for (int a = 0; a < d; ++a) g.range(a, L[a], lo[a], hi[a]);
            const bool rbc = isbc && isbc[e];
            row.clear();
            i64 J[3] = {0, 0, 0};
            for (J[0] = lo[0]; J[0] <= hi[0]; ++J[0])
                for (J[1] = lo[1]; J[1] <= hi[1]; ++J[1])
                    for (J[2] = lo[2]; J[2] <= hi[2]; ++J[2]) {
                        const i64 f = g.ravel(J);
                        double v = 0.0;
                        if (rbc || (isbc && isbc[f])) v = (f == e) ? 1.0 : 0.0;
                        else {
                            for (int a = 0; a < d; ++a) {
                                double t = K[a][L[a] * g.N[a] + J[a]];
                                for (int b = 0; b < d; ++b) if (b != a) t *= M[b][L[b] * g.N[b] + J[b]];
                                v += t;
                            }
                        }
                        row.emplace_back((int32_t)perm[f], v);
                    }
            std::sort(row.begin(), row.end(), [](const std::pair<int32_t, double> &x, const std::pair<int32_t, double> &y) { return x.first < y.first; });
            const i64 base = rowptr[perm[e]];
            for (size_t k = 0; k < row.size(); ++k) { colidx[base + (i64)k] = row[k].first; vals[base + (i64)k] = row[k].second; }
        }
    }
    return 0;
}

extern "C" void ssc_synth_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
