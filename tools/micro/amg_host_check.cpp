// Host-only check of the smoothed-aggregation hierarchy of ssc_b200/csrc/coarse_amg.inc: builds the Q1 stiffness matrix of an m^3
// hex mesh (Dirichlet boundary rows = identity), the hierarchy, and runs V(1,1)-preconditioned CG with the same cycle the device runs.
// Build: g++ -O2 -std=c++17 -pthread -o tools/micro/amg_host_check tools/micro/amg_host_check.cpp
// usage: amg_host_check [m]                 Q1 Poisson
//        amg_host_check m elasticity [bs] [rbm]   Q1 linear elasticity, hierarchy with block size bs (3: node aggregates, 1: scalar aggregates);
//                                               rbm: the six rigid-body modes as near-nullspace instead of the componentwise constants
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
typedef long long i64;
enum { SSC_ERR_ARG_INCOMP = 75 };
static int ssc_set_error(int code, const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return code; }
#include "../../ssc_b200/csrc/coarse_amg.inc"

static void spmv(const HostCSR &A, const std::vector<double> &x, std::vector<double> &y)
{
    y.assign((size_t)A.nr, 0.0);
    for (i64 i = 0; i < A.nr; ++i) { double a = 0; for (i64 k = A.rp[i]; k < A.rp[i + 1]; ++k) a += A.v[k] * x[A.ci[k]]; y[i] = a; }
}
static void dense_solve(const HostCSR &A, const std::vector<double> &b, std::vector<double> &x)
{
    const i64 n = A.nr;
    std::vector<double> M((size_t)n * n, 0.0);
    for (i64 i = 0; i < n; ++i) for (i64 k = A.rp[i]; k < A.rp[i + 1]; ++k) M[i * n + A.ci[k]] += A.v[k];
    x = b;
    for (i64 k = 0; k < n; ++k) {
        for (i64 i = k + 1; i < n; ++i) { const double f = M[i * n + k] / M[k * n + k]; if (f == 0) continue; for (i64 j = k; j < n; ++j) M[i * n + j] -= f * M[k * n + j]; x[i] -= f * x[k]; }
    }
    for (i64 i = n - 1; i >= 0; --i) { double s = x[i]; for (i64 j = i + 1; j < n; ++j) s -= M[i * n + j] * x[j]; x[i] = s / M[i * n + i]; }
}
static void vcycle(const std::vector<AmgHostLevel> &L, size_t l, const std::vector<double> &b, std::vector<double> &x)
{
    if (l + 1 == L.size()) { dense_solve(L[l].A, b, x); return; }
    const i64 n = L[l].A.nr;
    x.resize(n);
    for (i64 i = 0; i < n; ++i) x[i] = L[l].winv[i] * b[i];
    std::vector<double> t, bc, xc, pc;
    spmv(L[l].A, x, t);
    for (i64 i = 0; i < n; ++i) t[i] = b[i] - t[i];
    spmv(L[l].R, t, bc);
    vcycle(L, l + 1, bc, xc);
    spmv(L[l].P, xc, pc);
    for (i64 i = 0; i < n; ++i) x[i] += pc[i];
    spmv(L[l].A, x, t);
    for (i64 i = 0; i < n; ++i) x[i] += L[l].winv[i] * (b[i] - t[i]);
}
// Q1 linear elasticity (lambda = mu = 1) on an m^3 hex mesh, all boundary nodes clamped; dof = 3 * node + component
static HostCSR elasticity(int m)
{
    const int nv = m + 1;
    const i64 nn = (i64)nv * nv * nv, n = 3 * nn;
    double Ke[24][24] = {};
    const double gp[2] = {0.5 - 0.5 / sqrt(3.0), 0.5 + 0.5 / sqrt(3.0)};
    for (int gx = 0; gx < 2; ++gx) for (int gy = 0; gy < 2; ++gy) for (int gz = 0; gz < 2; ++gz) {
        const double x[3] = {gp[gx], gp[gy], gp[gz]};
        double dN[8][3];
        for (int a = 0; a < 8; ++a) for (int d = 0; d < 3; ++d) {
            double t = ((a >> d) & 1) ? 1.0 : -1.0;
            for (int e = 0; e < 3; ++e) if (e != d) t *= ((a >> e) & 1) ? x[e] : 1.0 - x[e];
            dN[a][d] = t;
        }
        for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) {
            double gg = 0; for (int d = 0; d < 3; ++d) gg += dN[a][d] * dN[b][d];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
                Ke[3 * a + i][3 * b + j] += 0.125 * (dN[a][i] * dN[b][j] + dN[a][j] * dN[b][i] + (i == j ? gg : 0.0));
        }
    }
    auto node = [&](int i, int j, int k) { return ((i64)i * nv + j) * nv + k; };
    auto bnd = [&](i64 r) { const int i = (int)(r / ((i64)nv * nv)), j = (int)((r / nv) % nv), k = (int)(r % nv); return i == 0 || j == 0 || k == 0 || i == m || j == m || k == m; };
    std::vector<std::vector<std::pair<int, double>>> rows((size_t)n);
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) for (int k = 0; k < m; ++k)
        for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) {
            const i64 ra = node(i + (a & 1), j + ((a >> 1) & 1), k + ((a >> 2) & 1)), rb = node(i + (b & 1), j + ((b >> 1) & 1), k + ((b >> 2) & 1));
            if (bnd(ra) || bnd(rb)) continue;
            for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) rows[(size_t)(3 * ra + c)].push_back({(int)(3 * rb + d), Ke[3 * a + c][3 * b + d]});
        }
    HostCSR A; A.nr = A.nc = n; A.rp.assign((size_t)n + 1, 0);
    for (i64 r = 0; r < n; ++r) {
        if (bnd(r / 3)) { A.ci.push_back((int)r); A.v.push_back(1.0); }
        else {
            std::sort(rows[(size_t)r].begin(), rows[(size_t)r].end());
            for (size_t q = 0; q < rows[(size_t)r].size();) { const int c = rows[(size_t)r][q].first; double v = 0; while (q < rows[(size_t)r].size() && rows[(size_t)r][q].first == c) v += rows[(size_t)r][q++].second; A.ci.push_back(c); A.v.push_back(v); }
        }
        A.rp[(size_t)r + 1] = (i64)A.ci.size();
    }
    return A;
}

static int solve_and_report(HostCSR A, int bs, const char *what, int m, const double *nns = nullptr, int k = 0);

int main(int argc, char **argv)
{
    if (argc > 2 && std::string(argv[2]) == "elasticity") {
        const int m = atoi(argv[1]), bs = argc > 3 ? atoi(argv[3]) : 3;
        if (argc > 4 && std::string(argv[4]) == "rbm") {
            // the six rigid-body modes as the near-nullspace: translations and rotations about the cube's centre
            const int nv = m + 1;
            const i64 nn = (i64)nv * nv * nv, n = 3 * nn;
            std::vector<double> nns((size_t)6 * n, 0.0);
            for (i64 r = 0; r < nn; ++r) {
                const double x = (double)(r / ((i64)nv * nv)) / m - 0.5, y = (double)((r / nv) % nv) / m - 0.5, z = (double)(r % nv) / m - 0.5;
                for (int c = 0; c < 3; ++c) nns[(size_t)c * n + 3 * r + c] = 1.0;
                nns[(size_t)3 * n + 3 * r + 0] = -y; nns[(size_t)3 * n + 3 * r + 1] = x;          // about z
                nns[(size_t)4 * n + 3 * r + 1] = -z; nns[(size_t)4 * n + 3 * r + 2] = y;          // about x
                nns[(size_t)5 * n + 3 * r + 2] = -x; nns[(size_t)5 * n + 3 * r + 0] = z;          // about y
            }
            return solve_and_report(elasticity(m), bs, "elasticity, node aggregates, rigid-body modes", m, nns.data(), 6);
        }
        return solve_and_report(elasticity(m), bs, bs == 3 ? "elasticity, block aggregates" : "elasticity, scalar aggregates", m);
    }
    const int m = argc > 1 ? atoi(argv[1]) : 16, nv = m + 1;
    // Q1 element stiffness on the unit cube scaled by h: entries by node distance class
    auto idx = [&](int i, int j, int k) { return ((i64)i * nv + j) * nv + k; };
    const i64 n = (i64)nv * nv * nv;
    std::vector<std::vector<std::pair<int, double>>> rows((size_t)n);
    double Ke[8][8];
    for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) {
        // exact Q1 stiffness (h = 1): sum over directions of (d/dx part) x (mass part)^2
        double s = 0;
        for (int d = 0; d < 3; ++d) {
            double t = 1;
            for (int e = 0; e < 3; ++e) { const int ba = (a >> e) & 1, bb = (b >> e) & 1; t *= (e == d) ? (ba == bb ? 1.0 : -1.0) : (ba == bb ? 1.0 / 3 : 1.0 / 6); }
            s += t;
        }
        Ke[a][b] = s;
    }
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) for (int k = 0; k < m; ++k)
        for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) {
            const i64 r = idx(i + (a & 1), j + ((a >> 1) & 1), k + ((a >> 2) & 1)), c = idx(i + (b & 1), j + ((b >> 1) & 1), k + ((b >> 2) & 1));
            rows[r].push_back({(int)c, Ke[a][b]});
        }
    HostCSR A; A.nr = A.nc = n; A.rp.assign(n + 1, 0);
    for (i64 r = 0; r < n; ++r) {
        const int i = r / ((i64)nv * nv), j = (r / nv) % nv, k = r % nv;
        const bool bnd = i == 0 || j == 0 || k == 0 || i == m || j == m || k == m;
        std::sort(rows[r].begin(), rows[r].end());
        if (bnd) { A.ci.push_back((int)r); A.v.push_back(1.0); }
        else {
            for (size_t q = 0; q < rows[r].size();) {
                int c = rows[r][q].first; double s = 0;
                while (q < rows[r].size() && rows[r][q].first == c) s += rows[r][q++].second;
                const int ci_ = c / (nv * nv), cj = (c / nv) % nv, ck = c % nv;
                const bool cb = ci_ == 0 || cj == 0 || ck == 0 || ci_ == m || cj == m || ck == m;
                if (!cb) { A.ci.push_back(c); A.v.push_back(s); }          // symmetric elimination of Dirichlet columns
            }
        }
        A.rp[r + 1] = (i64)A.ci.size();
    }
    return solve_and_report(std::move(A), 1, "Poisson", m);
}

static int solve_and_report(HostCSR A, int bs, const char *what, int m, const double *nns, int k)
{
    const i64 n = A.nr;
    std::vector<AmgHostLevel> L;
    if (amg_build_hierarchy(A, 512, 16, L, bs, nns, k)) return 1;          // kAmgCoarsestMax and the level cap of engine_loworder.inc
    for (size_t l = 0; l < L.size(); ++l) printf("level %zu: n %lld nnz %lld (%.1f per row)\n", l, L[l].A.nr, L[l].A.rp[L[l].A.nr], (double)L[l].A.rp[L[l].A.nr] / L[l].A.nr);
    std::vector<double> b(n), x(n, 0.0), r, z, p, q;
    srand(2); for (auto &v : b) v = rand() / (double)RAND_MAX - 0.5;
    r = b; vcycle(L, 0, r, z); p = z;
    double rz = 0, rr0 = 0; for (i64 i = 0; i < n; ++i) { rz += r[i] * z[i]; rr0 += r[i] * r[i]; }
    int it = 0;
    for (; it < 200; ++it) {
        spmv(L[0].A, p, q);
        double pq = 0; for (i64 i = 0; i < n; ++i) pq += p[i] * q[i];
        const double al = rz / pq; double rr = 0;
        for (i64 i = 0; i < n; ++i) { x[i] += al * p[i]; r[i] -= al * q[i]; rr += r[i] * r[i]; }
        if (rr <= 1e-20 * rr0) { ++it; break; }
        vcycle(L, 0, r, z);
        double rz2 = 0; for (i64 i = 0; i < n; ++i) rz2 += r[i] * z[i];
        const double be = rz2 / rz; rz = rz2;
        for (i64 i = 0; i < n; ++i) p[i] = z[i] + be * p[i];
    }
    printf("m = %d (%s): %lld dofs, PCG iterations to 1e-10: %d\n", m, what, n, it);
    return 0;
}
