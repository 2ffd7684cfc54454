// Small vector kernels of PCApply_PATCH / PCSetUp_PATCH, deterministic slot sum, SpMV helpers.
#pragma once
#include "kernels_common.cuh"

// ---------------------------------------------------------------------------------------------
// small vector kernels of PCApply_PATCH / PCSetUp_PATCH
// ---------------------------------------------------------------------------------------------
// y[idx] = x[idx] on owned Dirichlet dofs, ssc/libssc.c:1408-1413
__global__ void ssc_bc_kernel(const int *__restrict__ bc, ll nbc, const double *__restrict__ x, double *__restrict__ y)
{
    const ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nbc) { const int d = bc[i]; y[d] = x[d]; }
}
// y *= w, ssc/libssc.c:1418-1421
__global__ void ssc_scale_kernel(ll n, const double *__restrict__ w, double *__restrict__ y)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) y[i] *= w[i];
}
// multiplicity: scatter ones through every patch's dof list, ssc/libssc.c:1263-1273
__global__ void ssc_count_kernel(const int *__restrict__ gtol, ll total, double *__restrict__ m)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (ll)gridDim.x * blockDim.x) atomicAdd(m + gtol[i], 1.0);
}
// VecReciprocal (:1282): zero entries stay zero
__global__ void ssc_reciprocal_kernel(ll n, double *__restrict__ w)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) { const double v = w[i]; w[i] = v != 0.0 ? 1.0 / v : 0.0; }
}
// any non-zero flag? (per-patch failure flags of a -pc_patch_save_operators false apply)
__global__ void ssc_any_kernel(ll n, const int *__restrict__ flags, int *__restrict__ any)
{
    int f = 0;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) f |= flags[i];
    if (__any_sync(0xffffffffu, f) && (threadIdx.x & 31) == 0) *any = 1;
}
// star-forest pieces: dst[di[i]] (=|+=) src[si[i]]
__global__ void ssc_gather_kernel(ll n, const int *__restrict__ si, const double *__restrict__ src, double *__restrict__ dst)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) dst[i] = src[si[i]];
}
__global__ void ssc_scatter_kernel(ll n, const int *__restrict__ di, const double *__restrict__ src, double *__restrict__ dst, int add)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) {
        if (add) atomicAdd(dst + di[i], src[i]); else dst[di[i]] = src[i];
    }
}
__global__ void ssc_copy_indexed_kernel(ll n, const int *__restrict__ di, const int *__restrict__ si, const double *__restrict__ src,
                                        double *__restrict__ dst, int add)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) {
        if (add) atomicAdd(dst + di[i], src[si[i]]); else dst[di[i]] = src[si[i]];
    }
}
__global__ void ssc_fill_kernel(ll n, double *__restrict__ a, double v)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) a[i] = v;
}

// y = A x for a CSR matrix, one warp per row (restriction / prolongation of the low-order PC, ssc/lo.py:187-198)
__global__ void ssc_spmv_kernel(ll nrows, const ll *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
                                const double *__restrict__ x, double *__restrict__ y)
{
    const ll warp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    double s = 0.0;
    for (ll k = rowptr[warp] + lane; k < rowptr[warp + 1]; k += 32) s = fma(vals[k], __ldg(x + colidx[k]), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) y[warp] = s;
}

// ---------------------------------------------------------------------------------------------
// Coarse solve of the low-order PC for coarse spaces too large for a dense factorisation (ssc/lo.py:141-166 hands the P1
// operator to whatever PETSc solver the options name): Jacobi-preconditioned conjugate gradients, device-resident.
// Three kernels per iteration, no host round trip inside: the scalars (r.z, p.Ap, r.r) live in device memory, double-
// buffered by iteration parity b so that a kernel never zeroes an accumulator another kernel of the same iteration reads:
//   k1(b): q = A p;  pq[b] += p.q;                        zeroes rz[b^1], rr[b^1]  (accumulated by k2(b))
//   k2(b): alpha = rz[b]/pq[b];  x += alpha p;  r -= alpha q;  z = r/diag;  rz[b^1] += r.z;  rr[b^1] += r.r
//   k3(b): beta = rz[b^1]/rz[b];  p = z + beta p;         zeroes pq[b^1]
// s = {rz[2], pq[2], rr[2]}.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum_atomic(double v, double *dst)
{
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(dst, v);
    }
    __syncthreads();
}
__global__ void __launch_bounds__(256)
ssc_cg_spmv_kernel(ll n, const ll *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
                   const double *__restrict__ p, double *__restrict__ q, double *__restrict__ s, int b)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { s[b ^ 1] = 0.0; s[4 + (b ^ 1)] = 0.0; }
    const int lane = threadIdx.x & 31;
    const ll nw = ((ll)gridDim.x * blockDim.x) >> 5;
    double part = 0.0;
    for (ll row = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += nw) {
        double a = 0.0;
        for (ll k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) a = fma(vals[k], __ldg(p + colidx[k]), a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
        if (lane == 0) { q[row] = a; part = fma(a, p[row], part); }
    }
    block_sum_atomic(part, s + 2 + b);
}
__global__ void __launch_bounds__(256)
ssc_cg_update_kernel(ll n, const double *__restrict__ dinv, const double *__restrict__ p, const double *__restrict__ q,
                     double *__restrict__ x, double *__restrict__ r, double *__restrict__ z, double *__restrict__ s, int b)
{
    const double pq = s[2 + b];
    const double alpha = pq != 0.0 ? s[b] / pq : 0.0;
    double rz = 0.0, rr = 0.0;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, q[i], r[i]);
        const double zi = ri * dinv[i];
        r[i] = ri; z[i] = zi;
        rz = fma(ri, zi, rz); rr = fma(ri, ri, rr);
    }
    block_sum_atomic(rz, s + (b ^ 1));
    block_sum_atomic(rr, s + 4 + (b ^ 1));
}
__global__ void __launch_bounds__(256)
ssc_cg_direction_kernel(ll n, const double *__restrict__ z, double *__restrict__ p, double *__restrict__ s, int b)
{
    const double rz = s[b];
    const double beta = rz != 0.0 ? s[b ^ 1] / rz : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) s[2 + (b ^ 1)] = 0.0;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}
// start: x = 0, r = rhs, z = r/diag, p = z, rz[0] = r.z, rr[0] = r.r  (s zeroed by the caller)
__global__ void __launch_bounds__(256)
ssc_cg_start_kernel(ll n, const double *__restrict__ dinv, const double *__restrict__ rhs, double *__restrict__ x, double *__restrict__ r,
                    double *__restrict__ z, double *__restrict__ p, double *__restrict__ s)
{
    double rz = 0.0, rr = 0.0;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) {
        const double ri = rhs[i], zi = ri * dinv[i];
        x[i] = 0.0; r[i] = ri; z[i] = zi; p[i] = zi;
        rz = fma(ri, zi, rz); rr = fma(ri, ri, rr);
    }
    block_sum_atomic(rz, s + 0);
    block_sum_atomic(rr, s + 4);
}
// 1 / diagonal of a CSR matrix (a warp per row)
__global__ void ssc_csr_inv_diag_kernel(ll n, const ll *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals, double *__restrict__ dinv)
{
    const ll row = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double d = 0.0;
    for (ll k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) if (colidx[k] == row) d += vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    if (lane == 0) dinv[row] = d != 0.0 ? 1.0 / d : 0.0;
}
// y += a x
__global__ void ssc_axpy_kernel(ll n, double a, const double *__restrict__ x, double *__restrict__ y)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) y[i] = fma(a, x[i], y[i]);
}

// The whole coarse solve in ONE cooperative launch: the three phases of an iteration are separated by grid-wide barriers
// instead of kernel boundaries and the convergence test runs on the device (every thread reads the same r.r after the barrier),
// so a solve costs one launch and no host round trip: the Q1 operator of the 64^3 mesh (274 625 rows, 7.4 M nonzeros, L2-
// resident) takes ~160 iterations.  s = {rz[2], pq[2], rr[2]} as above; its_out receives the iteration count (-1: NaN).
#include <cooperative_groups.h>
__global__ void __launch_bounds__(256)
ssc_cg_persistent_kernel(ll n, const ll *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
                         const double *__restrict__ dinv, const double *__restrict__ rhs, double *__restrict__ x, double *__restrict__ r,
                         double *__restrict__ z, double *__restrict__ p, double *__restrict__ q, double *s, double rtol2, int max_it, int *its_out)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    volatile double *sv = s;
    const ll gtid = (ll)blockIdx.x * blockDim.x + threadIdx.x, gsize = (ll)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const ll gwarp = gtid >> 5, nwarps = gsize >> 5;
    {
        double rz = 0.0, rr = 0.0;
        for (ll i = gtid; i < n; i += gsize) {
            const double ri = rhs[i], zi = ri * dinv[i];
            x[i] = 0.0; r[i] = ri; z[i] = zi; p[i] = zi;
            rz = fma(ri, zi, rz); rr = fma(ri, ri, rr);
        }
        block_sum_atomic(rz, s + 0);
        block_sum_atomic(rr, s + 4);
    }
    grid.sync();
    const double rr0 = sv[4];
    int k = 0;
    if (rr0 > 0.0) {
        const double target = rtol2 * rr0;
        for (; k < max_it; ++k) {
            const int b = k & 1;
            if (gtid == 0) { sv[b ^ 1] = 0.0; sv[4 + (b ^ 1)] = 0.0; }          // accumulated in the second phase of this iteration
            {
                // eight lanes per row (a P1 / Q1 row has 7 .. 27 entries): four rows of a warp in flight at once
                double part = 0.0;
                const int sub = lane & 7;
                for (ll base = 4 * gwarp; base < n; base += 4 * nwarps) {          // warp-uniform trip count: the shuffles below need all 32 lanes
                    const ll row = base + (lane >> 3);
                    double a = 0.0;
                    if (row < n) for (ll e = rowptr[row] + sub; e < rowptr[row + 1]; e += 8) a = fma(vals[e], p[colidx[e]], a);
                    a += __shfl_down_sync(0xffffffffu, a, 4, 8);
                    a += __shfl_down_sync(0xffffffffu, a, 2, 8);
                    a += __shfl_down_sync(0xffffffffu, a, 1, 8);
                    if (sub == 0 && row < n) { q[row] = a; part = fma(a, p[row], part); }
                }
                block_sum_atomic(part, s + 2 + b);
            }
            grid.sync();
            {
                const double pq = sv[2 + b];
                const double alpha = pq != 0.0 ? sv[b] / pq : 0.0;
                double rz = 0.0, rr = 0.0;
                for (ll i = gtid; i < n; i += gsize) {
                    x[i] = fma(alpha, p[i], x[i]);
                    const double ri = fma(-alpha, q[i], r[i]);
                    const double zi = ri * dinv[i];
                    r[i] = ri; z[i] = zi;
                    rz = fma(ri, zi, rz); rr = fma(ri, ri, rr);
                }
                block_sum_atomic(rz, s + (b ^ 1));
                block_sum_atomic(rr, s + 4 + (b ^ 1));
            }
            grid.sync();
            const double rr = sv[4 + (b ^ 1)];
            if (!(rr == rr)) { k = -2; break; }                                  // uniform: every thread reads the same value
            if (rr <= target) { ++k; break; }
            {
                const double rz = sv[b];
                const double beta = rz != 0.0 ? sv[b ^ 1] / rz : 0.0;
                if (gtid == 0) sv[2 + (b ^ 1)] = 0.0;
                for (ll i = gtid; i < n; i += gsize) p[i] = fma(beta, p[i], z[i]);
            }
            grid.sync();
        }
    }
    if (gtid == 0) *its_out = k < 0 ? -1 : k;
}

// ---------------------------------------------------------------------------------------------
// Multigrid cycle and preconditioned CG of the coarse solve with -lo_pc_type gamg (coarse_amg.inc, engine_loworder.inc).
// One kernel shape for every sparse product of the cycle: eight lanes per row, four rows per warp (level operators have
// 24 - 60 entries per row, transfer operators 1 - 30).
//   kAmgProduct   out = A x               restriction  b_c = R t
//   kAmgResidual  out = b - A x
//   kAmgAdd       out = out + A x         prolongation x += P x_c
//   kAmgJacobi    out = x + w (b - A x)   damped Jacobi sweep, out != x
// ---------------------------------------------------------------------------------------------
enum { kAmgProduct = 0, kAmgResidual = 1, kAmgAdd = 2, kAmgJacobi = 3 };
template <int MODE>
__global__ void __launch_bounds__(256)
ssc_amg_row_kernel(ll n, const ll *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
                   const double *__restrict__ x, const double *__restrict__ b, const double *__restrict__ w, double *out)
{
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const ll gwarp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll base = 4 * gwarp; base < n; base += 4 * nwarps) {                 // warp-uniform trip count (full-mask shuffles below)
        const ll row = base + (lane >> 3);
        double a = 0.0;
        if (row < n) for (ll e = rowptr[row] + sub; e < rowptr[row + 1]; e += 8) a = fma(vals[e], __ldg(x + colidx[e]), a);
        a += __shfl_down_sync(0xffffffffu, a, 4, 8);
        a += __shfl_down_sync(0xffffffffu, a, 2, 8);
        a += __shfl_down_sync(0xffffffffu, a, 1, 8);
        if (sub == 0 && row < n) {
            if (MODE == kAmgProduct) out[row] = a;
            else if (MODE == kAmgResidual) out[row] = b[row] - a;
            else if (MODE == kAmgAdd) out[row] += a;
            else out[row] = fma(w[row], b[row] - a, x[row]);
        }
    }
}
// x = w b: a damped Jacobi sweep from a zero initial guess
__global__ void ssc_amg_scale_kernel(ll n, const double *__restrict__ w, const double *__restrict__ b, double *__restrict__ x)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) x[i] = w[i] * b[i];
}
// Preconditioned CG with the preconditioner applied between kernels (z = M^-1 r is a multigrid cycle).  s = {rz, pq, rz_new, rr}:
//   start:      x = 0, r = rhs, rr += r.r                                   (s zeroed by the caller)
//   [cycle]  dot: rz_new += r.z (zeroes pq)   direction(first): p = z
//   product:    rz = rz_new, rz_new = rr = 0;  q = A p;  pq += p.q
//   update:     alpha = rz / pq;  x += alpha p;  r -= alpha q;  rr += r.r     (the host reads rr here)
//   [cycle]  dot: rz_new += r.z (zeroes pq)   direction: beta = rz_new / rz;  p = z + beta p
// A kernel only zeroes or moves scalars that no thread of the same kernel reads.
__global__ void __launch_bounds__(256)
ssc_pcg_start_kernel(ll n, const double *__restrict__ rhs, double *__restrict__ x, double *__restrict__ r, double *__restrict__ s)
{
    double rr = 0.0;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) { const double ri = rhs[i]; x[i] = 0.0; r[i] = ri; rr = fma(ri, ri, rr); }
    block_sum_atomic(rr, s + 3);
}
__global__ void __launch_bounds__(256)
ssc_pcg_product_kernel(ll n, const ll *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
                       const double *__restrict__ p, double *__restrict__ q, double *s)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { s[0] = s[2]; s[2] = 0.0; s[3] = 0.0; }
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const ll gwarp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    double part = 0.0;
    for (ll base = 4 * gwarp; base < n; base += 4 * nwarps) {
        const ll row = base + (lane >> 3);
        double a = 0.0;
        if (row < n) for (ll e = rowptr[row] + sub; e < rowptr[row + 1]; e += 8) a = fma(vals[e], __ldg(p + colidx[e]), a);
        a += __shfl_down_sync(0xffffffffu, a, 4, 8);
        a += __shfl_down_sync(0xffffffffu, a, 2, 8);
        a += __shfl_down_sync(0xffffffffu, a, 1, 8);
        if (sub == 0 && row < n) { q[row] = a; part = fma(a, p[row], part); }
    }
    block_sum_atomic(part, s + 1);
}
__global__ void __launch_bounds__(256)
ssc_pcg_update_kernel(ll n, const double *__restrict__ p, const double *__restrict__ q, double *__restrict__ x, double *__restrict__ r, double *__restrict__ s)
{
    const double pq = s[1];
    const double alpha = pq != 0.0 ? s[0] / pq : 0.0;
    double rr = 0.0;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, q[i], r[i]);
        r[i] = ri;
        rr = fma(ri, ri, rr);
    }
    block_sum_atomic(rr, s + 3);
}
__global__ void __launch_bounds__(256)
ssc_pcg_dot_kernel(ll n, const double *__restrict__ r, const double *__restrict__ z, double *__restrict__ s)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) s[1] = 0.0;
    double rz = 0.0;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) rz = fma(r[i], z[i], rz);
    block_sum_atomic(rz, s + 2);
}
__global__ void __launch_bounds__(256)
ssc_pcg_direction_kernel(ll n, const double *__restrict__ z, double *__restrict__ p, const double *__restrict__ s, int first)
{
    const double rz = s[0];
    const double beta = (first || rz == 0.0) ? 0.0 : s[2] / rz;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) p[i] = first ? z[i] : fma(beta, p[i], z[i]);
}
// x = M b for a small dense row-major matrix (the inverse of the coarsest multigrid level): a warp per row
__global__ void ssc_dense_gemv_kernel(int n, const double *__restrict__ M, const double *__restrict__ b, double *__restrict__ x)
{
    const int row = (int)(((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    double a = 0.0;
    for (int j = lane; j < n; j += 32) a = fma(M[(ll)row * n + j], b[j], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (lane == 0) x[row] = a;
}
