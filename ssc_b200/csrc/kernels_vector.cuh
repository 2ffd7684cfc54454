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
