// MEASURED AND REJECTED (round 1; profiles/README.md): the two-chunk and the persistent ring-buffered forms of the packed-symmetric apply kernel.
// Kept out of the library's translation unit; they need ssc_b200/csrc/kernels_packed.cuh for their helpers.
#pragma once
#include "../../ssc_b200/csrc/kernels_packed.cuh"

// Two-chunk, two-threads-per-row form of the packed apply (the default): the triangle arrives as two bulk copies split
// at a column boundary j1 (about half the bytes each); everything that only needs columns < j1 is computed while the
// second copy is still in flight, and the n terms of a row are shared by two threads.  With
//   y_i = sum_m L[colstart(min(i, m)) + |i - m|] x_m
// the terms available after the first chunk are exactly those with min(i, m) < j1.
__global__ void __launch_bounds__(1024)
ssc_apply_sym2_kernel(const double *__restrict__ packed, ll pstride, unsigned copy_bytes, int j1, unsigned bytes0, const int *__restrict__ gtol, int n,
                      ll count, const double *__restrict__ x, double *__restrict__ y, double scale, const double *__restrict__ mult,
                      const PeerScatter *__restrict__ ps, double *__restrict__ ypatch)
{
    extern __shared__ __align__(128) unsigned char smem_bulk[];     // its own name: the other kernels declare the dynamic array with 16-byte alignment
    double *L = reinterpret_cast<double *>(smem_bulk);
    double *xs = L + (copy_bytes >> 3);
    double *ypart = xs + n;
    __shared__ __align__(8) unsigned long long bar[2];
    const ll p = blockIdx.x;
    const int t = threadIdx.x;
    const int rp = blockDim.x >> 1;                              // row slots; thread (i, q) = (t % rp, t / rp)
    const unsigned bytes1 = copy_bytes - bytes0;
    if (t == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
        const double *src = packed + p * pstride;
        mbar_expect_tx(&bar[0], bytes0);
        bulk_copy_g2s(L, src, bytes0, &bar[0]);
        if (bytes1) {
            mbar_expect_tx(&bar[1], bytes1);
            bulk_copy_g2s(L + (bytes0 >> 3), src + (bytes0 >> 3), bytes1, &bar[1]);
        }
    }
    const int *g = gtol + p * n;
    for (int i = t; i < n; i += blockDim.x) xs[i] = gather_x(x, g[i], ps);
    __syncthreads();
    const int i = t < rp ? t : t - rp, q = t < rp ? 0 : 1;
    const bool valid = i < n;
    double a0 = 0.0, a1 = 0.0;
    mbar_wait(&bar[0], 0);
    if (valid) a0 = sym_terms(L, xs, n, i, q, 2, 0, i < j1 ? n : j1);          // everything with min(i, m) < j1
    if (bytes1) {
        mbar_wait(&bar[1], 0);
        if (valid && i >= j1) a1 = sym_terms(L, xs, n, i, q, 2, j1, n);
    }
    if (q == 1 && valid) ypart[i] = a0 + a1;
    __syncthreads();
    if (q == 0 && valid) {
        const double sc = mult ? scale * mult[p] : scale;
        emit(y, g[i], (a0 + a1 + ypart[i]) * sc, ps, ypatch, p * n + i);
    }
}

// Persistent form of the packed-symmetric apply: one CTA per SM walks its share of the patches with a ring of kStages
// shared-memory buffers.  Thread 0 keeps kStages - 1 bulk copies (one whole packed triangle each) in flight while the
// CTA multiplies the patch that has landed; the residual entries of the next patch are gathered into registers before
// the current patch is computed, so neither the gather nor the scatter latency stalls the stream.
__global__ void __launch_bounds__(1024, 1)
ssc_apply_sym_persistent_kernel(const double *__restrict__ packed, ll pstride, unsigned copy_bytes, int stages, const int *__restrict__ gtol, int n, ll count,
                                const double *__restrict__ x, double *__restrict__ y, double scale, const double *__restrict__ mult,
                                const PeerScatter *__restrict__ ps, double *__restrict__ ypatch)
{
    extern __shared__ __align__(128) unsigned char smem_bulk[];     // its own name: the other kernels declare the dynamic array with 16-byte alignment
    __shared__ __align__(8) unsigned long long bar[8];
    const size_t stage_doubles = copy_bytes >> 3;
    double *ring = reinterpret_cast<double *>(smem_bulk);
    double *xsbuf = ring + stage_doubles * stages;            // 2 x n
    double *ypart = xsbuf + 2 * n;                             // (F - 1) x rp partial sums
    const int rp = (n + 31) & ~31;                             // row slots per part
    const int F = max(1, (int)blockDim.x / rp);                // threads per row
    const int t = threadIdx.x;
    const ll first = blockIdx.x, step = gridDim.x;
    const ll mine = first < count ? (count - first + step - 1) / step : 0;
    if (t == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
        for (int s = 0; s < stages - 1 && s < mine; ++s) {
            mbar_expect_tx(&bar[s], copy_bytes);
            bulk_copy_g2s(ring + stage_doubles * s, packed + (first + s * step) * pstride, copy_bytes, &bar[s]);
        }
    }
    // residual of the first patch
    if (mine > 0) for (int i = t; i < n; i += blockDim.x) xsbuf[i] = gather_x(x, gtol[first * n + i], ps);
    __syncthreads();
    for (ll k = 0; k < mine; ++k) {
        const ll p = first + k * step;
        const int s = (int)(k % stages);
        const unsigned parity = (unsigned)((k / stages) & 1);
        // keep the ring full: the stage freed at the end of the previous iteration takes patch k + stages - 1
        if (t == 0 && k + stages - 1 < mine) {
            const int sn = (int)((k + stages - 1) % stages);
            mbar_expect_tx(&bar[sn], copy_bytes);
            bulk_copy_g2s(ring + stage_doubles * sn, packed + (p + (stages - 1) * step) * pstride, copy_bytes, &bar[sn]);
        }
        const double *xs = xsbuf + (k & 1) * n;
        double *xn = xsbuf + ((k + 1) & 1) * n;
        // gather for the next patch into a register (rows beyond blockDim are fetched after the compute, rare)
        const int *gn = gtol + (p + step) * n;
        double xnext = 0.0;
        const bool have_next = k + 1 < mine;
        if (have_next && t < n) xnext = gather_x(x, gn[t], ps);
        const int *g = gtol + p * n;
        const double *L = ring + stage_doubles * s;
        mbar_wait(&bar[s], parity);
        const double sc = mult ? scale * mult[p] : scale;
        {
            // F threads per row; thread (i, q) adds the terms m = q, q + F, ... of
            //   y_i = sum_m L[colstart(min(i, m)) + |i - m|] x_m        (the symmetric product on the packed triangle)
            const int i = t % rp, q = t / rp;
            double a0 = 0.0, a1 = 0.0;
            if (i < n && q < F) {
                a0 = sym_terms(L, xs, n, i, q, F, 0, n);
                if (q > 0) ypart[(q - 1) * rp + i] = a0;
            }
            __syncthreads();
            if (i < n && q == 0) {
                double s = a0 + a1;
                for (int qq = 1; qq < F; ++qq) s += ypart[(qq - 1) * rp + i];
                emit(y, g[i], s * sc, ps, ypatch, p * n + i);
            }
        }
        if (have_next) {
            if (t < n) xn[t] = xnext;
            for (int i = t + blockDim.x; i < n; i += blockDim.x) xn[i] = gather_x(x, gn[i], ps);
        }
        __syncthreads();          // stage s and xs are free again
    }
}

