// The fused apply kernels (K3+K4+K5), general storage.
#pragma once
#include "kernels_common.cuh"

// ---------------------------------------------------------------------------------------------
// K3+K4+K5 apply: gather -> stored-inverse product -> scatter-add, fused (ssc/libssc.c:1359-1386).
// `ppb` patches per CTA, `tp` threads per patch (tp >= n except for very large patches, where a thread
// owns several rows).  The residual entries of a patch are gathered into shared memory (one index load and
// one x load per patch dof), then thread i streams column i of S with evict-first loads and adds its
// result into y with one fp64 reduction (RED.E.ADD.F64 at L2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
ssc_apply_kernel(const double *__restrict__ S, ll stride, int ld, const int *__restrict__ gtol, int n, ll count, int tp, int ppb,
                 const double *__restrict__ x, double *__restrict__ y, double scale, const double *__restrict__ mult,
                 const PeerScatter *__restrict__ ps, double *__restrict__ ypatch)
{
    extern __shared__ __align__(16) double xs_all[];
    const int lp = threadIdx.x / tp, t = threadIdx.x - lp * tp;
    const ll p = (ll)blockIdx.x * ppb + lp;
    const bool active = lp < ppb && p < count;
    double *xs = xs_all + (size_t)lp * n;
    const int *g = gtol + p * n;
    if (active) for (int i = t; i < n; i += tp) xs[i] = gather_x(x, g[i], ps);
    __syncthreads();
    if (!active) return;
    const double sc = mult ? scale * mult[p] : scale;
    for (int i = t; i < n; i += tp) {
        const double *M = S + p * stride + i;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int j = 0;
#pragma unroll 1
        for (; j + 8 <= n; j += 8) {
            const double m0 = __ldcs(M + (ll)(j + 0) * ld), m1 = __ldcs(M + (ll)(j + 1) * ld);
            const double m2 = __ldcs(M + (ll)(j + 2) * ld), m3 = __ldcs(M + (ll)(j + 3) * ld);
            const double m4 = __ldcs(M + (ll)(j + 4) * ld), m5 = __ldcs(M + (ll)(j + 5) * ld);
            const double m6 = __ldcs(M + (ll)(j + 6) * ld), m7 = __ldcs(M + (ll)(j + 7) * ld);
            a0 = fma(m0, xs[j + 0], a0); a1 = fma(m1, xs[j + 1], a1); a2 = fma(m2, xs[j + 2], a2); a3 = fma(m3, xs[j + 3], a3);
            a0 = fma(m4, xs[j + 4], a0); a1 = fma(m5, xs[j + 5], a1); a2 = fma(m6, xs[j + 6], a2); a3 = fma(m7, xs[j + 7], a3);
        }
        for (; j < n; ++j) a0 = fma(__ldcs(M + (ll)j * ld), xs[j], a0);
        emit(y, g[i], ((a0 + a1) + (a2 + a3)) * sc, ps, ypatch, p * n + i);
    }
}

// Software-pipelined variants: the next group of kU column loads is already in flight while the current group is
// consumed, so a thread always has kU..2*kU loads outstanding (the v0 loop drains its loads every iteration).
// kVec = 1: thread i owns row i (LDG.64).  kVec = 2: thread t owns rows 2t and 2t+1 and loads them with one
// 16-byte LDG.128; the row stride ld is even so every such load is aligned.
template <int kVec> struct VecT;
template <> struct VecT<1> { typedef double T; };
template <> struct VecT<2> { typedef double2 T; };
__device__ __forceinline__ void fma_vec(double m, double xv, double &a0, double &) { a0 = fma(m, xv, a0); }
__device__ __forceinline__ void fma_vec(double2 m, double xv, double &a0, double &a1) { a0 = fma(m.x, xv, a0); a1 = fma(m.y, xv, a1); }

template <int kVec, int kU>
__global__ void __launch_bounds__(256)
ssc_apply_pipe_kernel(const double *__restrict__ S, ll stride, int ld, const int *__restrict__ gtol, int n, ll count, int tp, int ppb,
                      const double *__restrict__ x, double *__restrict__ y, double scale, const double *__restrict__ mult,
                      const PeerScatter *__restrict__ ps, double *__restrict__ ypatch)
{
    typedef typename VecT<kVec>::T T;
    extern __shared__ __align__(16) double xs_all[];
    const int lp = threadIdx.x / tp, t = threadIdx.x - lp * tp;
    const ll p = (ll)blockIdx.x * ppb + lp;
    const bool active = lp < ppb && p < count;
    double *xs = xs_all + (size_t)lp * n;
    const int *g = gtol + p * n;
    if (active) for (int i = t; i < n; i += tp) xs[i] = gather_x(x, g[i], ps);
    __syncthreads();
    if (!active) return;
    const double sc = mult ? scale * mult[p] : scale;
    const int ngroups = n / kU;
    const ll ldv = ld / kVec;                         // row stride in units of T
    for (int i = t * kVec; i < n; i += tp * kVec) {
        const T *M = reinterpret_cast<const T *>(S + p * stride + i);
        T A[kU], B[kU];
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        if (ngroups > 0) {
#pragma unroll
            for (int u = 0; u < kU; ++u) A[u] = __ldcs(M + (ll)u * ldv);
        }
        for (int gi = 0; gi < ngroups; gi += 2) {
            const int j = gi * kU;
            const bool hasB = gi + 1 < ngroups, hasA = gi + 2 < ngroups;
            if (hasB) {
#pragma unroll
                for (int u = 0; u < kU; ++u) B[u] = __ldcs(M + (ll)(j + kU + u) * ldv);
            }
#pragma unroll
            for (int u = 0; u < kU; u += 2) {
                fma_vec(A[u], xs[j + u], a0, a1);
                fma_vec(A[u + 1], xs[j + u + 1], b0, b1);
            }
            if (hasA) {
#pragma unroll
                for (int u = 0; u < kU; ++u) A[u] = __ldcs(M + (ll)(j + 2 * kU + u) * ldv);
            }
            if (hasB) {
#pragma unroll
                for (int u = 0; u < kU; u += 2) {
                    fma_vec(B[u], xs[j + kU + u], a0, a1);
                    fma_vec(B[u + 1], xs[j + kU + u + 1], b0, b1);
                }
            }
        }
        for (int j = ngroups * kU; j < n; ++j) fma_vec(__ldcs(M + (ll)j * ldv), xs[j], a0, a1);
        emit(y, g[i], (a0 + b0) * sc, ps, ypatch, p * n + i);
        if (kVec == 2 && i + 1 < n) emit(y, g[i + 1], (a1 + b1) * sc, ps, ypatch, p * n + i + 1);
    }
}

// Persistent form for small and mid-size patches (n <= 96): the groups of a CTA walk the patch list in lockstep, and
// the two dependent gather loads of a patch (its dof index, then x at that index) are issued one and two patches ahead
// of the patch being streamed -- index of patch it + 2 and x of patch it + 1 are in flight while patch `it` streams its
// block -- so a group is never idle in a gather: with one patch per group and CTA (the kernels above) a 27-dof patch
// spends about as long fetching its 27 indices and residual entries as streaming its 5.8 KB block
// (0.67 of the measured HBM bandwidth at Q2 112^3, tools/gpu_apply_sizes.py).  tp >= n: thread t of a group owns row t.
template <int kU>
__global__ void __launch_bounds__(256)
ssc_apply_stream_kernel(const double *__restrict__ S, ll stride, int ld, const int *__restrict__ gtol, int n, ll count, int tp, int ppb,
                        const double *__restrict__ x, double *__restrict__ y, double scale, const double *__restrict__ mult,
                        const PeerScatter *__restrict__ ps, double *__restrict__ ypatch)
{
    extern __shared__ __align__(16) double xs_all[];          // [2][ppb][n]
    const int lp = threadIdx.x / tp, t = threadIdx.x - lp * tp;
    const ll ngr = (ll)gridDim.x * ppb;                        // groups in the grid
    const ll first = (ll)blockIdx.x * ppb + lp;
    const ll iters = (count + ngr - 1) / ngr;                  // the same for every group: the CTA stays in lockstep
    const bool lane_ok = lp < ppb && t < n;
    double *xs0 = xs_all + (size_t)lp * n, *xs1 = xs_all + (size_t)(ppb + lp) * n;
    int g0 = -1, g1 = -1;
    if (lane_ok && first < count) g0 = gtol[first * n + t];
    if (lane_ok && first + ngr < count) g1 = gtol[(first + ngr) * n + t];
    if (lane_ok) xs0[t] = g0 >= 0 ? gather_x(x, g0, ps) : 0.0;
    __syncthreads();
    const int ngroups = n / kU;
    for (ll it = 0; it < iters; ++it) {
        const ll p = first + it * ngr;
        const double *xs = (it & 1) ? xs1 : xs0;
        double *xn = (it & 1) ? xs0 : xs1;
        // in flight while this patch streams: the index two patches ahead, the residual entry one patch ahead
        int g2 = -1;
        if (lane_ok && p + 2 * ngr < count) g2 = gtol[(p + 2 * ngr) * n + t];
        double x1 = 0.0;
        if (g1 >= 0) x1 = gather_x(x, g1, ps);
        if (g0 >= 0) {
            const double *M = S + p * stride + t;
            double A[kU], B[kU];
            double a0 = 0.0, b0 = 0.0;
            if (ngroups > 0) {
#pragma unroll
                for (int u = 0; u < kU; ++u) A[u] = __ldcs(M + (ll)u * ld);
            }
            for (int gi = 0; gi < ngroups; gi += 2) {
                const int j = gi * kU;
                const bool hasB = gi + 1 < ngroups, hasA = gi + 2 < ngroups;
                if (hasB) {
#pragma unroll
                    for (int u = 0; u < kU; ++u) B[u] = __ldcs(M + (ll)(j + kU + u) * ld);
                }
#pragma unroll
                for (int u = 0; u < kU; u += 2) { a0 = fma(A[u], xs[j + u], a0); b0 = fma(A[u + 1], xs[j + u + 1], b0); }
                if (hasA) {
#pragma unroll
                    for (int u = 0; u < kU; ++u) A[u] = __ldcs(M + (ll)(j + 2 * kU + u) * ld);
                }
                if (hasB) {
#pragma unroll
                    for (int u = 0; u < kU; u += 2) { a0 = fma(B[u], xs[j + kU + u], a0); b0 = fma(B[u + 1], xs[j + kU + u + 1], b0); }
                }
            }
            for (int j = ngroups * kU; j < n; ++j) a0 = fma(__ldcs(M + (ll)j * ld), xs[j], a0);
            const double sc = mult ? scale * mult[p] : scale;
            emit(y, g0, (a0 + b0) * sc, ps, ypatch, p * n + t);
        }
        if (lane_ok) xn[t] = x1;
        g0 = g1; g1 = g2;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// -ssc_sub_solve refine: the stored-inverse product followed by ONE step of iterative refinement against the kept patch
// operator, the residual accumulated in double-double (error-free product through FMA, two-sum):
//     y = S x;   then `steps` times:  r = x - B y  (twice the working precision);   y += S r.
// The reference solves each patch problem with LU factors (KSPSolve of a preonly + lu sub-KSP, ssc/libssc.c:1374); a
// stored inverse loses kappa(B) * eps, which one refinement step with an accurately computed residual gives back
// (every step contracts the error by ||I - S B||; result accurate to a few eps as long as that is << 1), at 1 + 2 steps
// passes over n^2 doubles instead of one.
// One CTA per patch, thread i owns row i in the two S products (coalesced columns of S = (B^-1)^T); for the residual a
// warp owns a row of B at a time (B is kept row-major as extracted) and reduces its 32 partial double-double sums.
// Algorithmic bytes: 8 (1 + 2 steps) n^2 per patch in place of 8 n^2.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dd_add_product(double &s, double &e, double a, double b)
{
    const double p = a * b, pe = fma(a, b, -p);
    const double t = s + p, z = t - s;
    e += ((s - (t - z)) + (p - z)) + pe;
    s = t;
}
__device__ __forceinline__ void dd_add_dd(double &s, double &e, double s2, double e2)
{
    const double t = s + s2, z = t - s;
    e += ((s - (t - z)) + (s2 - z)) + e2;
    s = t;
}

__global__ void __launch_bounds__(1024)
ssc_apply_refine_kernel(const double *__restrict__ S, const double *__restrict__ B0, ll stride, int ld, const int *__restrict__ gtol, int n, ll count,
                        const double *__restrict__ x, double *__restrict__ y, double scale, const double *__restrict__ mult,
                        const PeerScatter *__restrict__ ps, double *__restrict__ ypatch, int steps)
{
    extern __shared__ __align__(16) double sh[];                // xs[n], y0[n], r[n]
    double *xs = sh, *y0 = sh + n, *r = sh + 2 * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        const int *g = gtol + p * n;
        const double *Sp = S + p * stride, *Bp = B0 + p * stride;
        for (int i = tid; i < n; i += blockDim.x) xs[i] = gather_x(x, g[i], ps);
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {
            const double *M = Sp + i;
            double a0 = 0.0, a1 = 0.0;
            int j = 0;
            for (; j + 2 <= n; j += 2) { a0 = fma(M[(ll)j * ld], xs[j], a0); a1 = fma(M[(ll)(j + 1) * ld], xs[j + 1], a1); }
            if (j < n) a0 = fma(M[(ll)j * ld], xs[j], a0);
            y0[i] = a0 + a1;
        }
        __syncthreads();
        const double sc = mult ? scale * mult[p] : scale;
        for (int step = 0; step < steps; ++step) {
            for (int i = warp; i < n; i += nwarps) {
                const double *row = Bp + (ll)i * ld;
                double s = 0.0, e = 0.0;
                for (int j = lane; j < n; j += 32) dd_add_product(s, e, -row[j], y0[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double s2 = __shfl_xor_sync(0xffffffffu, s, o), e2 = __shfl_xor_sync(0xffffffffu, e, o);
                    dd_add_dd(s, e, s2, e2);
                }
                if (lane == 0) { dd_add_dd(s, e, xs[i], 0.0); r[i] = s + e; }
            }
            __syncthreads();
            const bool last = step + 1 == steps;
            for (int i = tid; i < n; i += blockDim.x) {
                const double *M = Sp + i;
                double a0 = 0.0, a1 = 0.0;
                int j = 0;
                for (; j + 2 <= n; j += 2) { a0 = fma(M[(ll)j * ld], r[j], a0); a1 = fma(M[(ll)(j + 1) * ld], r[j + 1], a1); }
                if (j < n) a0 = fma(M[(ll)j * ld], r[j], a0);
                const double yi = y0[i] + (a0 + a1);
                if (last) emit(y, g[i], yi * sc, ps, ypatch, p * n + i);
                else y0[i] = yi;                                   // row i of y0 is read only by its owner in this loop
            }
            __syncthreads();
        }
    }
}

// kappa_1(B) = ||B||_1 ||B^-1||_1 of every patch from the kept operator and its stored inverse (-ssc_sub_solve auto
// refines the size classes whose worst patch exceeds a threshold); one CTA per patch, the class maximum through atomicMax
// on the bit pattern (non-negative doubles order like integers).
__global__ void __launch_bounds__(256)
ssc_kappa_kernel(const double *__restrict__ S, const double *__restrict__ B0, ll stride, int ld, int n, ll count, unsigned long long *__restrict__ kmax)
{
    __shared__ double red[2][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        const double *Sp = S + p * stride, *Bp = B0 + p * stride;
        // S = (B^-1)^T: column sums of |B^-1| are row sums of |S|; column sums of |B| run down the rows of B
        double mB = 0.0, mS = 0.0;
        for (int c = tid; c < n; c += blockDim.x) {
            double sb = 0.0, ss = 0.0;
            for (int k = 0; k < n; ++k) { sb += fabs(Bp[(ll)k * ld + c]); ss += fabs(Sp[(ll)k * ld + c]); }
            mB = fmax(mB, sb); mS = fmax(mS, ss);         // ||S||_inf over columns = ||B^-1||_inf; both norms bound kappa alike
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mB = fmax(mB, __shfl_xor_sync(0xffffffffu, mB, o)); mS = fmax(mS, __shfl_xor_sync(0xffffffffu, mS, o)); }
        if (lane == 0) { red[0][warp] = mB; red[1][warp] = mS; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a = fmax(a, red[0][w]); b = fmax(b, red[1][w]); }
            const double kappa = a * b;
            if (kappa == kappa) atomicMax(kmax, (unsigned long long)__double_as_longlong(kappa));
        }
        __syncthreads();
    }
}
