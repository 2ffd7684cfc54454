// Batched in-place inversion of the patch blocks (K2): the Gauss-Jordan family.
#pragma once
#include "kernels_common.cuh"

// ---------------------------------------------------------------------------------------------
// K2 factor: in-place Gauss-Jordan inversion, one CTA per patch.  With kPivot the pivot of step k is the
// largest entry of column k among rows >= k (partial pivoting: -sub_pc_type lu); without it the diagonal is
// used (SPD blocks: -sub_pc_type cholesky).  The block is loaded TRANSPOSED, so what is inverted is B^T and
// what is stored back is S = (B^T)^-1 = (B^-1)^T, the layout the apply kernel streams.
// kShared: the working copy lives in shared memory (n <= 168); otherwise the block is inverted in place in
// global memory (L2-resident) and transposed afterwards.
// ---------------------------------------------------------------------------------------------
template <bool kPivot>
__device__ __forceinline__ int gauss_jordan(double *A, int n, int ld, double *rowv, double *colv, int *piv,
                                            double *redv, int *redi)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    // thread -> fixed column, strided rows: the pivot-row entry stays in a register, the multiplier is a broadcast
    const int ncol_threads = n < (int)blockDim.x ? n : (int)blockDim.x;
    const int rgroups = blockDim.x / ncol_threads;          // >= 1
    const int mycol = tid % ncol_threads, myrg = tid / ncol_threads;
    int fail = 0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        if (kPivot) {
            double best = -1.0; int bi = k;
            for (int i = k + tid; i < n; i += blockDim.x) {
                const double v = fabs(A[(ll)i * ld + k]);
                if (v > best) { best = v; bi = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { redv[warp] = best; redi[warp] = bi; }
            __syncthreads();
            if (warp == 0) {
                best = lane < nwarps ? redv[lane] : -1.0; bi = lane < nwarps ? redi[lane] : n;
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_down_sync(0xffffffffu, best, o);
                    const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                if (lane == 0) { redi[32] = bi; piv[k] = bi; }
            }
            __syncthreads();
            p = redi[32];
            if (p != k) {
                for (int j = tid; j < n; j += blockDim.x) {
                    const double a = A[(ll)k * ld + j];
                    A[(ll)k * ld + j] = A[(ll)p * ld + j];
                    A[(ll)p * ld + j] = a;
                }
                __syncthreads();
            }
        }
        const double d = A[(ll)k * ld + k];
        if (!(fabs(d) > 0.0) || (!kPivot && !(d > 0.0))) { fail = 1; break; }   // uniform across the CTA
        const double dinv = 1.0 / d;
        __syncthreads();
        for (int j = tid; j < n; j += blockDim.x) {
            rowv[j] = A[(ll)k * ld + j] * dinv;
            colv[j] = A[(ll)j * ld + k];
        }
        __syncthreads();
        for (int j = mycol; j < n && myrg < rgroups; j += ncol_threads) {
            const double r = rowv[j];
            if (j == k) {
                for (int i = myrg; i < n; i += rgroups) A[(ll)i * ld + j] = (i == k) ? dinv : -colv[i] * dinv;
            } else {
                for (int i = myrg; i < n; i += rgroups) {
                    if (i == k) A[(ll)i * ld + j] = r;
                    else A[(ll)i * ld + j] = fma(-colv[i], r, A[(ll)i * ld + j]);
                }
            }
        }
        __syncthreads();
    }
    if (kPivot && !fail) {
        // (P M)^-1 = M^-1 P^T: undo the row exchanges as column exchanges, last first
        for (int k = n - 1; k >= 0; --k) {
            const int p = piv[k];
            if (p != k) {
                for (int i = tid; i < n; i += blockDim.x) {
                    const double a = A[(ll)i * ld + k];
                    A[(ll)i * ld + k] = A[(ll)i * ld + p];
                    A[(ll)i * ld + p] = a;
                }
                __syncthreads();
            }
        }
    }
    return fail;
}

template <bool kPivot, bool kShared>
__global__ void __launch_bounds__(1024)
ssc_invert_kernel(double *__restrict__ blocks, int n, int ldg, ll stride, ll count, int *__restrict__ fail_flags)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = kShared ? (n | 1) : ldg;
    double *rowv = reinterpret_cast<double *>(smem_raw);
    double *colv = rowv + n;
    double *redv = colv + n;
    int *redi = reinterpret_cast<int *>(redv + 32);
    int *piv = redi + 34;
    double *As = reinterpret_cast<double *>(smem_raw) + 2 * n + 32 + ((34 + n + 1) / 2 + 1);
    const int tid = threadIdx.x;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        double *B = blocks + p * stride;
        double *A = kShared ? As : B;
        if (kShared) {
            for (int e = tid; e < n * n; e += blockDim.x) { const int i = e / n, j = e - i * n; As[(ll)j * ld + i] = B[(ll)i * ldg + j]; }
        }
        __syncthreads();
        const int fail = gauss_jordan<kPivot>(A, n, ld, rowv, colv, piv, redv, redi);
        __syncthreads();
        if (tid == 0) fail_flags[p] = fail;
        if (kShared) {
            for (int e = tid; e < n * n; e += blockDim.x) { const int i = e / n, j = e - i * n; B[(ll)i * ldg + j] = As[(ll)i * ld + j]; }
        } else {
            // in-place transpose of the n x n block
            for (ll e = tid; e < (ll)n * n; e += blockDim.x) {
                const int i = (int)(e / n), j = (int)(e - (ll)i * n);
                if (i < j) { const double a = B[(ll)i * ldg + j]; B[(ll)i * ldg + j] = B[(ll)j * ldg + i]; B[(ll)j * ldg + i] = a; }
            }
        }
        __syncthreads();
    }
}

// Register-resident Gauss-Jordan for n <= 128: the whole block lives in registers, thread (tr, tc) owns the
// 8 x 4 tile of rows 8tr.. and columns 4tc..; per elimination step a thread reads only its 8 multipliers and its 4
// pivot-row entries from shared memory (32 FMAs per 12 loads, against 1 FMA per 3 shared accesses in the
// shared-memory kernel above).  Pivoting is implicit: rows are never moved, every warp finds the pivot row of the
// step redundantly (same deterministic arg-max), and the permutation is applied once when the result is stored.
// Two barriers per step; the two broadcast vectors are double-buffered so no third barrier is needed.
__device__ __forceinline__ double sel4(const double (&v)[4], int c)
{
    return c == 0 ? v[0] : (c == 1 ? v[1] : (c == 2 ? v[2] : v[3]));
}

template <bool kPivot>
__global__ void __launch_bounds__(512, 1)
ssc_invert_reg_kernel(double *__restrict__ blocks, int n, int ldg, ll stride, ll count, int *__restrict__ fail_flags, int TC)
{
    __shared__ __align__(16) double colv[2][128];
    __shared__ __align__(16) double rowv[2][128];
    __shared__ int perm[128], qinv[128];
    const int tid = threadIdx.x, lane = tid & 31;
    const int tr = tid / TC, tc = tid - tr * TC;
    const bool owner = tr < (n + 7) / 8;
    // non-owner threads (padding of the thread grid) read row/column slots 0.. harmlessly
    const int r0 = owner ? 8 * tr : 0, c0 = 4 * tc;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        double *B = blocks + p * stride;
        double a[8][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = c0 + c;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = r0 + r;
                a[r][c] = (owner && i < n && j < n) ? B[(ll)j * ldg + i] : 0.0;      // transposed load: A = B^T
            }
        }
        unsigned usedmask = 0;      // bit s of lane l <-> row l + 32 s has been a pivot row
        int fail = 0;
        for (int k = 0; k < n; ++k) {
            const int buf = k & 1;
            // (1) the owners of column k publish it and clear it in their registers (k & 3 is uniform: no divergence
            //     on the switch, static register indices inside)
            if (owner && tc == (k >> 2)) {
#define PUBLISH_COL(C)                                                                  \
    _Pragma("unroll") for (int r = 0; r < 8; ++r) { colv[buf][r0 + r] = a[r][C]; a[r][C] = 0.0; }
                switch (k & 3) {
                case 0: PUBLISH_COL(0) break;
                case 1: PUBLISH_COL(1) break;
                case 2: PUBLISH_COL(2) break;
                default: PUBLISH_COL(3) break;
                }
#undef PUBLISH_COL
            }
            __syncthreads();
            int prow = k;
            if (kPivot) {
                double best = -1.0;
                int bi = 1 << 30;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int i = lane + 32 * s;
                    if (i < n && !((usedmask >> s) & 1u)) {
                        const double v = fabs(colv[buf][i]);
                        if (v > best) { best = v; bi = i; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                prow = bi;
                if (prow >= n) { fail = 1; break; }                  // no candidate (NaNs)
                if ((prow & 31) == lane) usedmask |= 1u << (prow >> 5);
                if (tid == 0) perm[k] = prow;
            }
            const double d = colv[buf][prow];
            if (!(fabs(d) > 0.0) || (!kPivot && !(d > 0.0))) { fail = 1; break; }     // uniform across the CTA
            const double dinv = 1.0 / d;
            // (2) the owners of the pivot row publish it scaled (slot k carries 1/d) and clear it
            if (owner && tr == (prow >> 3)) {
#define PUBLISH_ROW(R)                                                                                          \
    _Pragma("unroll") for (int c = 0; c < 4; ++c) { rowv[buf][c0 + c] = (c0 + c == k) ? dinv : a[R][c] * dinv; a[R][c] = 0.0; }
                switch (prow & 7) {
                case 0: PUBLISH_ROW(0) break;
                case 1: PUBLISH_ROW(1) break;
                case 2: PUBLISH_ROW(2) break;
                case 3: PUBLISH_ROW(3) break;
                case 4: PUBLISH_ROW(4) break;
                case 5: PUBLISH_ROW(5) break;
                case 6: PUBLISH_ROW(6) break;
                default: PUBLISH_ROW(7) break;
                }
#undef PUBLISH_ROW
            }
            __syncthreads();
            // (3) rank-1 update of the whole tile.  Column k and row p were cleared above, the multiplier of row p is
            //     -1 and slot k of the pivot row is 1/d, so the same FMA yields the Gauss-Jordan special cases:
            //     a[p][j] = rowv[j],  a[i][k] = -colv[i]/d,  a[p][k] = 1/d.
            const double2 *cp = reinterpret_cast<const double2 *>(&colv[buf][r0]);
            const double2 *rp2 = reinterpret_cast<const double2 *>(&rowv[buf][c0]);
            double cv[8], rv[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) { const double2 t = cp[h]; cv[2 * h] = t.x; cv[2 * h + 1] = t.y; }
#pragma unroll
            for (int h = 0; h < 2; ++h) { const double2 t = rp2[h]; rv[2 * h] = t.x; rv[2 * h + 1] = t.y; }
            const int rp = prow - r0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double m = (r == rp) ? 1.0 : -cv[r];
#pragma unroll
                for (int c = 0; c < 4; ++c) a[r][c] = fma(m, rv[c], a[r][c]);
            }
        }
        __syncthreads();
        if (tid < n) { if (!kPivot) perm[tid] = tid; }
        __syncthreads();
        if (tid < n && !fail) qinv[perm[tid]] = tid;       // after an early failure perm[k+1..] was never written
        __syncthreads();
        if (tid == 0) fail_flags[p] = fail;
        if (!fail && owner) {
            // A^-1[kk][cc] = W[perm[kk]][qinv[cc]]  <=>  W[i][j] goes to row qinv[i], column perm[j]
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = r0 + r;
                if (i < n) {
                    const ll row = (ll)qinv[i] * ldg;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = c0 + c;
                        if (j < n) B[row + perm[j]] = a[r][c];
                    }
                }
            }
        }
        __syncthreads();
    }
}

// Blocked Gauss-Jordan for n > 128: the block stays in global memory (L2), 16 columns at a time are brought into
// shared memory and eliminated there (with partial pivoting, rows exchanged across the whole block), and the
// transformation they define is applied to all other columns as one (n x 16) x (16 x n) product on the fp64 tensor
// cores (mma.sync.m8n8k4.f64, DMMA): every element of the block is read and written once per 16 pivots instead of once
// per pivot, which is what bounds the unblocked global-memory kernel above.
//   panel P (n x 16) after its own elimination holds T's non-trivial columns:  rest <- rest + (P - E) * rest[mid rows]
__device__ __forceinline__ void dmma8x8x4(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <bool kPivot, int NB>
__global__ void __launch_bounds__(512, NB == 16 ? 2 : 1)
ssc_invert_blocked_kernel(double *__restrict__ blocks, int n, int ld, ll stride, ll count, int *__restrict__ fail_flags)
{
    constexpr int PP = NB + 1;                            // odd panel row stride: a thread per row walks it without bank conflicts
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int npad = (n + 7) & ~7;
    const int WP = ((npad + 15) & ~15) + 4;               // row stride of W, = 4 mod 16: conflict-free B fragments per half-warp
    double *P = reinterpret_cast<double *>(smem_raw);     // npad x PP
    double *W = P + (size_t)npad * PP;                    // NB x WP
    double *rowv = W + (size_t)NB * WP;                   // NB
    double *redv = rowv + NB;                             // 32
    int *redi = reinterpret_cast<int *>(redv + 32);       // 34
    int *piv = redi + 34;                                 // n
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        double *A = blocks + p * stride;
        int fail = 0;
        for (int kb = 0; kb < n && !fail; kb += NB) {
            const int bw = min(NB, n - kb);
            // (a) panel -> shared memory
            for (int e = tid; e < npad * NB; e += blockDim.x) {
                const int i = e / NB, j = e - i * NB;
                P[i * PP + j] = (i < n && j < bw) ? A[(ll)i * ld + kb + j] : 0.0;
            }
            __syncthreads();
            // (b) unblocked Gauss-Jordan on the panel columns, two barriers per pivot: the scaled pivot row is staged in
            //     rowv, then every thread updates its rows; the row exchange is implicit (the thread that owns position
            //     prow writes both exchanged rows) and the arg-max partials for the next column are collected on the fly.
            auto warp_argmax = [&](double &best, int &bi) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
            };
            if (kPivot) {
                double best = -1.0; int bi = n;
                for (int i = kb + tid; i < n; i += blockDim.x) { const double v = fabs(P[i * PP]); if (v > best) { best = v; bi = i; } }
                warp_argmax(best, bi);
                if (lane == 0) { redv[warp] = best; redi[warp] = bi; }
                __syncthreads();
            }
            for (int j = 0; j < bw; ++j) {
                const int c = kb + j;
                int prow = c;
                if (kPivot) {
                    double best = -1.0; int bi = n;
                    for (int w = 0; w < nwarps; ++w) { const double ov = redv[w]; const int oi = redi[w]; if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; } }
                    prow = bi;
                    if (prow >= n) { fail = 1; break; }
                }
                const double d = P[prow * PP + j];
                if (!(fabs(d) > 0.0) || (!kPivot && !(d > 0.0))) { fail = 1; break; }     // uniform
                const double dinv = 1.0 / d;
                if (tid < NB) rowv[tid] = (tid == j) ? dinv : P[prow * PP + tid] * dinv;
                if (tid == 0) piv[c] = prow;
                __syncthreads();
                double best = -1.0; int bi = n;
                for (int i = tid; i < n; i += blockDim.x) {
                    if (i == c && prow != c) continue;             // written by the owner of position prow
                    if (i == prow) {
                        double old[NB];
                        if (prow != c) {
#pragma unroll
                            for (int jj = 0; jj < NB; ++jj) old[jj] = P[c * PP + jj];
                        }
#pragma unroll
                        for (int jj = 0; jj < NB; ++jj) P[c * PP + jj] = rowv[jj];
                        if (prow != c) {
                            const double f = old[j];
                            double *r = P + prow * PP;
#pragma unroll
                            for (int jj = 0; jj < NB; ++jj) {
                                const double v = (jj == j) ? -f * dinv : fma(-f, rowv[jj], old[jj]);
                                r[jj] = v;
                                if (jj == j + 1) { const double av = fabs(v); if (av > best) { best = av; bi = prow; } }
                            }
                        }
                    } else {
                        double *r = P + i * PP;
                        const double f = r[j];
#pragma unroll
                        for (int jj = 0; jj < NB; ++jj) {
                            const double v = (jj == j) ? -f * dinv : fma(-f, rowv[jj], r[jj]);
                            r[jj] = v;
                            if (jj == j + 1 && i > c) { const double av = fabs(v); if (av > best || (av == best && i < bi)) { best = av; bi = i; } }
                        }
                    }
                }
                if (kPivot) {
                    warp_argmax(best, bi);
                    if (lane == 0) { redv[warp] = best; redi[warp] = bi; }
                }
                __syncthreads();
            }
            if (fail) break;
            // (c) the panel's row exchanges on all other columns: a thread owns a column and replays the exchanges in
            //     order for it, so no barrier is needed between them
            if (kPivot) {
                bool any = false;
                for (int j = 0; j < bw; ++j) any |= (piv[kb + j] != kb + j);
                if (any) {
                    for (int col = tid; col < n; col += blockDim.x) {
                        if (col >= kb && col < kb + bw) continue;
                        for (int j = 0; j < bw; ++j) {
                            const int c = kb + j, pr = piv[c];
                            if (pr != c) { const double t = A[(ll)c * ld + col]; A[(ll)c * ld + col] = A[(ll)pr * ld + col]; A[(ll)pr * ld + col] = t; }
                        }
                    }
                    __syncthreads();
                }
            }
            // (d) the pivot rows of all other columns -> shared memory
            for (int e = tid; e < NB * npad; e += blockDim.x) {
                const int r = e / npad, col = e - r * npad;
                W[r * WP + col] = (r < bw && col < n) ? A[(ll)(kb + r) * ld + col] : 0.0;
            }
            __syncthreads();
            // (e) rest <- rest + (P - E) W on the tensor cores.  A warp owns a strip of kCT consecutive 8 x 8 tiles of one
            //     tile row: the kCT accumulator fragments are fetched first (2*kCT independent loads in flight per lane),
            //     then 4 DMMAs per tile with the A fragments shared along the strip.
            constexpr int kCT = 8;
            const int tr_n = npad >> 3;
            const int nstrips = (tr_n + kCT - 1) / kCT;
            for (int t = warp; t < tr_n * nstrips; t += nwarps) {
                const int ti = t / nstrips, ts = t - ti * nstrips;
                const int i0 = ti << 3;
                const bool midrow = (i0 >= kb && i0 < kb + NB);
                const int ri = i0 + (lane >> 2);
                double c0[kCT], c1[kCT];
#pragma unroll
                for (int u = 0; u < kCT; ++u) {
                    const int j0 = (ts * kCT + u) << 3, cj = j0 + 2 * (lane & 3);
                    const bool skip = (j0 >= kb && j0 < kb + NB) || j0 >= npad;
                    c0[u] = 0.0; c1[u] = 0.0;
                    if (!skip && !midrow && ri < n) {
                        if (cj < n) c0[u] = A[(ll)ri * ld + cj];
                        if (cj + 1 < n) c1[u] = A[(ll)ri * ld + cj + 1];
                    }
                }
                double af[NB / 4];
#pragma unroll
                for (int q = 0; q < NB / 4; ++q) af[q] = P[(i0 + (lane >> 2)) * PP + 4 * q + (lane & 3)];
#pragma unroll
                for (int u = 0; u < kCT; ++u) {
                    const int j0 = (ts * kCT + u) << 3;
                    if (j0 >= npad) continue;
#pragma unroll
                    for (int q = 0; q < NB / 4; ++q) {
                        const double bfr = W[(4 * q + (lane & 3)) * WP + j0 + (lane >> 2)];
                        dmma8x8x4(c0[u], c1[u], af[q], bfr, c0[u], c1[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < kCT; ++u) {
                    const int j0 = (ts * kCT + u) << 3, cj = j0 + 2 * (lane & 3);
                    const bool skip = (j0 >= kb && j0 < kb + NB) || j0 >= npad;
                    if (!skip && ri < n) {
                        if (cj < n) A[(ll)ri * ld + cj] = c0[u];
                        if (cj + 1 < n) A[(ll)ri * ld + cj + 1] = c1[u];
                    }
                }
            }
            // (f) panel back
            for (int e = tid; e < n * bw; e += blockDim.x) {
                const int i = e / bw, j = e - i * bw;
                A[(ll)i * ld + kb + j] = P[i * PP + j];
            }
            __syncthreads();
        }
        __syncthreads();
        if (tid == 0) fail_flags[p] = fail;
        // The array now holds M' = (Pi A)^-1 = A^-1 Pi^T with (Pi A)[k] = A[pi(k)].  Wanted: S = (A^-1)^T, whose row pi(k)
        // is column k of M'.  So: transpose in place, then permute rows; a thread owns a column during the row
        // permutation and follows the cycles of pi, so neither step needs barriers inside.
        for (ll e = tid; e < (ll)n * n; e += blockDim.x) {
            const int i = (int)(e / n), j = (int)(e - (ll)i * n);
            if (i < j) { const double a = A[(ll)i * ld + j]; A[(ll)i * ld + j] = A[(ll)j * ld + i]; A[(ll)j * ld + i] = a; }
        }
        if (kPivot && !fail) {
            int *pi = redi + 34 + n;                      // second int array behind piv
            if (tid == 0) {
                for (int i = 0; i < n; ++i) pi[i] = i;
                for (int c = 0; c < n; ++c) { const int pr = piv[c]; const int t = pi[c]; pi[c] = pi[pr]; pi[pr] = t; }
            }
            __syncthreads();                              // also orders the transpose before the permutation
            for (int col = tid; col < n; col += blockDim.x) {
                // new row pi(k) = old row k.  Walk every cycle once, starting from its smallest member.
                for (int s = 0; s < n; ++s) {
                    if (pi[s] == s) continue;
                    bool smallest = true;
                    for (int q = pi[s]; q != s; q = pi[q]) if (q < s) { smallest = false; break; }
                    if (!smallest) continue;
                    double carry = A[(ll)s * ld + col];
                    int k = s;
                    do {
                        const int dst = pi[k];
                        const double nxt = A[(ll)dst * ld + col];
                        A[(ll)dst * ld + col] = carry;
                        carry = nxt;
                        k = dst;
                    } while (k != s);
                }
            }
        }
        __syncthreads();
    }
}
