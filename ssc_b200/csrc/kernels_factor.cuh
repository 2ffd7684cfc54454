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
        if (tid < n) qinv[perm[tid]] = tid;
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

// Look-ahead form of the register-resident Gauss-Jordan above.  ncu showed the kernel above issue- and latency-bound:
// 295 warp instructions per warp and elimination step for 32 FMAs, because every warp repeats the pivot search and
// the division, and the whole CTA waits for them twice per step.  Here an extra warp (the "searcher") owns the pivot
// search and the reciprocal.  While the 16 worker warps apply step k to their tiles, the owners of column k+1 update
// that column first, publish it and signal the searcher, which finds the pivot of step k+1 and its reciprocal in
// parallel with the bulk of the update.  Named barriers: 1 = column published (workers arrive, searcher waits),
// 2 = pivot ready (searcher arrives, workers wait), 3 = pivot row published (workers only).
// Tiles are 8 x 5 here (16 x 25 workers + the searcher = 448 threads: the register file holds 128 per thread).
// Same arithmetic, same pivot choice (largest magnitude, lowest row on ties), same stored layout as above.
__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <bool kPivot>
__global__ void __launch_bounds__(448, 1)
ssc_invert_la_kernel(double *__restrict__ blocks, int n, int ldg, ll stride, ll count, int *__restrict__ fail_flags, int TC, int W)
{
    __shared__ __align__(16) double colv[2][136];
    __shared__ __align__(16) double rowv[2][136];      // 26 column tiles of 5 reach slot 129
    __shared__ double s_dinv[2];
    __shared__ int s_prow[2];
    __shared__ int s_fail;
    __shared__ int perm[128], qinv[128];
    const int tid = threadIdx.x, lane = tid & 31;
    const int all = W + 32;
    const bool searcher = tid >= W;
    const int tr = tid / TC, tc = tid - tr * TC;
    const bool owner = !searcher && tr < (n + 7) / 8;
    const int r0 = owner ? 8 * tr : 0, c0 = owner ? 5 * tc : 0;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        double *B = blocks + p * stride;
        double a[8][5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int j = c0 + c;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = r0 + r;
                a[r][c] = (owner && i < n && j < n) ? B[(ll)j * ldg + i] : 0.0;      // transposed load: A = B^T
            }
        }
        if (tid == 0) s_fail = 0;
        if (owner && tc == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) { colv[0][r0 + r] = a[r][0]; a[r][0] = 0.0; }
        }
        __syncthreads();
        if (searcher) {
            unsigned usedmask = 0;      // bit s of lane l <-> row l + 32 s has been a pivot row
            for (int k = 0; k < n; ++k) {
                const int buf = k & 1;
                if (k > 0) nbar_sync(1, all);                        // column k is in colv[buf]
                int prow = k;
                if (kPivot) {
                    unsigned long long bestk = 0ull;
                    int bi = 1 << 30;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int i = lane + 32 * s;
                        if (i < n && !((usedmask >> s) & 1u)) {
                            const double v = fabs(colv[buf][i]);
                            const unsigned long long key = (v == v) ? (unsigned long long)__double_as_longlong(v) : 0ull;   // |v| orders like its bits
                            if (bi == (1 << 30) || key > bestk) { bestk = key; bi = i; }
                        }
                    }
                    const bool have = bi < (1 << 30);
                    const unsigned hi = (unsigned)(bestk >> 32), lo = (unsigned)bestk;
                    const unsigned mh = __reduce_max_sync(0xffffffffu, have ? hi : 0u);
                    const bool in1 = have && hi == mh;
                    const unsigned ml = __reduce_max_sync(0xffffffffu, in1 ? lo : 0u);
                    prow = (int)__reduce_min_sync(0xffffffffu, (in1 && lo == ml) ? (unsigned)bi : 0x7fffffffu);
                }
                bool bad = prow >= n;
                double dinv = 0.0;
                if (!bad) {
                    const double d = colv[buf][prow];
                    bad = !(fabs(d) > 0.0) || (!kPivot && !(d > 0.0));
                    dinv = 1.0 / d;
                    if ((prow & 31) == lane) usedmask |= 1u << (prow >> 5);
                }
                if (lane == 0) {
                    s_prow[buf] = bad ? -1 : prow; s_dinv[buf] = dinv; perm[k] = bad ? k : prow;
                    if (bad) s_fail = 1;
                }
                nbar_arrive(2, all);                                  // orders the stores above for the waiting workers
                if (bad) break;
            }
        } else {
            for (int k = 0; k < n; ++k) {
                const int buf = k & 1;
                nbar_sync(2, all);                                    // the pivot of step k is known
                const int prow = s_prow[buf];
                if (prow < 0) break;                                  // no usable pivot: uniform across the CTA
                const double dinv = s_dinv[buf];
                // the owners of the pivot row publish it scaled (slot k carries 1/d) and clear it; the multiplier of the
                // pivot row becomes -1 so that the update below writes the scaled row back without a special case
                if (owner && tr == (prow >> 3)) {
                    if (tc == 0) colv[buf][prow] = -1.0;
#define PUBLISH_ROW(R)                                                                                          \
    _Pragma("unroll") for (int c = 0; c < 5; ++c) { rowv[buf][c0 + c] = (c0 + c == k) ? dinv : a[R][c] * dinv; a[R][c] = 0.0; }
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
                nbar_sync(3, W);
                const double2 *cp = reinterpret_cast<const double2 *>(&colv[buf][r0]);
                double cv[8], rv[5];
#pragma unroll
                for (int h = 0; h < 4; ++h) { const double2 t = cp[h]; cv[2 * h] = t.x; cv[2 * h + 1] = t.y; }
#pragma unroll
                for (int c = 0; c < 5; ++c) rv[c] = rowv[buf][c0 + c];
                // look-ahead: column k+1 first, published for the searcher and cleared like column k was
                const int k1 = k + 1;
                const int t1 = k1 / 5, w1 = k1 - 5 * t1;
                const bool early = owner && k1 < n && tc == t1;
                if (early) {
#define EARLY_COL(C)                                                                                            \
    _Pragma("unroll") for (int r = 0; r < 8; ++r) { colv[buf ^ 1][r0 + r] = fma(-cv[r], rv[C], a[r][C]); a[r][C] = 0.0; }
                    switch (w1) {
                    case 0: EARLY_COL(0) break;
                    case 1: EARLY_COL(1) break;
                    case 2: EARLY_COL(2) break;
                    case 3: EARLY_COL(3) break;
                    default: EARLY_COL(4) break;
                    }
#undef EARLY_COL
                }
                if (k1 < n) nbar_arrive(1, all);
                const int skip = early ? w1 : -1;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    if (c != skip) {
#pragma unroll
                        for (int r = 0; r < 8; ++r) a[r][c] = fma(-cv[r], rv[c], a[r][c]);
                    }
                }
            }
        }
        __syncthreads();
        const int fail = s_fail;
        if (tid < n) qinv[perm[tid]] = tid;
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
                    for (int c = 0; c < 5; ++c) {
                        const int j = c0 + c;
                        if (j < n) B[row + perm[j]] = a[r][c];
                    }
                }
            }
        }
        __syncthreads();
    }
}

// Blocked form of the look-ahead kernel: panels of 5 columns.  Applying the 5 Gauss-Jordan steps of a panel to the
// whole matrix is  A <- T A  with  T = T_5 ... T_1, and T differs from the identity only in the columns of the 5 pivot
// rows.  So every other column gets the rank-5 update  A[:, j] += (T - I)[:, P] A_old[P, j]  (P = the pivot rows, raw
// values from before the panel), and the panel's own slots receive T[:, P] -- which is exactly what the same in-place
// Gauss-Jordan steps leave behind when they are run on the n x 5 panel alone.  The searcher warp does that small
// factorisation (pivot searches, reciprocals) for the next panel while the worker warps apply the current one:
// 200 FMAs per worker thread between barriers instead of 40, and the serial pivot chain runs beside them.
// A panel is NOT five neighbouring columns: it takes the same local column w of five neighbouring column tiles
// (columns 5 (5 g + q) + w, q = 0..4), so its columns belong to 80 threads instead of 16 and the look-ahead update of
// the next panel -- which sits on the critical path between two panel factorisations -- is 40 FMAs per owner, not 200.
// Columns are therefore eliminated in a different order than in the kernels above (any order is a valid partial-pivot
// elimination; perm[] is indexed by column); the inverse agrees to rounding.
// The searcher keeps its working panel row-major in shared memory (S) and re-reads its rows every step, so the row
// that just served as pivot row can be replaced by one lane without dynamic register indexing (a register-only
// version with selects measured 6 % slower).
// position of matrix row i inside a panel column in shared memory: the 16-byte chunks (row pairs) that the 16 row
// tiles of a warp read together are neighbours, so those reads are conflict-free (a plain row order puts them 64
// bytes apart: 8-way bank conflicts, which also starved the searcher's own shared-memory traffic)
__device__ __forceinline__ int ppos(int i) { return ((((i >> 1) & 3) * 16 + (i >> 3)) << 1) | (i & 1); }

template <bool kPivot, int NS>
__global__ void __launch_bounds__(512, 1)
ssc_invert_panel_kernel(double *__restrict__ blocks, int n, int ldg, ll stride, ll count, int *__restrict__ fail_flags, int TR, int TC, int W, int remap)
{
    constexpr int PB = 5, LDP = 136, SR = 6;
    __shared__ __align__(16) double Pcol[2][PB][LDP];      // panel columns: raw from their owners, then (T - I)[:, P] from the searcher
    __shared__ __align__(16) double Rrow[2][PB][LDP];      // the raw pivot rows of a panel
    __shared__ __align__(16) double S[128][SR];            // the searcher's working panel, one row per matrix row
    __shared__ int s_prow[2][8];
    __shared__ unsigned s_best[2][2][4];                   // two searcher warps: each warp's best candidate of a step (hi, lo, row)
    __shared__ __align__(16) double s_prv[2][SR];          // ... and the pivot row handed over by its lane
    __shared__ int s_pj[32], s_pbw[32], s_np;
    __shared__ int perm[128], qinv[128];
    // remap (13 worker warps, 512 threads): warp 13 is the searcher and warps 5 and 9 -- the other two warps of its SM
    // sub-partition besides worker warp 1 -- stay idle, so the serial pivot chain shares its scheduler and fp64 pipe
    // with one worker warp instead of three; the other sub-partitions run four worker warps each
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the non-empty panels in elimination order: panel j = (g, w) holds the columns 5 (5 g + q) + w < n of the tiles
    // 5 g + q < TC, a prefix in q; the list is the same for every block of the class
    if (threadIdx.x == 0) {
        int cntp = 0;
        for (int j = 0; j < PB * ((TC + PB - 1) / PB); ++j) {
            const int g = j / PB, w = j - PB * g;
            int bw = 0;
            for (int q = 0; q < PB; ++q) bw += (PB * g + q < TC && PB * (PB * g + q) + w < n) ? 1 : 0;
            if (bw) { s_pj[cntp] = j; s_pbw[cntp] = bw; ++cntp; }
        }
        s_np = cntp;
    }
    __syncthreads();
    const int NPE = s_np;
    if (remap && (wid == 5 || wid == 9)) {
        for (ll p = blockIdx.x; p < count; p += gridDim.x) __syncthreads();
        return;
    }
    const bool searcher = remap ? wid == 13 : (int)threadIdx.x >= W;
    const int tid = remap ? ((wid - (wid > 5) - (wid > 9) - (wid > 13)) << 5) + lane : (int)threadIdx.x;   // logical thread: tiles, permutation arrays
    const int all = W + 32 * NS;
    const int tc = tid / TR, tr = tid - tc * TR;
    const bool owner = !searcher && tc < TC;
    const int r0 = owner ? 8 * tr : 0, c0 = owner ? PB * tc : 0;
    const int gt = tc / PB, qt = tc - PB * gt;               // column-tile group of this thread, its panel column there
    const int rb = owner ? 2 * tr : 0;                       // ppos(8 tr + 2 h + e) = 2 tr + 32 h + e: this thread's row pairs inside a panel column
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        double *B = blocks + p * stride;
        if (searcher && NS == 2) {
            // Two searcher warps: lane l of warp sw holds rows l + 32 sw and l + 32 sw + 64 of the panel (2 x 5
            // values), so a step costs each warp half the instructions; the warps agree on the pivot through shared
            // memory (s_best) and the lane that holds the pivot row hands it over (s_prv), one 64-thread barrier each.
            const int sw = ((int)threadIdx.x - W) >> 5;
            const int i0 = lane + 32 * sw, i1 = i0 + 64;
            unsigned usedmask = (i0 >= n ? 1u : 0u) | (i1 >= n ? 2u : 0u);
            for (int it = 0; it < NPE; ++it) {
                const int j = s_pj[it], bw = s_pbw[it];
                const int buf = it & 1, g = j / PB, w = j - PB * g;
                nbar_sync(1, all);                                    // panel j is in Pcol[buf]
                double x[2][PB];
#pragma unroll
                for (int c = 0; c < PB; ++c) {
                    x[0][c] = (i0 < n && c < bw) ? Pcol[buf][c][ppos(i0)] : 0.0;
                    x[1][c] = (i1 < n && c < bw) ? Pcol[buf][c][ppos(i1)] : 0.0;
                }
                int pl[PB];
                bool bad = false;
#pragma unroll
                for (int t = 0; t < PB; ++t) {
                    pl[t] = 0;
                    if (t < bw && !bad) {
                        int prow = PB * (PB * g + t) + w;            // unpivoted: the diagonal entry of this column
                        if (kPivot) {
                            const unsigned h0 = (unsigned)__double2hiint(x[0][t]) & 0x7fffffffu, l0 = (unsigned)__double2loint(x[0][t]);
                            const unsigned h1 = (unsigned)__double2hiint(x[1][t]) & 0x7fffffffu, l1 = (unsigned)__double2loint(x[1][t]);
                            const bool ok0 = !(usedmask & 1u) && (h0 | l0) != 0u, ok1 = !(usedmask & 2u) && (h1 | l1) != 0u;
                            const bool take1 = ok1 && (!ok0 || h1 > h0 || (h1 == h0 && l1 > l0));      // ties keep the lower row
                            const bool have = ok0 || ok1;
                            const unsigned bh = take1 ? h1 : h0, bl = take1 ? l1 : l0;
                            const unsigned bi = take1 ? (unsigned)i1 : (unsigned)i0;
                            const unsigned mh = __reduce_max_sync(0xffffffffu, have ? bh : 0u);
                            const bool in1 = have && bh == mh;
                            const unsigned ml = __reduce_max_sync(0xffffffffu, in1 ? bl : 0u);
                            const unsigned mr = __reduce_min_sync(0xffffffffu, (in1 && bl == ml) ? bi : 0x7fffffffu);
                            if (lane == 0) { s_best[t & 1][sw][0] = mh; s_best[t & 1][sw][1] = ml; s_best[t & 1][sw][2] = mr; }
                            nbar_sync(4, 64);
                            const unsigned ah = s_best[t & 1][0][0], al = s_best[t & 1][0][1], ar = s_best[t & 1][0][2];
                            const unsigned bh2 = s_best[t & 1][1][0], bl2 = s_best[t & 1][1][1], br = s_best[t & 1][1][2];
                            const bool av = ar < 0x7fffffffu, bv = br < 0x7fffffffu;
                            const bool takeb = bv && (!av || bh2 > ah || (bh2 == ah && (bl2 > al || (bl2 == al && br < ar))));
                            prow = (int)(takeb ? br : ar);
                        }
                        if (prow >= n) bad = true;
                        else {
                            const bool mine = (prow & 63) == i0;     // rows i0 and i0 + 64 live in this lane
                            const int slot = prow >> 6;
                            if (mine) {
                                double2 *row = reinterpret_cast<double2 *>(&s_prv[t & 1][0]);
                                row[0] = make_double2(slot ? x[1][0] : x[0][0], slot ? x[1][1] : x[0][1]);
                                row[1] = make_double2(slot ? x[1][2] : x[0][2], slot ? x[1][3] : x[0][3]);
                                s_prv[t & 1][4] = slot ? x[1][4] : x[0][4];
                                usedmask |= 1u << slot;
                            }
                            nbar_sync(4, 64);
                            const double2 *prw = reinterpret_cast<const double2 *>(&s_prv[t & 1][0]);
                            const double2 p01 = prw[0], p23 = prw[1];
                            const double pr[PB] = {p01.x, p01.y, p23.x, p23.y, s_prv[t & 1][4]};
                            const double d = pr[t];
                            if (!(fabs(d) > 0.0) || (!kPivot && !(d > 0.0))) bad = true;
                            const double dinv = __drcp_rn(d);        // correctly rounded, = 1.0 / d
                            double rr[PB];
#pragma unroll
                            for (int c = 0; c < PB; ++c) rr[c] = (c == t) ? dinv : pr[c] * dinv;
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
                                const bool isp = mine && slot == s;
                                const double mlt = -x[s][t];
#pragma unroll
                                for (int c = 0; c < PB; ++c) {
                                    const double u = (c == t) ? mlt * dinv : fma(mlt, rr[c], x[s][c]);
                                    x[s][c] = isp ? rr[c] : u;
                                }
                            }
                            pl[t] = prow;
                        }
                    }
                }
                // (T - I)[:, P]: take the unit entries of the pivot rows out again
#pragma unroll
                for (int c = 0; c < PB; ++c) {
                    if (i0 < n && c < bw) Pcol[buf][c][ppos(i0)] = x[0][c] - ((i0 == pl[c]) ? 1.0 : 0.0);
                    if (i1 < n && c < bw) Pcol[buf][c][ppos(i1)] = x[1][c] - ((i1 == pl[c]) ? 1.0 : 0.0);
                }
                if (sw == 0 && lane == 0) {
#pragma unroll
                    for (int t = 0; t < PB; ++t) { s_prow[buf][t] = bad ? -1 : pl[t]; if (t < bw) perm[PB * (PB * g + t) + w] = bad ? PB * (PB * g + t) + w : pl[t]; }
                }
                nbar_arrive(2, all);                                  // orders the stores above for the waiting workers
                if (bad) break;
            }
        } else if (searcher) {
            unsigned usedmask = 0;      // bit s of lane l <-> row l + 32 s has been a pivot row (rows >= n count as used)
#pragma unroll
            for (int s = 0; s < 4; ++s) if (lane + 32 * s >= n) usedmask |= 1u << s;
            for (int it = 0; it < NPE; ++it) {
                const int j = s_pj[it], bw = s_pbw[it];
                const int buf = it & 1, g = j / PB, w = j - PB * g;
                nbar_sync(1, all);                                    // panel j is in Pcol[buf]
                double x[4][PB];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int i = lane + 32 * s;
#pragma unroll
                    for (int c = 0; c < PB; ++c) x[s][c] = (i < n && c < bw) ? Pcol[buf][c][ppos(i)] : 0.0;
                    double2 *row = reinterpret_cast<double2 *>(&S[i][0]);
                    row[0] = make_double2(x[s][0], x[s][1]); row[1] = make_double2(x[s][2], x[s][3]); S[i][4] = x[s][4];
                }
                __syncwarp();
                int pl[PB];
                bool bad = false;
#pragma unroll
                for (int t = 0; t < PB; ++t) {
                    pl[t] = 0;
                    if (t < bw && !bad) {
                        int prow = PB * (PB * g + t) + w;            // unpivoted: the diagonal entry of this column
                        if (kPivot) {
                            // largest |x|, lowest row on ties; integer compares on the bit pattern (|v| orders like its
                            // bits); exact zeros are no candidates, so a zero column ends with prow >= n (singular)
                            unsigned bh = 0u, bl = 0u;
                            int bi = 1 << 30;
#pragma unroll
                            for (int s = 0; s < 4; ++s) {
                                const unsigned kh = (unsigned)__double2hiint(x[s][t]) & 0x7fffffffu, kl = (unsigned)__double2loint(x[s][t]);
                                const bool better = !((usedmask >> s) & 1u) && (kh > bh || (kh == bh && kl > bl));
                                if (better) { bh = kh; bl = kl; bi = lane + 32 * s; }
                            }
                            const bool have = bi < (1 << 30);
                            const unsigned mh = __reduce_max_sync(0xffffffffu, have ? bh : 0u);
                            const bool in1 = have && bh == mh;
                            const unsigned ml = __reduce_max_sync(0xffffffffu, in1 ? bl : 0u);
                            prow = (int)__reduce_min_sync(0xffffffffu, (in1 && bl == ml) ? (unsigned)bi : 0x7fffffffu);
                        }
                        if (prow >= n) bad = true;
                        else {
                            const double2 *prw = reinterpret_cast<const double2 *>(&S[prow][0]);
                            const double2 p01 = prw[0], p23 = prw[1];
                            const double pr[PB] = {p01.x, p01.y, p23.x, p23.y, S[prow][4]};
                            const double d = pr[t];
                            if (!(fabs(d) > 0.0) || (!kPivot && !(d > 0.0))) bad = true;
                            const double dinv = __drcp_rn(d);        // correctly rounded, = 1.0 / d
                            double rr[PB];
#pragma unroll
                            for (int c = 0; c < PB; ++c) rr[c] = (c == t) ? dinv : pr[c] * dinv;
                            __syncwarp();                            // everybody has read the pivot row
#pragma unroll
                            for (int s = 0; s < 4; ++s) {
                                const int i = lane + 32 * s;
                                const double mlt = -x[s][t];
#pragma unroll
                                for (int c = 0; c < PB; ++c) x[s][c] = (c == t) ? mlt * dinv : fma(mlt, rr[c], x[s][c]);
                                double2 *row = reinterpret_cast<double2 *>(&S[i][0]);
                                row[0] = make_double2(x[s][0], x[s][1]); row[1] = make_double2(x[s][2], x[s][3]); S[i][4] = x[s][4];
                            }
                            // the pivot row itself becomes the scaled row, slot t carries 1/d (its lane stored it above:
                            // the same lane overwrites it here, in program order)
                            if ((prow & 31) == lane) {
                                double2 *row = reinterpret_cast<double2 *>(&S[prow][0]);
                                row[0] = make_double2(rr[0], rr[1]); row[1] = make_double2(rr[2], rr[3]); S[prow][4] = rr[4];
                                usedmask |= 1u << (prow >> 5);
                            }
                            __syncwarp();
                            // re-read the own rows: only the pivot row changed under its owner's feet
#pragma unroll
                            for (int s = 0; s < 4; ++s) {
                                const double2 *row = reinterpret_cast<const double2 *>(&S[lane + 32 * s][0]);
                                const double2 a01 = row[0], a23 = row[1];
                                x[s][0] = a01.x; x[s][1] = a01.y; x[s][2] = a23.x; x[s][3] = a23.y; x[s][4] = S[lane + 32 * s][4];
                            }
                            pl[t] = prow;
                        }
                    }
                }
                // (T - I)[:, P]: take the unit entries of the pivot rows out again
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int i = lane + 32 * s;
#pragma unroll
                    for (int c = 0; c < PB; ++c)
                        if (i < n && c < bw) Pcol[buf][c][ppos(i)] = x[s][c] - ((i == pl[c]) ? 1.0 : 0.0);
                }
                if (lane == 0) {
#pragma unroll
                    for (int t = 0; t < PB; ++t) { s_prow[buf][t] = bad ? -1 : pl[t]; if (t < bw) perm[PB * (PB * g + t) + w] = bad ? PB * (PB * g + t) + w : pl[t]; }
                }
                nbar_arrive(2, all);                                  // orders the stores above for the waiting workers
                if (bad) break;
            }
        } else {
            // the tile lives only in the workers' registers (the searcher's registers hold its panel)
            double a[8][PB];
#pragma unroll
            for (int c = 0; c < PB; ++c) {
                const int j = c0 + c;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = r0 + r;
                    a[r][c] = (owner && i < n && j < n) ? B[(ll)j * ldg + i] : 0.0;      // transposed load: A = B^T
                }
            }
            if (owner && gt == 0) {
#pragma unroll
                for (int h = 0; h < 4; ++h) *reinterpret_cast<double2 *>(&Pcol[0][qt][rb + 32 * h]) = make_double2(a[2 * h][0], a[2 * h + 1][0]);
            }
            nbar_arrive(1, all);
            int fail = 0;
            for (int it = 0; it < NPE; ++it) {
                const int j = s_pj[it], bw = s_pbw[it], buf = it & 1, g = j / PB, w = j - PB * g;
                const bool more = it + 1 < NPE;                        // the next non-empty panel
                const int jn = more ? s_pj[it + 1] : 0, bwn = more ? s_pbw[it + 1] : 0;
                const int gn = jn / PB, wn = jn - PB * gn;
                nbar_sync(2, all);                                    // the pivots of panel j are known
                int pl[PB];
#pragma unroll
                for (int t = 0; t < PB; ++t) pl[t] = s_prow[buf][t];
                if (pl[0] < 0) { fail = 1; break; }                   // no usable pivot: uniform across the CTA
                // the owners of the pivot rows publish them as they are
#pragma unroll
                for (int t = 0; t < PB; ++t) {
                    if (t < bw && owner && tr == (pl[t] >> 3)) {
#define PUBLISH_ROW(R) _Pragma("unroll") for (int c = 0; c < PB; ++c) Rrow[buf][t][c0 + c] = a[R][c];
                        switch (pl[t] & 7) {
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
                }
                nbar_sync(3, W);
                const bool in_panel = owner && gt == g && qt < bw;
                const bool in_next = owner && more && gt == gn && qt < bwn;
                // look-ahead: the next panel's columns first, then they go to the searcher
                if (in_next) {
#define EARLY_COL(C)                                                                                                    \
    _Pragma("unroll") for (int t = 0; t < PB; ++t) {                                                                   \
        if (t < bw) {                                                                                                   \
            const double rvc = Rrow[buf][t][c0 + C];                                                                    \
            _Pragma("unroll") for (int h = 0; h < 4; ++h) {                                                             \
                const double2 q2 = *reinterpret_cast<const double2 *>(&Pcol[buf][t][rb + 32 * h]);                      \
                a[2 * h][C] = fma(q2.x, rvc, a[2 * h][C]); a[2 * h + 1][C] = fma(q2.y, rvc, a[2 * h + 1][C]);           \
            }                                                                                                           \
        }                                                                                                               \
    }                                                                                                                   \
    _Pragma("unroll") for (int h = 0; h < 4; ++h) *reinterpret_cast<double2 *>(&Pcol[buf ^ 1][qt][rb + 32 * h]) = make_double2(a[2 * h][C], a[2 * h + 1][C]);
                    switch (wn) {
                    case 0: EARLY_COL(0) break;
                    case 1: EARLY_COL(1) break;
                    case 2: EARLY_COL(2) break;
                    case 3: EARLY_COL(3) break;
                    default: EARLY_COL(4) break;
                    }
#undef EARLY_COL
                }
                if (more) nbar_arrive(1, all);
                // branch-free: a column that must not be touched (this panel's own column: replaced below; the
                // look-ahead column: done above) gets the multiplier 0, so all nine loads of a rank are issued
                // together and the 40 FMAs follow back to back
                const int skip1 = in_panel ? w : -1, skip2 = in_next ? wn : -1;
                bool keep[PB];
#pragma unroll
                for (int c = 0; c < PB; ++c) keep[c] = c != skip1 && c != skip2;
                if (owner) {
#pragma unroll
                    for (int t = 0; t < PB; ++t) {
                        if (t < bw) {
                            double cv[8], rv[PB];
                            const double *pc_t = &Pcol[buf][t][rb];
                            const double *rr_t = &Rrow[buf][t][c0];
#pragma unroll
                            for (int h = 0; h < 4; ++h) { const double2 q = *reinterpret_cast<const double2 *>(pc_t + 32 * h); cv[2 * h] = q.x; cv[2 * h + 1] = q.y; }
#pragma unroll
                            for (int c = 0; c < PB; ++c) rv[c] = rr_t[c];
#pragma unroll
                            for (int c = 0; c < PB; ++c) {
                                const double m = keep[c] ? rv[c] : 0.0;
#pragma unroll
                                for (int r = 0; r < 8; ++r) a[r][c] = fma(cv[r], m, a[r][c]);
                            }
                        }
                    }
                }
                if (in_panel) {
                    // this thread's column w is column qt of the panel: it receives T[:, P] = (T - I)[:, P] + unit entry
                    const int prow = s_prow[buf][qt];
#define TAKE_COL(C) _Pragma("unroll") for (int r = 0; r < 8; ++r) a[r][C] = Pcol[buf][qt][rb + 32 * (r >> 1) + (r & 1)] + ((r0 + r == prow) ? 1.0 : 0.0);
                    switch (w) {
                    case 0: TAKE_COL(0) break;
                    case 1: TAKE_COL(1) break;
                    case 2: TAKE_COL(2) break;
                    case 3: TAKE_COL(3) break;
                    default: TAKE_COL(4) break;
                    }
#undef TAKE_COL
                }
            }
            // all pivots are known (the last pivot barrier has been passed): invert the permutation and store
            if (!fail && tid < n) qinv[perm[tid]] = tid;       // perm is incomplete after a failure
            nbar_sync(3, W);
            if (tid == 0) fail_flags[p] = fail;
            if (!fail && owner) {
                // A^-1[kk][cc] = W[perm[kk]][qinv[cc]]  <=>  W[i][j] goes to row qinv[i], column perm[j]
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = r0 + r;
                    if (i < n) {
                        const ll row = (ll)qinv[i] * ldg;
#pragma unroll
                        for (int c = 0; c < PB; ++c) {
                            const int j2 = c0 + c;
                            if (j2 < n) B[row + perm[j2]] = a[r][c];
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// The blocked look-ahead kernel with the searcher decoupled by one more panel.  Above, the searcher can only start
// panel j + 1 after the workers have passed the pivot barrier of panel j, published the pivot rows and updated the
// look-ahead columns: that front (~1/3 of the panel cycle) sits between two panel factorisations.  Here the workers
// publish the columns of panel j + 2 as they are after panel j, and the searcher applies panel j + 1 to them ITSELF
// (it holds (T - I)[:, P] of panel j + 1 in registers and finds the pivot rows' entries in the published columns:
// 100 FMAs per lane) before it factors them.  It therefore never waits for the workers' front, only for columns that
// were published a whole panel earlier; barriers alternate between two ids by panel parity so that a barrier is never
// arrived at twice within one phase.  Three panel buffers: M of panel j (workers reading), panel j + 1 (searcher),
// panel j + 2 (being published).  Same elimination order, pivots and arithmetic per step as the kernel above.
template <bool kPivot>
__global__ void __launch_bounds__(512, 1)
ssc_invert_panel2_kernel(double *__restrict__ blocks, int n, int ldg, ll stride, ll count, int *__restrict__ fail_flags, int TR, int TC, int W, int remap)
{
    constexpr int PB = 5, LDP = 136, SR = 6;
    __shared__ __align__(16) double Pcol[3][PB][LDP];      // panel columns: published by their owners, then (T - I)[:, P] from the searcher
    __shared__ __align__(16) double Rrow[2][PB][LDP];      // the raw pivot rows of a panel
    __shared__ __align__(16) double S[128][SR];            // the searcher's working panel, one row per matrix row
    __shared__ int s_prow[2][8];
    __shared__ int s_pg[32], s_pw[32], s_pbw[32], s_np;
    __shared__ int perm[128], qinv[128];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        int cntp = 0;
        for (int j = 0; j < PB * ((TC + PB - 1) / PB); ++j) {
            const int g = j / PB, w = j - PB * g;
            int bw = 0;
            for (int q = 0; q < PB; ++q) bw += (PB * g + q < TC && PB * (PB * g + q) + w < n) ? 1 : 0;
            if (bw) { s_pg[cntp] = g; s_pw[cntp] = w; s_pbw[cntp] = bw; ++cntp; }
        }
        s_np = cntp;
    }
    __syncthreads();
    const int NPE = s_np;
    if (remap && (wid == 5 || wid == 9)) {                   // see ssc_invert_panel_kernel
        for (ll p = blockIdx.x; p < count; p += gridDim.x) __syncthreads();
        return;
    }
    const bool searcher = remap ? wid == 13 : (int)threadIdx.x >= W;
    const int tid = remap ? ((wid - (wid > 5) - (wid > 9) - (wid > 13)) << 5) + lane : (int)threadIdx.x;
    const int all = W + 32;
    const int tc = tid / TR, tr = tid - tc * TR;
    const bool owner = !searcher && tc < TC;
    const int r0 = owner ? 8 * tr : 0, c0 = owner ? PB * tc : 0;
    const int gt = tc / PB, qt = tc - PB * gt;
    const int rb = owner ? 2 * tr : 0;                       // ppos(8 tr + 2 h + e) = 2 tr + 32 h + e: this thread's row pairs inside a panel column
    // barrier ids: 1 / 6 columns published (by parity of the panel they belong to), 2 / 5 pivots ready (parity of the panel), 3 pivot rows published
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        double *B = blocks + p * stride;
        if (searcher) {
            unsigned usedmask = 0;
#pragma unroll
            for (int s = 0; s < 4; ++s) if (lane + 32 * s >= n) usedmask |= 1u << s;
            double mprev[4][PB];                               // (T - I)[:, P] of the previous panel, this lane's rows
            int plprev[PB], bwprev = 0;
#pragma unroll
            for (int t = 0; t < PB; ++t) plprev[t] = 0;
            nbar_sync(1, all);                                    // the columns of the first two panels are there
            for (int it = 0; it < NPE; ++it) {
                const int bw = s_pbw[it], g = s_pg[it], w = s_pw[it];
                const int b3 = it % 3;
                if (it >= 2) nbar_sync((it & 1) ? 6 : 1, all);    // panel `it` was published (as of panel it - 2)
                double x[4][PB];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int i = lane + 32 * s;
#pragma unroll
                    for (int c = 0; c < PB; ++c) x[s][c] = (i < n && c < bw) ? Pcol[b3][c][ppos(i)] : 0.0;
                }
                if (it >= 1) {
                    // the previous panel applied to these columns: x += (T - I)[:, P_prev] * (rows P_prev of the published columns)
#pragma unroll
                    for (int t = 0; t < PB; ++t) {
                        if (t < bwprev) {
                            const int pp = ppos(plprev[t]);
#pragma unroll
                            for (int c = 0; c < PB; ++c) {
                                const double rv = c < bw ? Pcol[b3][c][pp] : 0.0;
#pragma unroll
                                for (int s = 0; s < 4; ++s) x[s][c] = fma(mprev[s][t], rv, x[s][c]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    double2 *row = reinterpret_cast<double2 *>(&S[lane + 32 * s][0]);
                    row[0] = make_double2(x[s][0], x[s][1]); row[1] = make_double2(x[s][2], x[s][3]); S[lane + 32 * s][4] = x[s][4];
                }
                __syncwarp();
                int pl[PB];
                bool bad = false;
#pragma unroll
                for (int t = 0; t < PB; ++t) {
                    pl[t] = 0;
                    if (t < bw && !bad) {
                        int prow = PB * (PB * g + t) + w;            // unpivoted: the diagonal entry of this column
                        if (kPivot) {
                            unsigned bh = 0u, bl = 0u;
                            int bi = 1 << 30;
#pragma unroll
                            for (int s = 0; s < 4; ++s) {
                                const unsigned kh = (unsigned)__double2hiint(x[s][t]) & 0x7fffffffu, kl = (unsigned)__double2loint(x[s][t]);
                                const bool better = !((usedmask >> s) & 1u) && (kh > bh || (kh == bh && kl > bl));
                                if (better) { bh = kh; bl = kl; bi = lane + 32 * s; }
                            }
                            const bool have = bi < (1 << 30);
                            const unsigned mh = __reduce_max_sync(0xffffffffu, have ? bh : 0u);
                            const bool in1 = have && bh == mh;
                            const unsigned ml = __reduce_max_sync(0xffffffffu, in1 ? bl : 0u);
                            prow = (int)__reduce_min_sync(0xffffffffu, (in1 && bl == ml) ? (unsigned)bi : 0x7fffffffu);
                        }
                        if (prow >= n) bad = true;
                        else {
                            const double2 *prw = reinterpret_cast<const double2 *>(&S[prow][0]);
                            const double2 p01 = prw[0], p23 = prw[1];
                            const double pr[PB] = {p01.x, p01.y, p23.x, p23.y, S[prow][4]};
                            const double d = pr[t];
                            if (!(fabs(d) > 0.0) || (!kPivot && !(d > 0.0))) bad = true;
                            const double dinv = __drcp_rn(d);        // correctly rounded, = 1.0 / d
                            double rr[PB];
#pragma unroll
                            for (int c = 0; c < PB; ++c) rr[c] = (c == t) ? dinv : pr[c] * dinv;
                            __syncwarp();                            // everybody has read the pivot row
#pragma unroll
                            for (int s = 0; s < 4; ++s) {
                                const int i = lane + 32 * s;
                                const double mlt = -x[s][t];
#pragma unroll
                                for (int c = 0; c < PB; ++c) x[s][c] = (c == t) ? mlt * dinv : fma(mlt, rr[c], x[s][c]);
                                double2 *row = reinterpret_cast<double2 *>(&S[i][0]);
                                row[0] = make_double2(x[s][0], x[s][1]); row[1] = make_double2(x[s][2], x[s][3]); S[i][4] = x[s][4];
                            }
                            if ((prow & 31) == lane) {
                                double2 *row = reinterpret_cast<double2 *>(&S[prow][0]);
                                row[0] = make_double2(rr[0], rr[1]); row[1] = make_double2(rr[2], rr[3]); S[prow][4] = rr[4];
                                usedmask |= 1u << (prow >> 5);
                            }
                            __syncwarp();
#pragma unroll
                            for (int s = 0; s < 4; ++s) {
                                const double2 *row = reinterpret_cast<const double2 *>(&S[lane + 32 * s][0]);
                                const double2 a01 = row[0], a23 = row[1];
                                x[s][0] = a01.x; x[s][1] = a01.y; x[s][2] = a23.x; x[s][3] = a23.y; x[s][4] = S[lane + 32 * s][4];
                            }
                            pl[t] = prow;
                        }
                    }
                }
                // (T - I)[:, P]: take the unit entries of the pivot rows out again; keep it for the next panel's columns
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int i = lane + 32 * s;
#pragma unroll
                    for (int c = 0; c < PB; ++c) {
                        mprev[s][c] = (c < bw) ? x[s][c] - ((i == pl[c]) ? 1.0 : 0.0) : 0.0;
                        if (i < n && c < bw) Pcol[b3][c][ppos(i)] = mprev[s][c];
                    }
                }
#pragma unroll
                for (int t = 0; t < PB; ++t) plprev[t] = pl[t];
                bwprev = bw;
                if (lane == 0) {
#pragma unroll
                    for (int t = 0; t < PB; ++t) { s_prow[it & 1][t] = bad ? -1 : pl[t]; if (t < bw) perm[PB * (PB * g + t) + w] = bad ? PB * (PB * g + t) + w : pl[t]; }
                }
                nbar_arrive((it & 1) ? 5 : 2, all);                // orders the stores above for the waiting workers
                if (bad) {
                    // the workers have already announced the columns of panel it + 1 (during panel it - 1): consume that
                    // arrival so the barrier starts clean for the next block
                    if (it >= 1 && it + 1 < NPE) nbar_sync(((it + 1) & 1) ? 6 : 1, all);
                    break;
                }
            }
        } else {
            double a[8][PB];
#pragma unroll
            for (int c = 0; c < PB; ++c) {
                const int j = c0 + c;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = r0 + r;
                    a[r][c] = (owner && i < n && j < n) ? B[(ll)j * ldg + i] : 0.0;      // transposed load: A = B^T
                }
            }
            // the first two panels as they are
#pragma unroll 1
            for (int it = 0; it < 2 && it < NPE; ++it) {
                const int g = s_pg[it], w = s_pw[it];
                if (owner && gt == g) {
#define PUT_COL(C) _Pragma("unroll") for (int h = 0; h < 4; ++h) *reinterpret_cast<double2 *>(&Pcol[it][qt][rb + 32 * h]) = make_double2(a[2 * h][C], a[2 * h + 1][C]);
                    switch (w) {
                    case 0: PUT_COL(0) break;
                    case 1: PUT_COL(1) break;
                    case 2: PUT_COL(2) break;
                    case 3: PUT_COL(3) break;
                    default: PUT_COL(4) break;
                    }
#undef PUT_COL
                }
            }
            nbar_arrive(1, all);
            int fail = 0;
            for (int it = 0; it < NPE; ++it) {
                const int bw = s_pbw[it], b3 = it % 3, b2 = it & 1, g = s_pg[it], w = s_pw[it];
                const bool more = it + 2 < NPE;                        // the panel two ahead
                const int bwn = more ? s_pbw[it + 2] : 0, gn = more ? s_pg[it + 2] : -1, wn = more ? s_pw[it + 2] : 0, bn3 = (it + 2) % 3;
                nbar_sync(b2 ? 5 : 2, all);                           // the pivots of this panel are known
                int pl[PB];
#pragma unroll
                for (int t = 0; t < PB; ++t) pl[t] = s_prow[b2][t];
                if (pl[0] < 0) { fail = 1; break; }                   // no usable pivot: uniform across the CTA
                const bool in_panel = owner && gt == g && qt < bw;
                const int myprow = in_panel ? s_prow[b2][qt] : -1;    // read now: the searcher may reuse the slot two panels on
#pragma unroll
                for (int t = 0; t < PB; ++t) {
                    if (t < bw && owner && tr == (pl[t] >> 3)) {
#define PUBLISH_ROW(R) _Pragma("unroll") for (int c = 0; c < PB; ++c) Rrow[b2][t][c0 + c] = a[R][c];
                        switch (pl[t] & 7) {
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
                }
                nbar_sync(3, W);
                const bool in_next = owner && more && gt == gn && qt < bwn;
                // look-ahead: the columns of the panel two ahead first, then they go to the searcher
                if (in_next) {
#define EARLY_COL(C)                                                                                                    \
    _Pragma("unroll") for (int t = 0; t < PB; ++t) {                                                                   \
        if (t < bw) {                                                                                                   \
            const double rvc = Rrow[b2][t][c0 + C];                                                                     \
            _Pragma("unroll") for (int h = 0; h < 4; ++h) {                                                             \
                const double2 q2 = *reinterpret_cast<const double2 *>(&Pcol[b3][t][rb + 32 * h]);                       \
                a[2 * h][C] = fma(q2.x, rvc, a[2 * h][C]); a[2 * h + 1][C] = fma(q2.y, rvc, a[2 * h + 1][C]);           \
            }                                                                                                           \
        }                                                                                                               \
    }                                                                                                                   \
    _Pragma("unroll") for (int h = 0; h < 4; ++h) *reinterpret_cast<double2 *>(&Pcol[bn3][qt][rb + 32 * h]) = make_double2(a[2 * h][C], a[2 * h + 1][C]);
                    switch (wn) {
                    case 0: EARLY_COL(0) break;
                    case 1: EARLY_COL(1) break;
                    case 2: EARLY_COL(2) break;
                    case 3: EARLY_COL(3) break;
                    default: EARLY_COL(4) break;
                    }
#undef EARLY_COL
                }
                if (more) nbar_arrive(((it + 2) & 1) ? 6 : 1, all);
                // branch-free: a column that must not be touched (this panel's own column: replaced below; the
                // look-ahead column: done above) gets the multiplier 0, so all nine loads of a rank are issued
                // together and the 40 FMAs follow back to back
                const int skip1 = in_panel ? w : -1, skip2 = in_next ? wn : -1;
                bool keep[PB];
#pragma unroll
                for (int c = 0; c < PB; ++c) keep[c] = c != skip1 && c != skip2;
                if (owner) {
#pragma unroll
                    for (int t = 0; t < PB; ++t) {
                        if (t < bw) {
                            double cv[8], rv[PB];
                            const double *pc_t = &Pcol[b3][t][rb];
                            const double *rr_t = &Rrow[b2][t][c0];
#pragma unroll
                            for (int h = 0; h < 4; ++h) { const double2 q = *reinterpret_cast<const double2 *>(pc_t + 32 * h); cv[2 * h] = q.x; cv[2 * h + 1] = q.y; }
#pragma unroll
                            for (int c = 0; c < PB; ++c) rv[c] = rr_t[c];
#pragma unroll
                            for (int c = 0; c < PB; ++c) {
                                const double m = keep[c] ? rv[c] : 0.0;
#pragma unroll
                                for (int r = 0; r < 8; ++r) a[r][c] = fma(cv[r], m, a[r][c]);
                            }
                        }
                    }
                }
                if (in_panel) {
#define TAKE_COL(C) _Pragma("unroll") for (int r = 0; r < 8; ++r) a[r][C] = Pcol[b3][qt][rb + 32 * (r >> 1) + (r & 1)] + ((r0 + r == myprow) ? 1.0 : 0.0);
                    switch (w) {
                    case 0: TAKE_COL(0) break;
                    case 1: TAKE_COL(1) break;
                    case 2: TAKE_COL(2) break;
                    case 3: TAKE_COL(3) break;
                    default: TAKE_COL(4) break;
                    }
#undef TAKE_COL
                }
            }
            if (!fail && tid < n) qinv[perm[tid]] = tid;       // perm is incomplete after a failure
            nbar_sync(3, W);
            if (tid == 0) fail_flags[p] = fail;
            if (!fail && owner) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = r0 + r;
                    if (i < n) {
                        const ll row = (ll)qinv[i] * ldg;
#pragma unroll
                        for (int c = 0; c < PB; ++c) {
                            const int j2 = c0 + c;
                            if (j2 < n) B[row + perm[j2]] = a[r][c];
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// Same algorithm with 8 x 8 tiles: (n/8)^2 threads, 64 FMAs per thread and step against 16 shared-memory loads and
// half as many warps repeating the pivot search -- the 8 x 4 form above is issue-bound on exactly that overhead.
template <bool kPivot>
__global__ void __launch_bounds__(256, 1)
ssc_invert_reg8_kernel(double *__restrict__ blocks, int n, int ldg, ll stride, ll count, int *__restrict__ fail_flags, int TC)
{
    __shared__ __align__(16) double colv[2][128];
    __shared__ __align__(16) double rowv[2][128];
    __shared__ int perm[128], qinv[128];
    const int tid = threadIdx.x, lane = tid & 31;
    const int tr = tid / TC, tc = tid - tr * TC;
    const bool owner = tr < TC;                              // TC = ceil(n / 8) tiles in both directions
    const int r0 = owner ? 8 * tr : 0, c0 = 8 * tc;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        double *B = blocks + p * stride;
        double a[8][8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = c0 + c;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = r0 + r;
                a[r][c] = (owner && i < n && j < n) ? B[(ll)j * ldg + i] : 0.0;      // transposed load: A = B^T
            }
        }
        unsigned usedmask = 0;
        int fail = 0;
        for (int k = 0; k < n; ++k) {
            const int buf = k & 1;
            if (owner && tc == (k >> 3)) {
#define PUBLISH_COL(C)                                                                  \
    _Pragma("unroll") for (int r = 0; r < 8; ++r) { colv[buf][r0 + r] = a[r][C]; a[r][C] = 0.0; }
                switch (k & 7) {
                case 0: PUBLISH_COL(0) break;
                case 1: PUBLISH_COL(1) break;
                case 2: PUBLISH_COL(2) break;
                case 3: PUBLISH_COL(3) break;
                case 4: PUBLISH_COL(4) break;
                case 5: PUBLISH_COL(5) break;
                case 6: PUBLISH_COL(6) break;
                default: PUBLISH_COL(7) break;
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
                if (prow >= n) { fail = 1; break; }
                if ((prow & 31) == lane) usedmask |= 1u << (prow >> 5);
                if (tid == 0) perm[k] = prow;
            }
            const double d = colv[buf][prow];
            if (!(fabs(d) > 0.0) || (!kPivot && !(d > 0.0))) { fail = 1; break; }     // uniform across the CTA
            const double dinv = 1.0 / d;
            if (owner && tr == (prow >> 3)) {
#define PUBLISH_ROW(R)                                                                                          \
    _Pragma("unroll") for (int c = 0; c < 8; ++c) { rowv[buf][c0 + c] = (c0 + c == k) ? dinv : a[R][c] * dinv; a[R][c] = 0.0; }
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
            const double2 *cp = reinterpret_cast<const double2 *>(&colv[buf][r0]);
            const double2 *rp2 = reinterpret_cast<const double2 *>(&rowv[buf][c0]);
            double cv[8], rv[8];
#pragma unroll
            for (int h = 0; h < 4; ++h) { const double2 t = cp[h]; cv[2 * h] = t.x; cv[2 * h + 1] = t.y; }
#pragma unroll
            for (int h = 0; h < 4; ++h) { const double2 t = rp2[h]; rv[2 * h] = t.x; rv[2 * h + 1] = t.y; }
            const int rp = prow - r0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double m = (r == rp) ? 1.0 : -cv[r];
#pragma unroll
                for (int c = 0; c < 8; ++c) a[r][c] = fma(m, rv[c], a[r][c]);
            }
        }
        __syncthreads();
        if (tid < n) { if (!kPivot) perm[tid] = tid; }
        __syncthreads();
        if (tid < n) qinv[perm[tid]] = tid;
        __syncthreads();
        if (tid == 0) fail_flags[p] = fail;
        if (!fail && owner) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = r0 + r;
                if (i < n) {
                    const ll row = (ll)qinv[i] * ldg;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
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
