// MEASURED AND REJECTED (round 2; kept for the probe harness tools/micro/factor_probe.cu, not part of the library):
// 142 ms (lu) / 122 ms (cholesky) at 64^3 Q3 against 118 / 111 ms for the L2-resident kernel that ships (ssc_b200/csrc/kernels_factor_l2.cuh).
// The pivot chain's fp64 FMAs share the FP64 pipe with the DMMA updates, so the look-ahead overlap is lost: cost = chain + update.
//
// K2 for n <= 128: blocked Gauss-Jordan with the whole block resident in REGISTERS as fp64 tensor-core accumulators,
// and the panel factorisation running one panel ahead of the update (look-ahead) on warps of its own.
//
// Replaces the lazy PCSetUp_LU inside the first KSPSolve of every patch (ssc/libssc.c:1374) by an explicit inverse.
//
// Layout.  One CTA per patch: G x G worker warps + G panel warps (G = 4: n <= 128, 3: n <= 96, 2: n <= 64).  Worker
// warp (wr, wc) owns the 32 x 32 sub-block (wr, wc) of M = B^T as 4 x 4 accumulator fragments of
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4): lane l holds M[i][j], M[i][j+1] with i = 32 wr + 8 ti + (l >> 2),
// j = 32 wc + 8 tj + 2 (l & 3) -- 64 registers per thread, no copy of the block anywhere else.  Shared memory only
// holds operand panels: three rotating n x 16 panel buffers and the 16 x n pivot rows W.
//
// Algorithm, NB = 16 columns per panel, rows never move (implicit pivoting, the permutation is applied at the store).
// Iteration j of the workers:
//   A. the owners of panel j+1's columns apply panel j to those columns first and publish them (n x 16) -> barrier B1;
//   B. all other columns get  M <- M + Q_j W_j  on the tensor cores (4 A fragments of Q + 4 B fragments of W from
//      shared memory per 16 DMMAs); panel j's own columns are reloaded from Q_j;
//   C. barrier B2: panel j+1 is factored;
//   D. every thread whose row is a pivot row of panel j+1 publishes its row entries to W and clears them -> barrier B3.
// Meanwhile the panel warps (thread = row, its 16 panel entries in registers) run the 16 Gauss-Jordan steps of panel
// j+1 between B1 and B2.  One named barrier per step: before it every row publishes its updated entries, the
// reciprocal of its candidate pivot and (per warp, one redux.sync.max.u32) a key that packs the high word of |a|
// (13 mantissa bits) with 127 - row, so the arg-max comes out of the same reduction (threshold partial pivoting: the
// chosen entry is within 2^-13 of the column maximum, lowest row on ties); after it every row reads the pivot row
// and its reciprocal and does 16 FMAs, the next pivot column first.  The panel then holds Q = the non-trivial columns
// of the panel's transformation T.
// 2 n^3 flop per block, of which all but the n x 16 panel steps are DMMA.  Result stored as S = (B^-1)^T row-major,
// like every other factor kernel: S[qinv[i]][perm[j]] = M[i][j].
#pragma once
#include "../../ssc_b200/csrc/kernels_factor_l2.cuh"      // ssc_dmma_acc, SSC_DMMA_PROBE

__device__ __forceinline__ void ssc_nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void ssc_nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int G> struct DmmaFactorSmem {
    static constexpr int NB = 16, N = 32 * G, QP = 20, WP = N + 4;      // QP, WP = 4 mod 16: conflict-free A / B fragment loads
    static constexpr size_t bytes = (size_t)(2 * N * QP + NB * WP + 2 * (NB + 2)) * sizeof(double) + (size_t)(8 + 3 * N + 4) * sizeof(int);
};

template <bool kPivot, int G>
__global__ void __launch_bounds__(32 * G * (G + 1), 1)
ssc_invert_dmma_kernel(double *__restrict__ blocks, int n, int ldg, ll stride, ll count, int *__restrict__ fail_flags)
{
    typedef DmmaFactorSmem<G> L;
    constexpr int NB = L::NB, N = L::N, QP = L::QP, WP = L::WP;
    constexpr int NW = 32 * G * G;                                // worker threads
    constexpr int NT = 32 * G * (G + 1);
    constexpr int B1 = 1, B2 = 2, B3 = 3, B4 = 4, B5 = 5;         // named barriers; B5 is CTA-wide (the two roles meet it from different code)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *R = reinterpret_cast<double *>(smem_raw);             // [2][N * QP] panel buffers: panel j (being applied) and panel j+1 (being factored)
    double *Wbuf = R + 2 * N * QP;                                // [NB][WP] pivot rows of the current panel, all columns
    double *rowv = Wbuf + NB * WP;                                // [2][NB + 2] the pivot row of a step and the reciprocal of its pivot, by step parity
    unsigned *keyb = reinterpret_cast<unsigned *>(rowv + 2 * (NB + 2));  // [2][4] per-warp maxima of the pivot keys
    int *perm = reinterpret_cast<int *>(keyb + 8);                // step -> pivot row
    int *rowk = perm + N;                                         // row -> step at which it was the pivot (-1: not yet)
    int *qinv = rowk + N;
    volatile int *s_fail = qinv + N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool worker = tid < NW;
    const int wr = warp / G, wc = warp - wr * G, gr = lane >> 2, gc = lane & 3;
    const int npanels = (n + NB - 1) / NB;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        double *B = blocks + p * stride;
        if (worker) {
            double2 c[4][4];                                      // lives in the worker branch only: the panel warps' registers hold panel rows
            // M = B^T: lane reads B[j][i], B[j+1][i]; per load instruction a warp touches 4 rows of B x 64 contiguous bytes
#pragma unroll
            for (int ti = 0; ti < 4; ++ti) {
                const int i = 32 * wr + 8 * ti + gr;
#pragma unroll
                for (int tj = 0; tj < 4; ++tj) {
                    const int j = 32 * wc + 8 * tj + 2 * gc;
                    c[ti][tj].x = (i < n && j < n) ? B[(ll)j * ldg + i] : 0.0;
                    c[ti][tj].y = (i < n && j + 1 < n) ? B[(ll)(j + 1) * ldg + i] : 0.0;
                }
            }
            // iteration j applies panel j (j >= 0) and prepares panel j + 1
            for (int j = -1; j < npanels; ++j) {
                const bool have_next = j + 1 < npanels;
                const double *Q = R + ((j + 2) % 2) * (N * QP);
                // halves of this warp's columns: h = 0 -> fragment columns 0, 1; h = 1 -> 2, 3; half h is panel 2 wc + h
                bool doh[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) doh[h] = j >= 0 && 2 * wc + h != j && 2 * wc + h != j + 1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (have_next && 2 * wc + h == j + 1) {
                        // A. panel j+1's columns first
                        if (j >= 0) {
#pragma unroll
                            for (int q = 0; q < NB / 4; ++q) {
                                double a[4], b[2];
#pragma unroll
                                for (int ti = 0; ti < 4; ++ti) a[ti] = Q[(32 * wr + 8 * ti + gr) * QP + 4 * q + gc];
#pragma unroll
                                for (int t = 0; t < 2; ++t) b[t] = Wbuf[(4 * q + gc) * WP + 32 * wc + 8 * (2 * h + t) + gr];
#pragma unroll
                                for (int t = 0; t < 2; ++t)
#pragma unroll
                                    for (int ti = 0; ti < 4; ++ti) ssc_dmma_acc(c[ti][2 * h + t], a[ti], b[t]);
                            }
                        }
                        double *In = R + ((j + 1) % 2) * (N * QP);
#pragma unroll
                        for (int ti = 0; ti < 4; ++ti) {
                            double2 *dst = reinterpret_cast<double2 *>(In + (32 * wr + 8 * ti + gr) * QP + 2 * gc);
                            dst[0] = c[ti][2 * h];
                            dst[4] = c[ti][2 * h + 1];
                        }
                        ssc_nbar_arrive(B1, 64 * G);
                    }
                }
                if (j >= 0) {
                    // B. all other columns; panel j's own columns become Q_j
                    if (doh[0] || doh[1]) {
#pragma unroll
                        for (int q = 0; q < NB / 4; ++q) {
                            double a[4], b[4];
#pragma unroll
                            for (int ti = 0; ti < 4; ++ti) a[ti] = Q[(32 * wr + 8 * ti + gr) * QP + 4 * q + gc];
#pragma unroll
                            for (int tj = 0; tj < 4; ++tj) b[tj] = Wbuf[(4 * q + gc) * WP + 32 * wc + 8 * tj + gr];
#pragma unroll
                            for (int tj = 0; tj < 4; ++tj) {
                                if (!doh[tj >> 1]) continue;                     // warp-uniform
#pragma unroll
                                for (int ti = 0; ti < 4; ++ti) ssc_dmma_acc(c[ti][tj], a[ti], b[tj]);
                            }
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (2 * wc + h == j) {
#pragma unroll
                            for (int ti = 0; ti < 4; ++ti) {
                                const double2 *src = reinterpret_cast<const double2 *>(Q + (32 * wr + 8 * ti + gr) * QP + 2 * gc);
                                c[ti][2 * h] = src[0];
                                c[ti][2 * h + 1] = src[4];
                            }
                        }
                    }
                }
                if (!have_next) break;
                ssc_nbar_sync(B2, NW + 32 * G);                                  // C. panel j+1 is factored, everybody is done with W_j
                if (*s_fail) break;
                // D. the pivot rows of panel j+1: publish (all columns) and clear; rows of W a short last panel does not use are zero
                const int kb = NB * (j + 1), bw = min(NB, n - kb);
                if (bw < NB) for (int e = tid; e < (NB - bw) * WP; e += NW) Wbuf[bw * WP + e] = 0.0;
#pragma unroll
                for (int ti = 0; ti < 4; ++ti) {
                    const int k = rowk[32 * wr + 8 * ti + gr];
                    if (k >= kb) {
                        double2 *dst = reinterpret_cast<double2 *>(Wbuf + (k - kb) * WP + 32 * wc + 2 * gc);
#pragma unroll
                        for (int tj = 0; tj < 4; ++tj) { dst[4 * tj] = c[ti][tj]; c[ti][tj] = make_double2(0.0, 0.0); }
                    }
                }
                ssc_nbar_sync(B3, NW);
            }
            ssc_nbar_sync(B5, NT);
            const int fail = *s_fail;
            if (tid < n && !fail) qinv[perm[tid]] = tid;
            ssc_nbar_sync(B5, NT);
            if (tid == 0) fail_flags[p] = fail;
            if (!fail) {
                // M[i][j] = (B^T)^-1 [qinv[i]][perm[j]] = S[qinv[i]][perm[j]]
#pragma unroll
                for (int ti = 0; ti < 4; ++ti) {
                    const int i = 32 * wr + 8 * ti + gr;
                    if (i < n) {
                        double *row = B + (ll)qinv[i] * ldg;
#pragma unroll
                        for (int tj = 0; tj < 4; ++tj) {
                            const int j = 32 * wc + 8 * tj + 2 * gc;
                            if (j < n) row[perm[j]] = c[ti][tj].x;
                            if (j + 1 < n) row[perm[j + 1]] = c[ti][tj].y;
                        }
                    }
                }
            }
            ssc_nbar_sync(B5, NT);
        } else {
            const int i = tid - NW, pwarp = warp - G * G;
            perm[i] = i; rowk[i] = -1;
            if (i == 0) *s_fail = 0;
            bool used = i >= n;                                   // row i cannot be a pivot row (any more)
            for (int j = -1; j + 1 < npanels; ++j) {
                const int kb = NB * (j + 1), bw = min(NB, n - kb);
                double *Pj = R + ((j + 1) % 2) * (N * QP);        // panel j+1: published by its owners, factored in registers, written back as Q
                ssc_nbar_sync(B1, 64 * G);
                double pr[NB];
                {
                    const double2 *src = reinterpret_cast<const double2 *>(Pj + i * QP);
#pragma unroll
                    for (int h = 0; h < NB / 2; ++h) { const double2 t = src[h]; pr[2 * h] = t.x; pr[2 * h + 1] = t.y; }
                }
                // candidates of step 0: this row's key (one redux per warp) and the reciprocal of its entry
                if (kPivot) {
                    unsigned key = used ? 0u : (((unsigned)__double2hiint(pr[0]) & 0x7FFFFF80u) | (unsigned)(127 - i));
                    key = __reduce_max_sync(0xffffffffu, key);
                    if (lane == 0) keyb[pwarp] = key;
                }
                double myinv = 1.0 / pr[0];
                bool bad = false;
#pragma unroll
                for (int s = 0; s < NB; ++s) {
                    if (s >= bw || SSC_DMMA_PROBE >= 2) break;
                    const int cur = s & 1;
                    int prow = kb + s;
                    if (kPivot) {
                        ssc_nbar_sync(B4, 32 * G);
                        unsigned m = keyb[4 * cur];
#pragma unroll
                        for (int w = 1; w < G; ++w) m = max(m, keyb[4 * cur + w]);
                        prow = 127 - (int)(m & 127u);
                        if (m == 0) { bad = true; break; }                       // no candidate left (uniform)
                    }
                    const bool win = i == prow;
                    if (win) {
                        // the pivot row as it stands, and the reciprocal of the pivot
                        double2 *dst = reinterpret_cast<double2 *>(rowv + cur * (NB + 2));
#pragma unroll
                        for (int h = 0; h < NB / 2; ++h) dst[h] = make_double2(pr[2 * h], pr[2 * h + 1]);
                        rowv[cur * (NB + 2) + NB] = myinv;
                        used = true; perm[kb + s] = i; rowk[i] = kb + s;
                    }
                    ssc_nbar_sync(B4, 32 * G);
                    const double dinv = rowv[cur * (NB + 2) + NB];
                    if (kPivot ? !(fabs(dinv) > 0.0 && fabs(dinv) < 1.0e300) : !(dinv > 0.0 && dinv < 1.0e300)) { bad = true; break; }   // uniform
                    double rv[NB];
                    {
                        const double2 *src = reinterpret_cast<const double2 *>(rowv + cur * (NB + 2));
#pragma unroll
                        for (int h = 0; h < NB / 2; ++h) { const double2 t = src[h]; rv[2 * h] = t.x; rv[2 * h + 1] = t.y; }
                    }
                    // one update formula for all rows: the pivot row starts from zero with the multiplier 1/d, every other row
                    // from itself with -a[i][s]/d; entry s becomes the multiplier itself (1/d, resp. -a[i][s]/d)
                    const double g = win ? dinv : -pr[s] * dinv;
                    if (win) {
#pragma unroll
                        for (int jj = 0; jj < NB; ++jj) pr[jj] = 0.0;
                    }
                    if (s + 1 < NB) pr[s + 1] = fma(g, rv[s + 1], pr[s + 1]);                // the next pivot column first
                    if (s + 1 < bw) {
                        if (kPivot) {
                            unsigned key = used ? 0u : (((unsigned)__double2hiint(pr[s + 1]) & 0x7FFFFF80u) | (unsigned)(127 - i));
                            key = __reduce_max_sync(0xffffffffu, key);
                            if (lane == 0) keyb[4 * (cur ^ 1) + pwarp] = key;
                        }
                        myinv = 1.0 / pr[s + 1];
                    }
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) if (jj != s && jj != s + 1) pr[jj] = fma(g, rv[jj], pr[jj]);
                    pr[s] = g;
                }
                {
                    double2 *dst = reinterpret_cast<double2 *>(Pj + i * QP);
#pragma unroll
                    for (int h = 0; h < NB / 2; ++h) dst[h] = make_double2(pr[2 * h], pr[2 * h + 1]);
                }
                if (bad) *s_fail = 1;
                ssc_nbar_sync(B2, NW + 32 * G);
                if (*s_fail) break;
            }
            ssc_nbar_sync(B5, NT);                                // the three CTA-wide barriers of the workers' epilogue
            ssc_nbar_sync(B5, NT);
            ssc_nbar_sync(B5, NT);
        }
    }
}
