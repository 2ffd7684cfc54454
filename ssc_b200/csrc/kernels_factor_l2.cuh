// K2 for 32 < n <= 512, throughput form: blocked Gauss-Jordan with the block resident in L2 and SEVERAL patches in flight per SM.
//
// Replaces the lazy PCSetUp_LU inside the first KSPSolve of every patch (ssc/libssc.c:1374) by an explicit inverse.
//
// Why.  A pivoted elimination is a chain of n dependent steps (search, reciprocal, broadcast, update: ~700 cycles each on a
// lone warp group), and fp64 FMAs of that chain share the FP64 pipe with the tensor-core updates, so a design with ONE patch
// per SM (tools/micro/factor_reg_dmma.cuh: block in registers as DMMA accumulators, panel one ahead; measured 142 ms at 64^3 Q3
// against 118 ms here) cannot overlap the two: it costs chain + update.  Here a CTA is
// 128 threads and four of them share an SM (256 threads and two for 128 < n <= 256, 512 threads and one beyond), each on its own patch, so the SM always has other patches' work to issue while
// one waits in its chain; the price is that the block (125 KB at n = 125) lives in L2 and every panel's update streams it
// through the SM once (2 MB of L2 traffic per patch; measured L2 read+write bandwidth 14 TB/s, tools/micro/l2_micro.cu).
//
// Algorithm (M = B row-major in place, row stride ld EVEN so that every row and every fragment pair starts on 16 bytes,
// NB = 16 columns per panel, rows never move: implicit pivoting):
//   1. thread i loads the 16 panel entries of row i into registers;
//   2. the 16 Gauss-Jordan steps of the panel, thread = row: key = high word of |a| (13 mantissa bits) packed
//      with 127 - row, one redux.sync.max.u32 per warp + one shared-memory exchange gives the pivot row (threshold partial
//      pivoting, 2^-13), the winner publishes its row and the reciprocal of the pivot (computed ahead), two barriers per step;
//   3. the panel (= Q, the non-trivial columns of the panel's transformation) goes to shared memory and back to the block;
//      the 16 pivot rows W are read from the block (contiguous rows);
//   4. all other columns:  M <- M + Q W  with pivot rows starting from zero, on the tensor cores (mma.sync.m8n8k4.f64, SASS
//      DMMA.8x8x4): a warp owns tile rows, keeps the four A fragments of a tile row and walks the tile columns in strips of four.
// The array then holds M[i][j] = B^-1[qinv[i]][perm[j]]; an in-place transpose and (only if some pivot was off the diagonal)
// a row and a column permutation turn it into S = (B^-1)^T row-major, the layout every factor kernel leaves.
#pragma once
#include "kernels_common.cuh"

#ifndef SSC_DMMA_PROBE
#define SSC_DMMA_PROBE 0        // tools/micro/factor_probe.cu: 1 skips the tensor-core updates, 2 the panel steps, 3 both (timing experiments only)
#endif
__device__ __forceinline__ void ssc_dmma_acc(double2 &c, double a, double b)
{
    if (SSC_DMMA_PROBE == 1 || SSC_DMMA_PROBE == 3) return;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}

// N = 128 (n <= 128, 128 threads, four CTAs per SM), 256 (n <= 256, two CTAs per SM) or 512 (n <= 512, one CTA per SM): N threads, thread = row.
template <bool kPivot, int NB, int N>
__global__ void __launch_bounds__(N, N == 512 ? 1 : (N == 256 ? 2 : (NB == 16 ? 4 : 3)))
ssc_invert_l2_kernel(double *__restrict__ blocks, int n, int ld, ll stride, ll count, int *__restrict__ fail_flags)
{
    constexpr int QP = NB + 4, WP = N + 4, NWARP = N / 32;        // QP, WP = 4 mod 16: conflict-free A / B fragment loads
    constexpr unsigned kRowMask = N - 1;                          // the pivot key packs N - 1 - row into its low bits
    extern __shared__ __align__(16) unsigned char l2_smem[];      // (N * QP + NB * WP) doubles
    double *Pbuf = reinterpret_cast<double *>(l2_smem);           // Q of the current panel
    double *Wbuf = Pbuf + N * QP;                                 // its pivot rows, all columns
    __shared__ __align__(16) double rowv[2][NB + 2];              // the pivot row of a step + the reciprocal of its pivot, by step parity
    __shared__ __align__(16) unsigned keyb[2][NWARP];
    __shared__ int perm[N], rowk[N], qinv[N], leaders[N];
    __shared__ int s_fail, s_offdiag, s_nleaders;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gr = lane >> 2, gc = lane & 3;
    const int i = tid;                                             // the row this thread owns in the panel steps
    const int TR = (n + 7) >> 3;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        double *B = blocks + p * stride;
        perm[i] = i; rowk[i] = -1;
        if (i == 0) { s_fail = 0; s_offdiag = 0; }
        bool used = i >= n;
        __syncthreads();
        for (int kb = 0; kb < n; kb += NB) {
            const int bw = min(NB, n - kb);
            double pr[NB];
            {
                // ld is even and kb a multiple of 16: 16-byte loads; pairs that start inside the row are whole (the last may end in the padding column)
                const double2 *src = reinterpret_cast<const double2 *>(B + (ll)i * ld + kb);
#pragma unroll
                for (int h = 0; h < NB / 2; ++h) {
                    double2 t = make_double2(0.0, 0.0);
                    if (i < n && kb + 2 * h < ld) t = src[h];
                    pr[2 * h] = 2 * h < bw ? t.x : 0.0;
                    pr[2 * h + 1] = 2 * h + 1 < bw ? t.y : 0.0;
                }
            }
            if (kPivot) {
                unsigned key = used ? 0u : (((unsigned)__double2hiint(pr[0]) & (0x7FFFFFFFu & ~kRowMask)) | (kRowMask - (unsigned)i));
                key = __reduce_max_sync(0xffffffffu, key);
                if (lane == 0) keyb[0][warp] = key;
            }
            double myinv = 1.0 / pr[0];
            bool bad = false;
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                if (s >= bw || SSC_DMMA_PROBE >= 2) break;
                const int cur = s & 1;
                int prow = kb + s;
                if (kPivot) {
                    __syncthreads();
                    unsigned m = 0;
#pragma unroll
                    for (int w4 = 0; w4 < NWARP / 4; ++w4) {
                        const uint4 k4 = reinterpret_cast<const uint4 *>(keyb[cur])[w4];
                        m = max(m, max(max(k4.x, k4.y), max(k4.z, k4.w)));
                    }
                    prow = (int)(kRowMask - (m & kRowMask));
                    if (m == 0) { bad = true; break; }                           // no candidate left (uniform)
                }
                const bool win = i == prow;
                if (win) {
                    double2 *dst = reinterpret_cast<double2 *>(rowv[cur]);
#pragma unroll
                    for (int h = 0; h < NB / 2; ++h) dst[h] = make_double2(pr[2 * h], pr[2 * h + 1]);
                    rowv[cur][NB] = myinv;
                    used = true; perm[kb + s] = i; rowk[i] = kb + s;
                    if (i != kb + s) s_offdiag = 1;
                }
                __syncthreads();
                const double dinv = rowv[cur][NB];
                if (kPivot ? !(fabs(dinv) > 0.0 && fabs(dinv) < 1.0e300) : !(dinv > 0.0 && dinv < 1.0e300)) { bad = true; break; }   // uniform
                double rv[NB];
                {
                    const double2 *src = reinterpret_cast<const double2 *>(rowv[cur]);
#pragma unroll
                    for (int h = 0; h < NB / 2; ++h) { const double2 t = src[h]; rv[2 * h] = t.x; rv[2 * h + 1] = t.y; }
                }
                if (win) {
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) pr[jj] = (jj == s) ? dinv : rv[jj] * dinv;
                } else {
                    const double g = -pr[s] * dinv;
                    if (s + 1 < NB) pr[s + 1] = fma(g, rv[s + 1], pr[s + 1]);                // the next pivot column first
#pragma unroll
                    for (int jj = 0; jj < NB; ++jj) if (jj != s && jj != s + 1) pr[jj] = fma(g, rv[jj], pr[jj]);
                    pr[s] = g;
                }
                if (s + 1 < bw) {
                    if (kPivot) {
                        unsigned key = used ? 0u : (((unsigned)__double2hiint(pr[s + 1]) & (0x7FFFFFFFu & ~kRowMask)) | (kRowMask - (unsigned)i));
                        key = __reduce_max_sync(0xffffffffu, key);
                        if (lane == 0) keyb[cur ^ 1][warp] = key;
                    }
                    myinv = 1.0 / pr[s + 1];
                }
            }
            if (bad) s_fail = 1;
            {
                double2 *dst = reinterpret_cast<double2 *>(Pbuf + i * QP);
#pragma unroll
                for (int h = 0; h < NB / 2; ++h) dst[h] = make_double2(pr[2 * h], pr[2 * h + 1]);
            }
            __syncthreads();
            if (s_fail) break;
            {
                double2 *dst = reinterpret_cast<double2 *>(B + (ll)i * ld + kb);
#pragma unroll
                for (int h = 0; h < NB / 2; ++h) if (i < n && 2 * h < bw) dst[h] = make_double2(pr[2 * h], pr[2 * h + 1]);   // pr beyond bw is zero: the padding column stays zero
            }
            // the pivot rows of this panel, all columns but the panel's own (those tiles are not updated)
            {
                // thread tid owns column j = tid of W: 16 independent row reads in flight, then 16 stores
                const int j = tid;
                const bool colok = j < n && (j < kb || j >= kb + NB);   // (the panel's own tile columns are not updated)
                double v[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) v[k] = (k < bw && colok) ? B[(ll)perm[kb + k] * ld + j] : 0.0;
#pragma unroll
                for (int k = 0; k < NB; ++k) Wbuf[k * WP + j] = v[k];
            }
            __syncthreads();
            // M <- M + Q W on the tensor cores; pivot rows of this panel start from zero
            const int ptile = kb >> 3;                                           // the panel is tile columns ptile .. ptile + NB/8 - 1
            for (int ti = warp; ti < TR; ti += NWARP) {
                const int r = 8 * ti + gr;
                const bool valid = r < n;
                const bool fresh = valid && rowk[r] >= kb;
                double a[NB / 4];
#pragma unroll
                for (int q = 0; q < NB / 4; ++q) a[q] = Pbuf[r * QP + 4 * q + gc];
                double *Mrow = B + (ll)r * ld;
                // strips of 8 tiles, software-pipelined: the 16 loads of the next strip are in flight while this strip's DMMAs issue
                auto load_strip = [&](int tj0, double2 (&c)[8]) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int tj = tj0 + u, j = 8 * tj + 2 * gc;
                        const bool skip = tj >= TR || (tj >= ptile && tj < ptile + NB / 8);
                        c[u] = make_double2(0.0, 0.0);
                        if (!skip && valid && !fresh && j < ld) c[u] = *reinterpret_cast<const double2 *>(Mrow + j);     // j, ld even: the pair is inside the row
                    }
                };
                auto finish_strip = [&](int tj0, double2 (&c)[8]) {
#pragma unroll
                    for (int q = 0; q < NB / 4; ++q) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int tj = tj0 + u;
                            if (tj >= TR || (tj >= ptile && tj < ptile + NB / 8)) continue;      // warp-uniform
                            ssc_dmma_acc(c[u], a[q], Wbuf[(4 * q + gc) * WP + 8 * tj + gr]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int tj = tj0 + u, j = 8 * tj + 2 * gc;
                        const bool skip = tj >= TR || (tj >= ptile && tj < ptile + NB / 8);
                        if (!skip && valid && j < ld) *reinterpret_cast<double2 *>(Mrow + j) = c[u];
                    }
                };
                for (int t0 = 0; t0 < TR; t0 += 16) {
                    double2 c0[8], c1[8];
                    load_strip(t0, c0);
                    if (TR > t0 + 8) load_strip(t0 + 8, c1);
                    finish_strip(t0, c0);
                    if (TR > t0 + 8) finish_strip(t0 + 8, c1);
                }
            }
            __syncthreads();
        }
        const int fail = s_fail;
        if (tid == 0) fail_flags[p] = fail;
        if (!fail) {
            // M[i][j] = B^-1[qinv[i]][perm[j]]  ->  S[a][b] = B^-1[b][a] = M[perm[b]][qinv[a]]:
            // transpose in place (T[a][b] = M[b][a]), then S[a][b] = T[qinv[a]][perm[b]]
            {
                // 32 x 32 tile pairs through shared memory: 16 coalesced loads in flight per thread, coalesced stores
                double *tA = Pbuf, *tB = Pbuf + 32 * 33;                         // two 32 x 33 tiles (17 KB of the 20 KB panel buffer)
                const int nt = (n + 31) >> 5, cx = tid & 31, ry = tid >> 5;      // thread: column cx, rows ry, ry + NWARP, ...
                constexpr int RPT = 32 / NWARP;                                  // rows of a 32 x 32 tile per thread
                for (int ta = 0; ta < nt; ++ta)
                    for (int tb = ta; tb < nt; ++tb) {
                        double va[RPT], vb[RPT];
#pragma unroll
                        for (int k = 0; k < RPT; ++k) {
                            const int r = ry + NWARP * k;
                            const int ia = 32 * ta + r, ja = 32 * tb + cx;       // tile (ta, tb)
                            const int ib = 32 * tb + r, jb = 32 * ta + cx;       // tile (tb, ta)
                            va[k] = (ia < n && ja < n) ? B[(ll)ia * ld + ja] : 0.0;
                            vb[k] = (ib < n && jb < n) ? B[(ll)ib * ld + jb] : 0.0;
                        }
#pragma unroll
                        for (int k = 0; k < RPT; ++k) { tA[(ry + NWARP * k) * 33 + cx] = va[k]; tB[(ry + NWARP * k) * 33 + cx] = vb[k]; }
                        __syncthreads();
#pragma unroll
                        for (int k = 0; k < RPT; ++k) {
                            const int r = ry + NWARP * k;
                            const int ia = 32 * ta + r, ja = 32 * tb + cx, ib = 32 * tb + r, jb = 32 * ta + cx;
                            if (ia < n && ja < n) B[(ll)ia * ld + ja] = tB[cx * 33 + r];     // (ta, tb)[r][cx] = (tb, ta)[cx][r]
                            if (ta != tb && ib < n && jb < n) B[(ll)ib * ld + jb] = tA[cx * 33 + r];
                        }
                        __syncthreads();
                    }
            }
            if (kPivot && s_offdiag) {
                if (i < n) qinv[perm[i]] = i;
                __syncthreads();
                // rows: new row a = old row qinv[a].  Cycle leaders (smallest member of every non-trivial cycle) once, then
                // every thread follows the cycles for its columns
                if (tid == 0) {
                    int nl = 0;
                    for (int s0 = 0; s0 < n; ++s0) {
                        if (qinv[s0] == s0) continue;
                        bool smallest = true;
                        for (int q = qinv[s0]; q != s0; q = qinv[q]) if (q < s0) { smallest = false; break; }
                        if (smallest) leaders[nl++] = s0;
                    }
                    s_nleaders = nl;
                }
                __syncthreads();                                                 // also orders the transpose before the permutations
                for (int col = tid; col < n; col += N) {
                    for (int l = 0; l < s_nleaders; ++l) {
                        const int s0 = leaders[l];
                        // new[a] = old[qinv[a]]: walk a = s0, qinv[s0], ... keeping the first value for the last write
                        const double first = B[(ll)s0 * ld + col];
                        int a = s0;
                        while (true) {
                            const int src = qinv[a];
                            if (src == s0) { B[(ll)a * ld + col] = first; break; }
                            B[(ll)a * ld + col] = B[(ll)src * ld + col];
                            a = src;
                        }
                    }
                }
                __syncthreads();
                // columns: S[a][b] = row[perm[b]], a warp per row through a shared-memory copy of the row
                double *rowbuf = Wbuf + warp * WP;
                for (int a = warp; a < n; a += NWARP) {
                    double *row = B + (ll)a * ld;
                    for (int b = lane; b < n; b += 32) rowbuf[b] = row[b];
                    __syncwarp();
                    for (int b = lane; b < n; b += 32) row[b] = rowbuf[perm[b]];
                    __syncwarp();
                }
            }
        }
        __syncthreads();
    }
}
