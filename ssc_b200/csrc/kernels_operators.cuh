// Patch operators: extraction from the assembled CSR (K1) and assembly from element matrices (K1e).
#pragma once
#include "kernels_common.cuh"

// ---------------------------------------------------------------------------------------------
// K1 extract: B_p[i][j] = A[gtol_p[i], gtol_p[j]]   (replaces MatZeroEntries + callback assembly,
// ssc/libssc.c:1286-1291, :1120-1156).  One CTA per patch.  A shared-memory hash table (one 64-bit entry = global
// column, patch column; load factor <= 1/4, so a miss -- two thirds of the probes: the columns of a row that lie outside
// the patch -- costs 1.4 probes) maps global column -> patch column.  A warp owns one output row at a time: it has four
// 32-entry chunks of the CSR row (index + value) in flight per lane, drops the hits into a row buffer in shared memory and
// streams the finished row out with coalesced evict-first stores (the blocks are written once and must not push the CSR
// lines, which the neighbouring patches read again, out of L2), zeroing the buffer in the same pass.
// kAccumulate = false: every row of the CSR has strictly increasing columns (checked once on the device when the operator
// is set, ssc_csr_sorted_kernel), so a hit is a plain store; true: duplicates are summed with shared-memory atomics.
// rowbuf_warps = 0: the patch is too large for row buffers; hits go straight to the zero-filled row in global memory.
// ---------------------------------------------------------------------------------------------
__global__ void ssc_csr_sorted_kernel(const ll *__restrict__ rowptr, const int *__restrict__ colidx, ll nrows, int *__restrict__ unsorted)
{
    // a warp per row: columns strictly increasing?
    const ll warp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((ll)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    int bad = 0;
    for (ll r = warp; r < nrows; r += nw) {
        const ll a = rowptr[r], b = rowptr[r + 1];
        for (ll k = a + 1 + lane; k < b; k += 32) bad |= colidx[k] <= colidx[k - 1];
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) *unsorted = 1;
}

template <bool kAccumulate>
__global__ void __launch_bounds__(256)
ssc_extract_kernel(const ll *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
                   const int *__restrict__ gtol, double *__restrict__ blocks, int n, int ld, ll stride, ll count, int H, int rowbuf_warps)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int2 *table = reinterpret_cast<int2 *>(smem_raw);           // H entries: (global column, patch column); x = -1: empty
    ll *rbeg = reinterpret_cast<ll *>(table + H);               // n: start of every patch row in the CSR arrays
    int *rlen = reinterpret_cast<int *>(rbeg + n);              // n (+1 pad): its length
    double *rowbuf = reinterpret_cast<double *>(rlen + ((n + 2) & ~1));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const bool buffered = rowbuf_warps > 0;
    if (buffered) for (int t = tid; t < nwarps * ld; t += blockDim.x) rowbuf[t] = 0.0;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        const int *g = gtol + p * n;
        for (int t = tid; t < H; t += blockDim.x) table[t].x = -1;
        __syncthreads();
        // one round of dependent loads for the whole patch: dof -> row extent; and the column table
        for (int t = tid; t < n; t += blockDim.x) {
            const int key = g[t];
            const ll a = rowptr[key];
            rbeg[t] = a;
            rlen[t] = (int)(rowptr[key + 1] - a);
            unsigned s = hash_dof(key) & (H - 1);
            while (true) {
                const int prev = atomicCAS(&table[s].x, -1, key);
                if (prev == -1 || prev == key) { table[s].y = t; break; }
                s = (s + 1) & (H - 1);
            }
        }
        __syncthreads();
        double *B = blocks + p * stride;
        double *rb = rowbuf + warp * ld;                       // entries n .. ld-1 stay zero
        for (int i = warp; i < n; i += nwarps) {
            const int len = rlen[i];
            const int *ci = colidx + rbeg[i] + lane;            // 64-bit row start once, 32-bit offsets below
            const double *va = vals + rbeg[i] + lane;
            double *Brow = B + (ll)i * ld;
            if (!buffered) {
                for (int t = lane; t < ld; t += 32) Brow[t] = 0.0;
                __syncwarp();
            }
            double *dst = buffered ? rb : Brow;
            for (int k0 = lane; k0 < len + lane; k0 += 128) {   // uniform trip count; four chunks of the row in flight per lane
                int c[4];
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool in = k0 + 32 * u < len;
                    c[u] = in ? __ldg(ci + (k0 - lane) + 32 * u) : -1;
                    v[u] = in ? __ldg(va + (k0 - lane) + 32 * u) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (c[u] < 0) continue;
                    unsigned s = hash_dof(c[u]) & (H - 1);
                    int2 e = table[s];
                    while (e.x != c[u] && e.x != -1) { s = (s + 1) & (H - 1); e = table[s]; }
                    if (e.x == c[u]) {
                        if (kAccumulate) atomicAdd(dst + e.y, v[u]);
                        else dst[e.y] = v[u];
                    }
                }
            }
            if (buffered) {
                __syncwarp();
                for (int t = lane; t < ld; t += 32) {
                    const double val = rb[t];
                    rb[t] = 0.0;
                    __stcs(Brow + t, val);
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K1e assemble: B_p from ELEMENT matrices, the reference's own route to the patch operators
// (PCPatchComputeOperator ssc/libssc.c:1120-1156 + the callback ssc/PatchPC.pyx:54-72, ssc/patch.py:28-62,210-213:
// MatSetValues(ADD_VALUES) of one tdpc x tdpc matrix per patch cell through the per-cell `dofs` table, negative
// indices dropped).  One CTA per patch.  The patch-local index of a cell dof is looked up in a shared-memory hash of
// the patch's dof list (a dof that is not in the list is a global or artificial BC dof: the -1 of the `dofs` table).
// Cells are added one after the other in the order of the `cells` list, so every entry is summed in the
// reference's order and the result is bit-identical to the oracle's (oracle_patch_matrix_from_elements).
// Inside one cell all targets are distinct, so no atomics are needed.
// ---------------------------------------------------------------------------------------------
struct ElemSpaces {
    const int *cnm[8];            // cell-node map of every subspace, rows follow the cell numbering
    int bs[8], npc[8], subOff[8], tOff[9];
    int nsub, t;                  // t = totalDofsPerCell
};

__global__ void __launch_bounds__(1024)
ssc_assemble_kernel(ElemSpaces E, const double *__restrict__ Ke, ll kstride, const ll *__restrict__ celloff, const int *__restrict__ cells,
                    const int *__restrict__ gtol, double *blocks, int n, int ld, ll stride, ll count, int H, int acc_shared)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *keys = reinterpret_cast<int *>(smem_raw);
    int *slotval = keys + H;
    int *lidx = slotval + H;                                   // t: patch-local index of every cell dof (-1: dropped)
    int *vi = lidx + ((E.t + 1) & ~1);                         // t: the cell dofs that are kept
    double *acc_s = reinterpret_cast<double *>(vi + ((E.t + 1) & ~1));
    __shared__ int nvalid[2];                                  // kept-dof counters of even and odd cells
    const int tid = threadIdx.x;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        const int *g = gtol + p * n;
        for (int t = tid; t < H; t += blockDim.x) keys[t] = -1;
        __syncthreads();
        for (int t = tid; t < n; t += blockDim.x) {
            const int key = g[t];
            unsigned s = hash_dof(key) & (H - 1);
            while (true) {
                const int prev = atomicCAS(&keys[s], -1, key);
                if (prev == -1 || prev == key) { slotval[s] = t; break; }
                s = (s + 1) & (H - 1);
            }
        }
        double *B = blocks + p * stride;
        double *acc = acc_shared ? acc_s : B;
        for (int e = tid; e < n * ld; e += blockDim.x) acc[e] = 0.0;
        if (tid == 0) { nvalid[0] = 0; nvalid[1] = 0; }
        __syncthreads();
        int par = 0;
        for (ll c = celloff[p]; c < celloff[p + 1]; ++c, par ^= 1) {
            const ll cell = cells[c];
            for (int t = tid; t < E.t; t += blockDim.x) {
                int k = 0;
                while (t >= E.tOff[k + 1]) ++k;
                const int r = t - E.tOff[k], j = r / E.bs[k], l = r - j * E.bs[k];
                const int key = E.cnm[k][cell * E.npc[k] + j] * E.bs[k] + l + E.subOff[k];
                int loc = -1;
                unsigned s = hash_dof(key) & (H - 1);
                int kk;
                while ((kk = keys[s]) != -1) { if (kk == key) { loc = slotval[s]; break; } s = (s + 1) & (H - 1); }
                lidx[t] = loc;
                if (loc >= 0) vi[atomicAdd(&nvalid[par], 1)] = t;
            }
            __syncthreads();
            const int nv = nvalid[par];
            if (tid == 0) nvalid[par ^ 1] = 0;                 // nobody touches the other counter before the next barrier
            const double *K = Ke + cell * kstride;
            for (int e = tid; e < nv * nv; e += blockDim.x) {
                const int a = e / nv, b = e - a * nv;
                const int i = vi[a], j = vi[b];
                acc[lidx[i] * ld + lidx[j]] += K[i * E.t + j];
            }
            __syncthreads();
        }
        if (acc_shared) {
            for (int e = tid; e < n * ld; e += blockDim.x) B[e] = acc_s[e];
            __syncthreads();
        }
    }
}
