// Patch operators: extraction from the assembled CSR (K1) and assembly from element matrices (K1e).
#pragma once
#include "kernels_common.cuh"

// ---------------------------------------------------------------------------------------------
// K1 extract: B_p[i][j] = A[gtol_p[i], gtol_p[j]]   (replaces MatZeroEntries + callback assembly,
// ssc/libssc.c:1286-1291, :1120-1156).  One CTA per patch; a shared-memory hash maps global column -> patch
// column; each warp owns one output row at a time, accumulates it in shared memory and stores it coalesced.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ssc_extract_kernel(const ll *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
                   const int *__restrict__ gtol, double *__restrict__ blocks, int n, int ld, ll stride, ll count, int H)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *keys = reinterpret_cast<int *>(smem_raw);
    int *slotval = keys + H;
    ll *rbeg = reinterpret_cast<ll *>(slotval + H);            // n: start of every patch row in the CSR arrays
    int *rlen = reinterpret_cast<int *>(rbeg + n);            // n (+1 pad): its length
    double *rowbuf = reinterpret_cast<double *>(rlen + ((n + 2) & ~1));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (ll p = blockIdx.x; p < count; p += gridDim.x) {
        const int *g = gtol + p * n;
        for (int t = tid; t < H; t += blockDim.x) keys[t] = -1;
        __syncthreads();
        // one round of dependent loads for the whole patch: dof -> row extent; and the column hash
        for (int t = tid; t < n; t += blockDim.x) {
            const int key = g[t];
            const ll a = rowptr[key];
            rbeg[t] = a;
            rlen[t] = (int)(rowptr[key + 1] - a);
            unsigned s = hash_dof(key) & (H - 1);
            while (true) {
                const int prev = atomicCAS(&keys[s], -1, key);
                if (prev == -1 || prev == key) { slotval[s] = t; break; }
                s = (s + 1) & (H - 1);
            }
        }
        __syncthreads();
        double *B = blocks + p * stride;
        // a warp owns two output rows at a time: both rows' index/value loads are issued before either is consumed
        double *rb0 = rowbuf + (2 * warp) * n, *rb1 = rb0 + n;
        for (int i0 = 2 * warp; i0 < n; i0 += 2 * nwarps) {
            const int i1 = i0 + 1;
            const bool two = i1 < n;
            for (int t = lane; t < n; t += 32) { rb0[t] = 0.0; rb1[t] = 0.0; }
            __syncwarp();
            const ll a0 = rbeg[i0], a1 = two ? rbeg[i1] : 0;
            const int l0 = rlen[i0], l1 = two ? rlen[i1] : 0;
            const int lmax = l0 > l1 ? l0 : l1;
            for (int k = lane; k < lmax; k += 32) {
                int c0 = -1, c1 = -1;
                double v0 = 0.0, v1 = 0.0;
                if (k < l0) { c0 = colidx[a0 + k]; v0 = vals[a0 + k]; }
                if (k < l1) { c1 = colidx[a1 + k]; v1 = vals[a1 + k]; }
                if (c0 >= 0) {
                    unsigned s = hash_dof(c0) & (H - 1);
                    int kk;
                    while ((kk = keys[s]) != -1) { if (kk == c0) { atomicAdd(&rb0[slotval[s]], v0); break; } s = (s + 1) & (H - 1); }
                }
                if (c1 >= 0) {
                    unsigned s = hash_dof(c1) & (H - 1);
                    int kk;
                    while ((kk = keys[s]) != -1) { if (kk == c1) { atomicAdd(&rb1[slotval[s]], v1); break; } s = (s + 1) & (H - 1); }
                }
            }
            __syncwarp();
            for (int t = lane; t < ld; t += 32) {
                B[(ll)i0 * ld + t] = t < n ? rb0[t] : 0.0;
                if (two) B[(ll)i1 * ld + t] = t < n ? rb1[t] : 0.0;
            }
            __syncwarp();
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
