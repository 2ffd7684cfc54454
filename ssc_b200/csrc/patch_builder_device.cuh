// Device-side patch construction: the star-patch case of PCPatchCreateCellPatches
// (ssc/libssc.c:452-580) and PCPatchCreateCellPatchDiscretisationInfo (ssc/libssc.c:598-900) as one CUDA kernel,
// one CTA per patch.  Same outputs, bit for bit, as the host builder (patch_builder.cpp) and the oracle:
//   * the owned points are {v} + star(v) (breadth-first over supports), the seen points their closure (over cones);
//   * cells of the patch ascend by mesh point;
//   * a dof of a patch cell is dropped if it is a global Dirichlet dof or "seen but not owned" (artificial BC), and the
//     kept ones are numbered in first-encounter order walking subspaces -> cells -> nodes -> components.
// The first-encounter order is made parallel: every candidate position records itself with atomicMin in a
// shared-memory hash keyed by dof; positions that hold their dof's minimum are the first encounters, and a block-wide
// exclusive scan over positions turns them into local numbers.
// Anything the kernel cannot hold in shared memory (very large patches) raises an overflow flag and the host builder
// takes over; Vanka / user patches, exclude_subspace and the per-cell dofs table always use the host builder.
#pragma once
#include <cuda_runtime.h>

namespace devbuild {

constexpr int kThreads = 128;
constexpr int kMaxPoints = 1024;     // points in the closure of a star
constexpr int kPointHash = 2048;
constexpr int kMaxCells = 256;
constexpr int kDofHash = 4096;       // dofs seen by a patch (<= kDofHash / 2)
constexpr int kMaxSub = 8;

struct Topo {
    const int *coneOff, *cones, *suppOff, *supports;
    const unsigned char *ghost;      // may be null
    const int *cnOff;                // cell numbering: plex cell point -> row of the cell-node maps
    const unsigned char *isBc;       // one byte per local dof: global Dirichlet dof (ghostBcNodes)
    int cStart, cEnd, vStart, npatch, nlocal;
    int nsub;
    const int *secDof[kMaxSub], *secOff[kMaxSub], *cnm[kMaxSub];
    int bs[kMaxSub], npc[kMaxSub], subOff[kMaxSub];
};

__device__ __forceinline__ unsigned hash32(int key, int mask) { return ((unsigned)key * 2654435761u >> 7) & (unsigned)mask; }

// insert key; returns the slot (-1 if the table is full); *fresh tells whether this call created the entry
__device__ __forceinline__ int hset(int *keys, int mask, int key, bool *fresh)
{
    unsigned s = hash32(key, mask);
    for (int probes = 0; probes <= mask; ++probes) {
        const int prev = atomicCAS(&keys[s], -1, key);
        if (prev == -1) { *fresh = true; return (int)s; }
        if (prev == key) { *fresh = false; return (int)s; }
        s = (s + 1) & (unsigned)mask;
    }
    *fresh = false;
    return -1;
}
__device__ __forceinline__ int hfind(const int *keys, int mask, int key)
{
    unsigned s = hash32(key, mask);
    for (int probes = 0; probes <= mask; ++probes) {
        const int k = keys[s];
        if (k == key) return (int)s;
        if (k == -1) return -1;
        s = (s + 1) & (unsigned)mask;
    }
    return -1;
}

// kFill = false: counts only (gtolCount, cellCount).  kFill = true: writes gtol / cells at the given offsets.
template <bool kFill>
__global__ void __launch_bounds__(kThreads)
build_kernel(Topo T, int *__restrict__ gtolCount, int *__restrict__ cellCount, const long long *__restrict__ gtolOff,
             const long long *__restrict__ cellOff, int *__restrict__ gtol, int *__restrict__ cells, int *__restrict__ overflow)
{
    extern __shared__ int sm[];
    int *ptlist = sm;                              // kMaxPoints
    int *pthash = ptlist + kMaxPoints;             // kPointHash
    int *cellsU = pthash + kPointHash;             // kMaxCells (unsorted)
    int *cellsS = cellsU + kMaxCells;              // kMaxCells (sorted)
    int *dkeys = cellsS + kMaxCells;               // kDofHash
    int *dflag = dkeys + kDofHash;                 // kDofHash: bit 0 seen, bit 1 owned
    int *dmin = dflag + kDofHash;                  // kDofHash: first position of the dof
    __shared__ int nList, nCells, nDofs, warpTot[kThreads / 32], bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int p = blockIdx.x; p < T.npatch; p += gridDim.x) {
        const int v = T.vStart + p;
        if (T.ghost && T.ghost[v]) {               // only owned entities get patches, ssc/libssc.c:515-521
            if (!kFill && tid == 0) { gtolCount[p] = 0; cellCount[p] = 0; }
            continue;
        }
        for (int i = tid; i < kPointHash; i += kThreads) pthash[i] = -1;
        for (int i = tid; i < kDofHash; i += kThreads) { dkeys[i] = -1; dflag[i] = 0; dmin[i] = 0x7fffffff; }
        if (tid == 0) { nList = 1; nCells = 0; nDofs = 0; bad = 0; ptlist[0] = v; }
        __syncthreads();
        if (tid == 0) { bool f; hset(pthash, kPointHash - 1, v, &f); }
        __syncthreads();
        // star(v): breadth-first over supports (PCPatchConstruct_Star, ssc/libssc.c:1447-1469)
        int begin = 0, nOwned = 1;
        for (int pass = 0; pass < 2; ++pass) {
            // pass 0: supports from v -> owned points; pass 1: cones from every owned point -> seen points (:325-363)
            const int *off = pass == 0 ? T.suppOff : T.coneOff, *adj = pass == 0 ? T.supports : T.cones;
            begin = 0;
            while (true) {
                const int end = nList;
                __syncthreads();
                if (begin >= end || bad) break;
                for (int idx = begin + tid; idx < end; idx += kThreads) {
                    const int q = ptlist[idx];
                    for (int k = off[q]; k < off[q + 1]; ++k) {
                        const int r = adj[k];
                        bool fresh;
                        if (hset(pthash, kPointHash - 1, r, &fresh) < 0) bad = 1;
                        if (fresh) {
                            const int pos = atomicAdd(&nList, 1);
                            if (pos < kMaxPoints) ptlist[pos] = r; else bad = 1;
                        }
                    }
                }
                __syncthreads();
                begin = end;
            }
            if (pass == 0) nOwned = nList;
            __syncthreads();
        }
        const int np = min(nList, kMaxPoints);
        // cells of the patch, ascending
        for (int idx = tid; idx < np; idx += kThreads) {
            const int q = ptlist[idx];
            if (q >= T.cStart && q < T.cEnd) { const int c = atomicAdd(&nCells, 1); if (c < kMaxCells) cellsU[c] = q; else bad = 1; }
        }
        __syncthreads();
        const int nc = min(nCells, kMaxCells);
        for (int i = tid; i < nc; i += kThreads) {
            const int c = cellsU[i];
            int rank = 0;
            for (int j = 0; j < nc; ++j) rank += cellsU[j] < c;
            cellsS[rank] = c;
        }
        // seen / owned dofs (PCPatchGetPointDofs, ssc/libssc.c:370-416)
        for (int idx = tid; idx < np; idx += kThreads) {
            const int q = ptlist[idx];
            const int fl = idx < nOwned ? 3 : 1;
            for (int k = 0; k < T.nsub; ++k) {
                const int ld = T.secDof[k][q], lo = T.secOff[k][q], bs = T.bs[k];
                for (int j = lo; j < lo + ld; ++j)
                    for (int l = 0; l < bs; ++l) {
                        bool fresh;
                        const int s = hset(dkeys, kDofHash - 1, bs * j + l + T.subOff[k], &fresh);
                        if (s < 0) { bad = 1; continue; }
                        if (fresh && atomicAdd(&nDofs, 1) >= kDofHash / 2) bad = 1;
                        atomicOr(&dflag[s], fl);
                    }
            }
        }
        __syncthreads();
        if (bad) {                                  // uniform: shared flag read after the barrier
            if (tid == 0) *overflow = 1;
            if (!kFill && tid == 0) { gtolCount[p] = 0; cellCount[p] = 0; }
            __syncthreads();
            continue;
        }
        // candidate positions in the reference's walking order (ssc/libssc.c:731-778)
        int total = 0;
        for (int k = 0; k < T.nsub; ++k) total += nc * T.npc[k] * T.bs[k];
        auto dof_at = [&](int pos) -> int {
            int k = 0, base = 0;
            while (true) { const int blk = nc * T.npc[k] * T.bs[k]; if (pos < base + blk) break; base += blk; ++k; }
            const int local = pos - base, per = T.npc[k] * T.bs[k];
            const int ci = local / per, rem = local - ci * per, j = rem / T.bs[k], l = rem - j * T.bs[k];
            const int cell = T.cnOff[cellsS[ci]];
            return T.cnm[k][(long long)cell * T.npc[k] + j] * T.bs[k] + T.subOff[k] + l;
        };
        for (int pos = tid; pos < total; pos += kThreads) {
            const int g = dof_at(pos);
            if (g < 0 || g >= T.nlocal) { bad = 1; continue; }        // invalid cell-node map: let the host builder report it
            int s = hfind(dkeys, kDofHash - 1, g);
            const int fl = s >= 0 ? dflag[s] : 0;
            const bool keep = !T.isBc[g] && !((fl & 1) && !(fl & 2));
            if (keep) {
                if (s < 0) { bool fresh; s = hset(dkeys, kDofHash - 1, g, &fresh); if (fresh && atomicAdd(&nDofs, 1) >= kDofHash / 2) bad = 1; }
                if (s < 0) { bad = 1; continue; }
                atomicMin(&dmin[s], pos);
            }
        }
        __syncthreads();
        if (bad) {
            if (tid == 0) *overflow = 1;
            if (!kFill && tid == 0) { gtolCount[p] = 0; cellCount[p] = 0; }
            __syncthreads();
            continue;
        }
        // first encounters -> local numbers by an exclusive scan over positions
        int base = 0;
        const long long goff = kFill ? gtolOff[p] : 0;
        for (int chunk = 0; chunk < total; chunk += kThreads) {
            const int pos = chunk + tid;
            int g = -1;
            bool first = false;
            if (pos < total) {
                g = dof_at(pos);
                const int s = hfind(dkeys, kDofHash - 1, g);
                const int fl = s >= 0 ? dflag[s] : 0;
                const bool keep = !T.isBc[g] && !((fl & 1) && !(fl & 2));
                first = keep && s >= 0 && dmin[s] == pos;
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, first);
            if (lane == 0) warpTot[warp] = __popc(ballot);
            __syncthreads();
            int before = 0, all = 0;
            for (int w = 0; w < kThreads / 32; ++w) { if (w < warp) before += warpTot[w]; all += warpTot[w]; }
            if (kFill && first) gtol[goff + base + before + __popc(ballot & ((1u << lane) - 1u))] = g;
            base += all;
            __syncthreads();
        }
        if (!kFill && tid == 0) { gtolCount[p] = base; cellCount[p] = nc; }
        if (kFill) { const long long co = cellOff[p]; for (int i = tid; i < nc; i += kThreads) cells[co + i] = T.cnOff[cellsS[i]]; }
        __syncthreads();
    }
}

constexpr size_t kSmemBytes = (size_t)(kMaxPoints + kPointHash + 2 * kMaxCells + 3 * kDofHash) * sizeof(int);

__global__ void narrow_kernel(long long n, const long long *__restrict__ src, int *__restrict__ dst)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = (int)src[i];
}

}  // namespace devbuild
