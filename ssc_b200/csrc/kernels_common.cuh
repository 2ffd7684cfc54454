// sm_100a kernels of the patch preconditioner: shared declarations (layouts, peer scatter helpers).
// kernels.cuh includes the thematic headers in order so that every launch site sees the same layouts.
//
// Layout in HBM (DESIGN.md "data layout"):
//   * patches are sorted by size n into classes; inside a class patch q has its dof list at
//     gtol[goff + q*n .. +n) (int32) and its dense block at blocks[moff + q*stride .. + n*n) (fp64),
//     stride = n*n rounded up to 16 doubles so every block starts on a 128-byte line.
//   * after setup a block holds S = (B^-1)^T row-major, i.e. S[j*n + i] = (B^-1)[i][j]: in the apply kernel
//     thread i of a patch reads S[j*n + i] for j = 0..n-1, so a warp reads consecutive doubles (coalesced),
//     every byte of the block is read exactly once, and no cross-thread reduction is needed.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef long long ll;

__device__ __forceinline__ unsigned hash_dof(int key) { return (unsigned)key * 2654435761u; }

// Peer-memory scatter (multi-GPU, NVLink P2P): a patch dof >= nowned is a ghost whose owner lives on another GPU of
// the box; its correction is added straight into the owner's local vector through a mapped peer pointer, so the
// overlap sum of ssc/libssc.c:1399-1400 travels over NVLink while the patch kernel is still streaming blocks.
struct PeerScatter {
    double *yl[16];          // "incoming" vector (one entry per owned dof) of every neighbour (cudaIpcOpenMemHandle)
    const int *ghostPeer;    // ghost slot -> neighbour number
    const int *ghostRemote;  // ghost slot -> owned index on that neighbour
    const double *xg;        // this rank's ghost values of x, written by the neighbours (halo fill)
    int nowned;
};
// residual entry of local dof gi: owned ones come from the caller's vector, ghosts from the array the neighbours filled
__device__ __forceinline__ double gather_x(const double *__restrict__ x, int gi, const PeerScatter *__restrict__ ps)
{
    if (ps != nullptr && gi >= ps->nowned) return ps->xg[gi - ps->nowned];
    return __ldg(x + gi);
}
// Deterministic mode: instead of reducing into y, every patch dof stores its correction in its own slot of `ypatch`
// (coalesced plain stores); ssc_slot_sum_kernel then adds the slots of each dof in a fixed order.
__device__ __forceinline__ void scatter_add(double *__restrict__ y, int gi, double v, const PeerScatter *__restrict__ ps);
__device__ __forceinline__ void emit(double *__restrict__ y, int gi, double v, const PeerScatter *__restrict__ ps, double *__restrict__ ypatch, ll slot)
{
    if (ypatch != nullptr) ypatch[slot] = v;
    else scatter_add(y, gi, v, ps);
}
__device__ __forceinline__ void scatter_add(double *__restrict__ y, int gi, double v, const PeerScatter *__restrict__ ps)
{
    if (ps != nullptr && gi >= ps->nowned) {
        const int s = gi - ps->nowned;
        atomicAdd(ps->yl[ps->ghostPeer[s]] + ps->ghostRemote[s], v);
    } else {
        atomicAdd(y + gi, v);
    }
}
