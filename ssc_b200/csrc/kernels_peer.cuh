// NVLink peer exchange kernels.
#pragma once
#include "kernels_common.cuh"

// ---------------------------------------------------------------------------------------------
// NVLink peer exchange (one process per GPU, buffers mapped with CUDA IPC)
// ---------------------------------------------------------------------------------------------
struct PeerPush { double *xl[16]; };
// halo fill (SFBcast, ssc/libssc.c:1323-1324) as direct stores into the neighbours' ghost slots
__global__ void ssc_peer_push_kernel(ll n, const int *__restrict__ sendRoot, const int *__restrict__ sendPeer, const int *__restrict__ remoteLeaf,
                                     const double *__restrict__ x, PeerPush peers)
{
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x)
        peers.xl[sendPeer[i]][remoteLeaf[i]] = x[sendRoot[i]];
}
struct PeerFlags { unsigned long long *flags[16]; int slot[16]; };
// Neighbour barrier: thread k tells neighbour k "everything this rank enqueued before epoch e is done" and waits for
// the same word from it.  Each rank runs on its own GPU; a bounded spin turns a lost peer into an error flag, not a hang.
__global__ void ssc_peer_barrier_kernel(int nneigh, PeerFlags peers, volatile unsigned long long *mine, unsigned long long epoch, int *timeout)
{
    const int k = threadIdx.x;
    if (k >= nneigh) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flags[k] + peers.slot[k]), "l"(epoch) : "memory");
    const long long t0 = clock64();
    unsigned long long v;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + k) : "memory");
        if (clock64() - t0 > 8000000000LL) { *timeout = 1; break; }      // ~4 s
    } while (v < epoch);
    __threadfence_system();
}

// The two halves of the neighbour barrier as separate launches: a rank can announce "my patches that touch other
// ranks' dofs are done" early and collect the neighbours' announcements late (pipelined host-space apply).
__global__ void ssc_peer_signal_kernel(int nneigh, PeerFlags peers, unsigned long long epoch)
{
    const int k = threadIdx.x;
    if (k >= nneigh) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flags[k] + peers.slot[k]), "l"(epoch) : "memory");
}
__global__ void ssc_peer_wait_kernel(int nneigh, volatile unsigned long long *mine, unsigned long long epoch, int *timeout)
{
    const int k = threadIdx.x;
    if (k >= nneigh) return;
    const long long t0 = clock64();
    unsigned long long v;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + k) : "memory");
        if (clock64() - t0 > 8000000000LL) { *timeout = 1; break; }      // ~4 s
    } while (v < epoch);
    __threadfence_system();
}

// r = x - A y, one warp per row (residual between the colours of a multiplicative sweep)
__global__ void ssc_residual_kernel(ll nrows, const ll *__restrict__ rowptr, const int *__restrict__ colidx, const double *__restrict__ vals,
                                    const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ r)
{
    const ll warp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    double s = 0.0;
    for (ll k = rowptr[warp] + lane; k < rowptr[warp + 1]; k += 32) s = fma(vals[k], y[colidx[k]], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) r[warp] = x[warp] - s;
}

// deterministic mode: y[g] = sum of the slots of dof g, in ascending slot order (no atomics)
__global__ void ssc_slot_sum_kernel(ll n, const ll *__restrict__ off, const int *__restrict__ slots, const double *__restrict__ ypatch, double *__restrict__ y)
{
    for (ll g = (ll)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (ll)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (ll k = off[g]; k < off[g + 1]; ++k) s += ypatch[slots[k]];
        y[g] = s;
    }
}
