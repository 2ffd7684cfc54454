// Micro-measurements behind the factor kernel design: fp64 FMA throughput and latency, division, redux, barrier.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp64_micro tools/micro/fp64_micro.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_tput(double *out, int iters)
{
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void latencies(long long *res, double *sink, double seed)
{
    __shared__ double sh[64];
    double x = seed;
    long long t0, t1;
    // dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = fma(x, 1.0000001, 1e-9);
    t1 = clock64();
    if (threadIdx.x == 0) res[0] = (t1 - t0) / 64;
    // dependent division chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) x = 1.0 / (x + 1.5);
    t1 = clock64();
    if (threadIdx.x == 0) res[1] = (t1 - t0) / 16;
    // redux chain
    unsigned u = (unsigned)threadIdx.x + (unsigned)x;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) u = __reduce_max_sync(0xffffffffu, u + threadIdx.x);
    t1 = clock64();
    if (threadIdx.x == 0) res[2] = (t1 - t0) / 16;
    // shuffle pair chain (double)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64();
    if (threadIdx.x == 0) res[3] = (t1 - t0) / 16;
    // __syncthreads round trip
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) res[4] = (t1 - t0) / 16;
    // shared store -> barrier -> load
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) { sh[(threadIdx.x + i) & 63] = x; __syncthreads(); x += sh[(threadIdx.x + i + 1) & 63]; __syncthreads(); }
    t1 = clock64();
    if (threadIdx.x == 0) res[5] = (t1 - t0) / 16;
    // block fence
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) { sh[threadIdx.x & 63] = x; __threadfence_block(); }
    t1 = clock64();
    if (threadIdx.x == 0) res[6] = (t1 - t0) / 16;
    sink[threadIdx.x] = x + u;
}

int main()
{
    double *out; long long *res;
    cudaMalloc(&out, 148 * 16 * 1024 * sizeof(double));
    cudaMallocManaged(&res, 8 * sizeof(long long));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 20000, blocks = 148 * (2048 / threads);
        dfma_tput<<<blocks, threads>>>(out, 100);
        cudaEventRecord(e0);
        dfma_tput<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * iters * (double)blocks * threads;
        printf("DFMA throughput, %4d threads x %3d blocks: %.2f TFLOP/s (%.1f FMA/clk/SM at 1.9 GHz)\n", threads, blocks, flops / ms * 1e-9, flops / 2 / (ms * 1e-3) / 148 / 1.9e9);
    }
    for (int threads : {32, 448, 512}) {
        latencies<<<1, threads>>>(res, out, 1.25);
        cudaDeviceSynchronize();
        printf("%3d threads: dependent DFMA %lld clk, 1/x %lld clk, redux.max %lld clk, shfl(double)+add %lld clk, __syncthreads %lld clk, st-bar-ld-bar %lld clk, st+fence.cta %lld clk\n",
               threads, res[0], res[1], res[2], res[3], res[4], res[5], res[6]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
