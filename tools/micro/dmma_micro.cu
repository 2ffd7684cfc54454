// DMMA (mma.sync.m8n8k4.f64) throughput on sm_100a with register-resident accumulators, against DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/dmma_micro tools/micro/dmma_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void dmma_tput(double *out, int iters)
{
    double c0[NACC], c1[NACC];
    for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = -c0[i]; }
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x & 7);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c0[i], c1[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// with operand loads from shared memory: 4 A fragments + 4 B fragments per 16 DMMAs (a 32x32 warp block, one k4 step)
__global__ void dmma_smem(double *out, int iters)
{
    __shared__ double As[128 * 4], Bs[4 * 132];
    for (int i = threadIdx.x; i < 128 * 4; i += blockDim.x) As[i] = 1e-9 * i;
    for (int i = threadIdx.x; i < 4 * 132; i += blockDim.x) Bs[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = (warp >> 2) & 3, wc = warp & 3;
    double c0[16], c1[16];
    for (int i = 0; i < 16; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[(32 * wr + 8 * i + (lane >> 2)) * 4 + (lane & 3)];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[(lane & 3) * 132 + 32 * wc + 8 * j + (lane >> 2)];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma(c0[4 * i + j], c1[4 * i + j], a[i], b[j]);
        if (it == iters + 5) { As[threadIdx.x] = c0[0]; __syncthreads(); }
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out; cudaMalloc(&out, 148 * 2048 * sizeof(double) * 2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 4000, blocks = 148;
        dmma_tput<16><<<blocks, threads>>>(out, 10);
        cudaEventRecord(e0);
        dmma_tput<16><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 256 * 16 * iters * (double)blocks * (threads / 32);
        printf("DMMA reg-only, 1 CTA/SM x %4d threads, 16 acc: %.2f TFLOP/s (%.1f FMA/clk/SM at 1.9 GHz)\n", threads, flops / ms * 1e-9, flops / 2 / (ms * 1e-3) / 148 / 1.9e9);
    }
    for (int threads : {256, 512}) {
        const int iters = 4000, blocks = 148;
        dmma_tput<4><<<blocks, threads>>>(out, 10);
        cudaEventRecord(e0);
        dmma_tput<4><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 256 * 4 * iters * (double)blocks * (threads / 32);
        printf("DMMA reg-only, 1 CTA/SM x %4d threads,  4 acc: %.2f TFLOP/s\n", threads, flops / ms * 1e-9);
    }
    for (int threads : {256, 512}) {
        const int iters = 4000, blocks = 148;
        dmma_smem<<<blocks, threads>>>(out, 10);
        cudaEventRecord(e0);
        dmma_smem<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 256 * 16 * iters * (double)blocks * (threads / 32);
        printf("DMMA + smem fragments (8 LDS per 16 DMMA), %4d threads: %.2f TFLOP/s\n", threads, flops / ms * 1e-9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
