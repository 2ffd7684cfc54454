// L2-resident read+write streaming bandwidth on B200: each CTA repeatedly updates (x = x*a + b) its own 128 KB slice of a buffer that fits L2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/l2_micro tools/micro/l2_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rw(double2 *buf, int per_cta, int iters)
{
    double2 *p = buf + (size_t)blockIdx.x * per_cta;
    for (int it = 0; it < iters; ++it)
        for (int i = threadIdx.x; i < per_cta; i += blockDim.x) { double2 v = p[i]; v.x = v.x * 1.0000001 + 1e-9; v.y = v.y * 0.9999999 - 1e-9; p[i] = v; }
}
__global__ void ro(const double2 *buf, int per_cta, int iters, double *out)
{
    const double2 *p = buf + (size_t)blockIdx.x * per_cta;
    double s = 0;
    for (int it = 0; it < iters; ++it)
        for (int i = threadIdx.x; i < per_cta; i += blockDim.x) { double2 v = p[i]; s += v.x + v.y; }
    if (s == 1.2345) out[0] = s;
}
int main()
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double *out; cudaMalloc(&out, 8);
    for (int ctas_per_sm : {1, 2, 4}) for (int kb : {64, 128}) {
        const int ctas = 148 * ctas_per_sm, per = kb * 1024 / 16;
        const size_t total = (size_t)ctas * per * 16;
        double2 *buf; cudaMalloc(&buf, total); cudaMemset(buf, 0, total);
        const int iters = 200;
        for (int mode = 0; mode < 2; ++mode) {
            if (mode) rw<<<ctas, 256>>>(buf, per, 2); else ro<<<ctas, 256>>>(buf, per, 2, out);
            cudaEventRecord(e0);
            if (mode) rw<<<ctas, 256>>>(buf, per, iters); else ro<<<ctas, 256>>>(buf, per, iters, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%s  %d CTA/SM x %3d KB = %6.1f MB working set: %.2f TB/s (%s)\n", mode ? "read+write" : "read only ", ctas_per_sm, kb, total / 1e6,
                   (mode ? 2.0 : 1.0) * total * iters / (ms * 1e-3) / 1e12, mode ? "bytes read + written" : "bytes read");
        }
        cudaFree(buf);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
