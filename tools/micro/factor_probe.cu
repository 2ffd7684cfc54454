// Stand-alone timing harness of the DMMA factor kernels: the L2-resident one that ships (ssc_b200/csrc/kernels_factor_l2.cuh, kernel = 4)
// and the register-resident one that was measured and rejected (tools/micro/factor_reg_dmma.cuh, kernel = 2).
// usage: factor_probe [n] [count] [pivot 0|1] [kernel 2|3|4] [CTAs per SM of kernel 3 / 4] [panel width of kernel 4: 16|32] [row stride]   (3: the blocked shared-memory-panel kernel of kernels_factor.cuh)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo [-DSSC_DMMA_PROBE=k] -o tools/micro/factor_probe tools/micro/factor_probe.cu
//   SSC_DMMA_PROBE = 0 the kernel as shipped; 1 tensor-core updates skipped (what the pivot chain alone costs);
//                    2 panel steps skipped (what the updates and the data movement alone cost).  Results of 1 and 2 are garbage by design.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../ssc_b200/csrc/kernels_factor_l2.cuh"
#include "../../ssc_b200/csrc/kernels_factor.cuh"
#include "factor_reg_dmma.cuh"

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 125;
    const long long count = argc > 2 ? atoll(argv[2]) : 148 * 40;
    const int pivot = argc > 3 ? atoi(argv[3]) : 1;
    const int kernel = argc > 4 ? atoi(argv[4]) : 2;      // 2 register-resident, 4 L2-resident
    const int ctas = argc > 5 ? atoi(argv[5]) : 4;
    const int nbarg = argc > 6 ? atoi(argv[6]) : 16;     // panel width of kernel 4 (32: n <= 128, pivoting only)
    const int ld = argc > 7 ? atoi(argv[7]) : ((n + 1) & ~1);      // row stride (even; default: the library's n rounded up to even)
    const long long stride = ((long long)n * ld + 15) / 16 * 16;
    std::vector<double> h((size_t)stride * count, 0.0);
    srand(1);
    for (long long p = 0; p < count; ++p)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) h[(size_t)(p * stride + (long long)i * ld + j)] = (i == j ? 0.5 * n : 0.0) + rand() / (double)RAND_MAX - 0.5;
    double *d, *d0; int *fail;
    cudaMalloc(&d, h.size() * sizeof(double)); cudaMalloc(&d0, h.size() * sizeof(double)); cudaMalloc(&fail, count * sizeof(int));
    cudaMemcpy(d0, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = DmmaFactorSmem<4>::bytes;
    cudaFuncSetAttribute(ssc_invert_dmma_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(ssc_invert_dmma_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemcpy(d, d0, h.size() * sizeof(double), cudaMemcpyDeviceToDevice);
        cudaEventRecord(e0);
        if (kernel == 4) {
#define RUN_L2(NV)                                                                                                          \
    do {                                                                                                                    \
        const size_t sm4 = (size_t)(NV * 20 + 16 * (NV + 4)) * sizeof(double);                                               \
        cudaFuncSetAttribute(ssc_invert_l2_kernel<true, 16, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4);     \
        cudaFuncSetAttribute(ssc_invert_l2_kernel<true, 16, NV>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);       \
        cudaFuncSetAttribute(ssc_invert_l2_kernel<false, 16, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4);    \
        cudaFuncSetAttribute(ssc_invert_l2_kernel<false, 16, NV>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);      \
        if (pivot) ssc_invert_l2_kernel<true, 16, NV><<<148 * ctas, NV, sm4>>>(d, n, ld, stride, count, fail);                \
        else ssc_invert_l2_kernel<false, 16, NV><<<148 * ctas, NV, sm4>>>(d, n, ld, stride, count, fail);                     \
    } while (0)
            if (nbarg == 32 && n <= 128) {
                const size_t sm4 = (size_t)(128 * 36 + 32 * 132) * sizeof(double);
                cudaFuncSetAttribute(ssc_invert_l2_kernel<true, 32, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4);
                cudaFuncSetAttribute(ssc_invert_l2_kernel<true, 32, 128>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                ssc_invert_l2_kernel<true, 32, 128><<<148 * ctas, 128, sm4>>>(d, n, ld, stride, count, fail);
            } else if (n <= 128) RUN_L2(128); else if (n <= 256) RUN_L2(256); else RUN_L2(512);
#undef RUN_L2
        } else if (kernel == 3) {
            const int npad8 = (n + 7) & ~7, WPb = ((npad8 + 15) & ~15) + 4, nb = 16;
            const size_t smem_b = ((size_t)npad8 * (nb + 1) + (size_t)nb * WPb + nb + 32) * sizeof(double) + (size_t)(34 + 2 * n + 2) * sizeof(int);
            cudaFuncSetAttribute(ssc_invert_blocked_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
            cudaFuncSetAttribute(ssc_invert_blocked_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
            if (pivot) ssc_invert_blocked_kernel<true, 16><<<148 * ctas, 512, smem_b>>>(d, n, ld, stride, count, fail);
            else ssc_invert_blocked_kernel<false, 16><<<148 * ctas, 512, smem_b>>>(d, n, ld, stride, count, fail);
        } else if (pivot) ssc_invert_dmma_kernel<true, 4><<<148, 640, smem>>>(d, n, ld, stride, count, fail);
        else ssc_invert_dmma_kernel<false, 4><<<148, 640, smem>>>(d, n, ld, stride, count, fail);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    std::vector<int> hf(count);
    cudaMemcpy(hf.data(), fail, count * sizeof(int), cudaMemcpyDeviceToHost);
    long long nf = 0; for (int f : hf) nf += f;
    // residual of patch 0
    std::vector<double> S((size_t)n * ld);
    cudaMemcpy(S.data(), d, (size_t)n * ld * sizeof(double), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double a = 0; for (int k = 0; k < n; ++k) a += h[(size_t)i * ld + k] * S[(size_t)j * ld + k]; a -= (i == j); if (a < 0) a = -a; if (a > worst) worst = a; }
    const double per = best * 1e3 / (count / 148.0);
    printf("kernel %d ctas/SM %d probe %d  n %d  count %lld  pivot %d: %.3f ms  = %.2f us per patch per SM = %.0f clk at 1.9 GHz  (%.2f TFLOP/s of 2n^3)  failed %lld  max|B S^T - I| %.2e  %s\n",
           kernel, ctas, SSC_DMMA_PROBE, n, count, pivot, best, per, per * 1900, 2.0 * n * n * n * count / (best * 1e-3) * 1e-12, nf, worst, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
