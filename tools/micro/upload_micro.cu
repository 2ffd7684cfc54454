// Pageable host -> device through a ring of pinned buffers filled by T host threads: which T / chunk / ring depth reaches the PCIe rate?
// usage: upload_micro <GB> ; prints a table.  Build: nvcc -O2 -o upload_micro upload_micro.cu -lpthread
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include <condition_variable>
#include <mutex>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static double run(char *dst, const char *src, size_t bytes, int T, size_t chunk, int R, bool pool)
{
    std::vector<void *> stage(R); std::vector<cudaEvent_t> ev(R); std::vector<bool> busy(R, false);
    for (int r = 0; r < R; ++r) { cudaMallocHost(&stage[r], chunk); cudaEventCreateWithFlags(&ev[r], cudaEventDisableTiming); memset(stage[r], 0, chunk); }
    cudaStream_t s; cudaStreamCreate(&s);
    double t0 = now();
    if (!pool) {
        size_t off = 0;
        for (int it = 0; off < bytes; ++it) {
            int r = it % R; size_t len = std::min(chunk, bytes - off);
            if (busy[r]) cudaEventSynchronize(ev[r]);
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) { size_t a = len * t / T, b = len * (t + 1) / T; th.emplace_back([=]() { memcpy((char *)stage[r] + a, src + off + a, b - a); }); }
            for (auto &x : th) x.join();
            cudaMemcpyAsync(dst + off, stage[r], len, cudaMemcpyHostToDevice, s); cudaEventRecord(ev[r], s); busy[r] = true; off += len;
        }
    } else {
        // persistent threads: each owns a 1/T slice of every chunk; the last one to finish a chunk issues its DMA (in chunk order);
        // a thread starts chunk c once the DMA of chunk c-R (same buffer) has been issued and has completed
        const size_t nch = (bytes + chunk - 1) / chunk;
        std::vector<std::atomic<int>> done(nch);
        for (auto &d : done) d = 0;
        std::atomic<long> issued(0);
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t]() {
            for (size_t c = 0; c < nch; ++c) {
                const size_t off = c * chunk, len = std::min(chunk, bytes - off); const int r = (int)(c % R);
                if (c >= (size_t)R) { while (issued.load(std::memory_order_acquire) < (long)(c - R + 1)) std::this_thread::yield(); cudaEventSynchronize(ev[r]); }
                const size_t a = len * t / T, b = len * (t + 1) / T;
                memcpy((char *)stage[r] + a, src + off + a, b - a);
                if (done[c].fetch_add(1) == T - 1) {
                    while (issued.load(std::memory_order_acquire) != (long)c) std::this_thread::yield();
                    cudaMemcpyAsync(dst + off, stage[r], len, cudaMemcpyHostToDevice, s); cudaEventRecord(ev[r], s);
                    issued.store((long)c + 1, std::memory_order_release);
                }
            }
        });
        for (auto &x : th) x.join();
    }
    cudaStreamSynchronize(s);
    double t1 = now();
    for (int r = 0; r < R; ++r) { cudaFreeHost(stage[r]); cudaEventDestroy(ev[r]); }
    cudaStreamDestroy(s);
    return t1 - t0;
}

int main(int argc, char **argv)
{
    const double gb = argc > 1 ? atof(argv[1]) : 4.0;
    const size_t bytes = (size_t)(gb * (1 << 30));
    char *src = (char *)malloc(bytes);
    { std::vector<std::thread> th; for (int t = 0; t < 16; ++t) th.emplace_back([=]() { memset(src + bytes * t / 16, t + 1, bytes * (t + 1) / 16 - bytes * t / 16); }); for (auto &x : th) x.join(); }
    char *dst; cudaMalloc(&dst, bytes);
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
    { double t0 = now(); cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice); printf("plain cudaMemcpy pageable: %.1f GB/s\n", bytes / (now() - t0) / 1e9); }
    { // host memcpy rate alone
        char *tmp = (char *)malloc(bytes); memset(tmp, 0, bytes);
        for (int T : {1, 4, 8, 16, 32}) { double t0 = now(); std::vector<std::thread> th; for (int t = 0; t < T; ++t) th.emplace_back([=]() { memcpy(tmp + bytes * t / T, src + bytes * t / T, bytes * (t + 1) / T - bytes * t / T); }); for (auto &x : th) x.join(); printf("host memcpy T=%d: %.1f GB/s\n", T, bytes / (now() - t0) / 1e9); }
        free(tmp);
    }
    { double t0 = now(); cudaHostRegister(src, bytes, cudaHostRegisterDefault); double t1 = now(); cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice); double t2 = now(); cudaHostUnregister(src); double t3 = now();
      printf("cudaHostRegister %.3f s (%.1f GB/s), copy %.1f GB/s, unregister %.3f s\n", t1 - t0, bytes / (t1 - t0) / 1e9, bytes / (t2 - t1) / 1e9, t3 - t2); }
    for (int pool = 0; pool < 2; ++pool)
        for (size_t mb : {16, 32, 64, 128})
            for (int T : {4, 8, 16, 32})
                for (int R : {3, 4}) {
                    if (pool && R == 3) continue;
                    double s = run(dst, src, bytes, T, mb << 20, R, pool);
                    printf("pool=%d chunk=%zu MB T=%d R=%d: %.1f GB/s\n", pool, mb, T, R, bytes / s / 1e9);
                    fflush(stdout);
                }
    return 0;
}
