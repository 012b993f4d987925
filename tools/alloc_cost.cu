// alloc_cost.cu -- what a sketcher handle's set-up calls cost on this box, alone and from T threads at once
// (fb2_sketch_files creates one handle per worker thread).  nvcc -O2 -o alloc_cost alloc_cost.cu; ./alloc_cost [threads]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void spin_kernel(int n, int *out) { int x = 0; for (int i = 0; i < n; ++i) x += i * i; if (x == 42) *out = x; }
static void one(int T, int tid, double *res) {
    cudaSetDevice(0);
    void *p; double t;
    t = now_ms(); cudaHostAlloc(&p, 32u << 20, cudaHostAllocDefault); res[0] = now_ms() - t; void *h32 = p;
    t = now_ms(); cudaHostAlloc(&p, 6u << 20, cudaHostAllocDefault); res[1] = now_ms() - t; void *h6 = p;
    t = now_ms(); for (int i = 0; i < 4; ++i) { cudaHostAlloc(&p, 4096, cudaHostAllocDefault); } res[2] = (now_ms() - t) / 4;
    std::vector<void *> d;
    t = now_ms(); for (int i = 0; i < 5; ++i) { cudaMalloc(&p, 9u << 20); d.push_back(p); } res[3] = (now_ms() - t) / 5;
    t = now_ms(); for (int i = 0; i < 20; ++i) { cudaMalloc(&p, 64u << 10); d.push_back(p); } res[4] = (now_ms() - t) / 20;
    t = now_ms(); for (int i = 0; i < 4; ++i) { cudaMalloc(&p, 64u << 20); d.push_back(p); } res[5] = (now_ms() - t) / 4;
    cudaStream_t st[3];
    t = now_ms(); for (int i = 0; i < 3; ++i) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking); res[6] = (now_ms() - t) / 3;
    int *o; cudaMalloc(&o, 4);
    for (int i = 0; i < 20; ++i) spin_kernel<<<148, 256, 0, st[0]>>>(20000, o);   // ~ a file's worth of kernels in flight
    t = now_ms(); for (void *q : d) cudaFree(q); res[7] = (now_ms() - t) / (double)d.size();
    t = now_ms(); cudaFreeHost(h32); cudaFreeHost(h6); res[8] = (now_ms() - t) / 2;
    cudaStreamSynchronize(st[0]);
    for (int i = 0; i < 3; ++i) cudaStreamDestroy(st[i]);
    cudaFree(o);
    (void)T; (void)tid;
}
int main(int argc, char **argv) {
    cudaFree(0);
    const char *names[9] = {"cudaHostAlloc 32 MiB", "cudaHostAlloc 6 MiB", "cudaHostAlloc 4 KiB", "cudaMalloc 9 MiB", "cudaMalloc 64 KiB", "cudaMalloc 64 MiB",
                            "cudaStreamCreate", "cudaFree (kernels of other threads in flight)", "cudaFreeHost"};
    for (int T : {1, 8, 16}) {
        if (argc > 1 && atoi(argv[1]) != T) continue;
        std::vector<double> res(9 * T, 0.0);
        std::vector<std::thread> th;
        const double t0 = now_ms();
        for (int t = 0; t < T; ++t) th.emplace_back(one, T, t, &res[9 * t]);
        for (auto &x : th) x.join();
        printf("%d thread(s), %.1f ms wall; mean ms per call:\n", T, now_ms() - t0);
        for (int k = 0; k < 9; ++k) { double s = 0; for (int t = 0; t < T; ++t) s += res[9 * t + k]; printf("  %-46s %8.3f\n", names[k], s / T); }
    }
    return 0;
}
