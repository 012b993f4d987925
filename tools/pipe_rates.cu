// Micro-benchmark: issue rate (warp instructions per clock per SM) of the integer instructions the hash
// kernel is made of, on the device it runs on.  nvcc -arch=sm_100a -O3 tools/pipe_rates.cu -o /tmp/pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed * (i + 1) + threadIdx.x; b[i] = seed ^ (i * 77u + 1u); }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, 0x9E3779B9, 0x96;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, 0x2745937f, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 3) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, 0x2745937f;" : "=l"(w) : "r"(a[i])); a[i] = (uint32_t)w ^ (uint32_t)(w >> 32); }
            if (OP == 4) asm volatile("mul.hi.u32 %0, %0, 0x80000001;" : "+r"(a[i]));
            if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 7) { asm volatile("lop3.b32 %0, %0, %1, 0x9E3779B9, 0x96;" : "+r"(a[i]) : "r"(b[i])); asm volatile("mad.lo.u32 %0, %0, 0x2745937f, %1;" : "+r"(b[i]) : "r"(a[i])); }
            if (OP == 8) asm volatile("shr.u32 %0, %0, 1;" : "+r"(a[i]));
            if (OP == 9) { uint32_t p; asm volatile("{.reg .pred q; setp.gt.u32 q, %1, %2; selp.u32 %0, %1, %2, q;}" : "=r"(p) : "r"(a[i]), "r"(b[i])); a[i] = p + 1; }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, int per_iter) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * sms);
    k<OP><<<sms, 1024>>>(out, 12345u, cyc); cudaDeviceSynchronize();
    k<OP><<<sms, 1024>>>(out, 12345u, cyc); cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    const double warp_instr = 32.0 * ITERS * 8 * per_iter;   // 32 warps per SM
    printf("%-28s %.3f warp-instr/clk/SM  (%.3f per SMSP)\n", name, warp_instr / avg, warp_instr / avg / 4);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("LOP3", 1); run<1>("SHF.L.W (funnel)", 1); run<2>("IMAD (mad.lo)", 1); run<3>("IMAD.WIDE.U32 (+xor)", 2);
    run<4>("IMAD.HI.U32", 1); run<5>("PRMT", 1); run<6>("IADD", 1); run<7>("LOP3 + IMAD interleaved", 2);
    run<8>("SHR (SHF.R)", 1); run<9>("ISETP+SEL+IADD", 3);
    return 0;
}
