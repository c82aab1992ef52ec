// Integer-pipe microbenchmark for sm_100a: issue rates of the multiply flavours the NTT and MAC kernels are made of.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/pipe_probe tools/probe/pipe_probe.cu
// Prints warp-instructions per clock per SM for each flavour (8 independent chains per thread, 1024 threads per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024, 1) probe(uint32_t *sink, int iters, long long *clk) {
    uint32_t a[8], b = threadIdx.x * 2654435761u + 12345u, c = blockIdx.x + 7u;
    uint64_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i * 977u; w[i] = a[i]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 4) asm volatile("{.reg .u32 t; mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;}" : "+r"(a[i]), "+r"(a[(i + 1) & 7]) : "r"(b), "r"(c));
            if (OP == 5) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b) : "r"(a[i])); }
            if (OP == 6) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
            if (OP == 7) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    if (s == 0x12345678u) sink[0] = s + b;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter) {
    uint32_t *sink; long long *clk;
    cudaMalloc(&sink, 4); cudaMalloc(&clk, 148 * 8);
    const int iters = 4096;
    probe<OP><<<148, 1024>>>(sink, 16, clk);
    probe<OP><<<148, 1024>>>(sink, iters, clk);
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    const double warp_instr = 32.0 * iters * 8 * per_iter;   // per SM: 32 warps
    printf("%-28s %8.3f warp-instr/clk/SM  (%.2f clk per warp-instr per SMSP)\n", name, warp_instr / avg, avg / (warp_instr / 4));
    cudaFree(sink); cudaFree(clk);
}

int main() {
    run<0>("IMAD (mad.lo.u32)", 1);
    run<1>("IMAD.HI (mad.hi.u32)", 1);
    run<7>("IMAD.HI (mul.hi.u32)", 1);
    run<2>("IMAD.WIDE.U32 (+64b acc)", 1);
    run<4>("mad.lo.cc + madc.hi pair", 1);
    run<3>("IADD3 (add.u32)", 1);
    run<6>("SHF (funnel shift)", 1);
    run<5>("IMAD + IADD3 interleaved", 2);
    return 0;
}
