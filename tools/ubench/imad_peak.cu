// Integer-multiply issue-rate micro-benchmark for sm_100a (B200).
// Measures, per SM and per clock, how many 32x32 products the integer pipe retires for the
// instruction forms a multi-precision Montgomery multiply can be built from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o imad_peak imad_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NCH 8   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t seed, unsigned long long *cyc)
{
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t lo[NCH], hi[NCH], x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { lo[i] = a + i; hi[i] = b ^ i; x[i] = a * (2 * i + 3) + seed; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            b += 0x9e3779b9u;   // new multiplier every row (like b[i] / m in a Montgomery row)
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                if (MODE == 0) {          // IMAD (lo)
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(x[i]), "r"(b));
                } else if (MODE == 1) {   // IMAD.HI
                    asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(x[i]), "r"(b));
                } else if (MODE == 2) {   // IMAD.WIDE 64-bit accumulate, no carry
                    uint64_t acc = ((uint64_t)hi[i] << 32) | lo[i];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x[i]), "r"(b));
                    lo[i] = (uint32_t)acc; hi[i] = (uint32_t)(acc >> 32);
                }
            }
            if (MODE == 3) {              // one carry chain across the NCH pairs: lo/hi fused by ptxas?
                asm volatile(
                    "mad.lo.cc.u32 %0, %16, %24, %0;\n\t"
                    "madc.hi.cc.u32 %1, %16, %24, %1;\n\t"
                    "madc.lo.cc.u32 %2, %17, %24, %2;\n\t"
                    "madc.hi.cc.u32 %3, %17, %24, %3;\n\t"
                    "madc.lo.cc.u32 %4, %18, %24, %4;\n\t"
                    "madc.hi.cc.u32 %5, %18, %24, %5;\n\t"
                    "madc.lo.cc.u32 %6, %19, %24, %6;\n\t"
                    "madc.hi.cc.u32 %7, %19, %24, %7;\n\t"
                    "madc.lo.cc.u32 %8, %20, %24, %8;\n\t"
                    "madc.hi.cc.u32 %9, %20, %24, %9;\n\t"
                    "madc.lo.cc.u32 %10, %21, %24, %10;\n\t"
                    "madc.hi.cc.u32 %11, %21, %24, %11;\n\t"
                    "madc.lo.cc.u32 %12, %22, %24, %12;\n\t"
                    "madc.hi.cc.u32 %13, %22, %24, %13;\n\t"
                    "madc.lo.cc.u32 %14, %23, %24, %14;\n\t"
                    "madc.hi.u32 %15, %23, %24, %15;\n\t"
                    : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
                      "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
                    : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(b));
            }
            if (MODE == 4) {              // two independent carry chains of 4 pairs each
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
                    "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                    "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
                    "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                    "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
                    "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                    "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
                    "madc.hi.u32 %7, %11, %12, %7;\n\t"
                    : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3])
                    : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(b));
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
                    "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                    "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
                    "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                    "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
                    "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                    "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
                    "madc.hi.u32 %7, %11, %12, %7;\n\t"
                    : "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
                    : "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(b));
            }
            if (MODE == 5) {              // IMAD.WIDE (no carry) interleaved 1:1 with IADD3
#pragma unroll
                for (int i = 0; i < NCH; i++) {
                    uint64_t acc = ((uint64_t)hi[i] << 32) | lo[i];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x[i]), "r"(b));
                    lo[i] = (uint32_t)acc; hi[i] = (uint32_t)(acc >> 32);
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(lo[i]));
                }
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = a;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= lo[i] ^ hi[i] ^ x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char *name, int threads, int blocks_per_sm, double prods_per_inner)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * blocks_per_sm;
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 12345, cyc); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 12345, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long *h = new unsigned long long[blocks];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double prods = (double)ITERS * 4 * prods_per_inner * threads * (double)blocks;
    double per_sm_clk = (double)ITERS * 4 * prods_per_inner * threads * blocks_per_sm / avg;
    printf("%-34s thr=%4d bps=%d  %.3f ms  %8.2f Gprod/s  %6.2f prod/clk/SM (block cycles %.0f)  err=%s\n", name, threads, blocks_per_sm,
           ms, prods / ms * 1e-6, per_sm_clk, avg, cudaGetErrorString(cudaGetLastError()));
    delete[] h; cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int ths[3] = {256, 512, 1024};
    for (int t = 0; t < 3; t++) {
        int th = ths[t];
        run<0>("IMAD lo (mad.lo.u32)", th, 1, NCH);
        run<1>("IMAD.HI (mad.hi.u32)", th, 1, NCH);
        run<2>("IMAD.WIDE (mad.wide.u32)", th, 1, NCH);
        run<3>("IMAD.WIDE.X one 8-pair carry chain", th, 1, NCH);
        run<4>("IMAD.WIDE.X two 4-pair carry chains", th, 1, NCH);
        run<5>("IMAD.WIDE + IADD 1:1", th, 1, NCH);
    }
    run<2>("IMAD.WIDE (mad.wide.u32)", 128, 1, NCH);
    run<3>("IMAD.WIDE.X one 8-pair carry chain", 128, 1, NCH);
    run<2>("IMAD.WIDE (mad.wide.u32)", 512, 2, NCH);
    return 0;
}
