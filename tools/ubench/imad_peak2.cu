// variants of dependency-free IMAD.WIDE.U32.X chains: find the best achievable issue rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int NP, int NCHAIN, bool UNI>   // NP pairs per chain, NCHAIN chains per thread
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t seed, unsigned long long *cyc)
{
    uint32_t lo[NP * NCHAIN], hi[NP * NCHAIN], x[NP * NCHAIN];
    uint32_t b = seed * 3u + (UNI ? blockIdx.x : threadIdx.x) * 0x9e3779b9u;
#pragma unroll
    for (int i = 0; i < NP * NCHAIN; i++) { lo[i] = seed + i; hi[i] = b ^ i; x[i] = (seed + threadIdx.x) * (2 * i + 3) + 1; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            b += 0x9e3779b9u;
#pragma unroll
            for (int c = 0; c < NCHAIN; c++) {
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    const int i = c * NP + j;
                    if (j == 0) asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(x[i]), "r"(b));
                    else asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(x[i]), "r"(b));
                    if (j == NP - 1) asm volatile("madc.hi.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(x[i]), "r"(b));
                    else asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(x[i]), "r"(b));
                }
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NP * NCHAIN; i++) s ^= lo[i] ^ hi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}
template <int NP, int NCHAIN, bool UNI> void run(int threads)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, (size_t)sms * threads * 4); cudaMalloc(&cyc, sms * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<NP, NCHAIN, UNI><<<sms, threads>>>(out, 123 + rep, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    unsigned long long c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
    double prods = (double)ITERS * 2 * NP * NCHAIN * threads * sms;
    printf("NP=%2d NCHAIN=%d uni=%d thr=%4d: %.3f ms %8.1f Gprod/s  %.2f prod/clk/SM\n", NP, NCHAIN, (int)UNI, threads, best, prods / best * 1e-6,
           (double)ITERS * 2 * NP * NCHAIN * threads / (double)c0);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int th : {512, 1024}) {
        run<8, 1, true>(th); run<8, 1, false>(th); run<16, 1, true>(th); run<16, 1, false>(th);
        run<4, 4, true>(th); run<4, 4, false>(th); run<8, 2, true>(th); run<8, 2, false>(th); run<13, 2, true>(th); run<7,4,false>(th);
    }
    return 0;
}
