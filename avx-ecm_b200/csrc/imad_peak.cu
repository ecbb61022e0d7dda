// imad_peak.cu -- live integer-multiply roofline of the GPU this process runs on.
// Retires dependency-free chains of IMAD.WIDE.U32.X (mad.lo.cc/madc.hi.cc pairs, the exact
// instruction form mont_mul is built from; per-thread multiplicands in vector registers) and
// reports 32x32->64 products per second, plus the SM clock derived from clock64().
#include "../../include/ecm_b200.h"
#include <cuda_runtime.h>
#include <cstdint>

namespace {

constexpr int kIters = 2048;

__global__ void __launch_bounds__(1024) k_imad_peak(uint32_t *out, uint32_t seed, unsigned long long *cyc)
{
    uint32_t lo[8], hi[8], x[8];
    uint32_t b = seed * 3u + threadIdx.x * 0x9e3779b9u;
#pragma unroll
    for (int i = 0; i < 8; i++) { lo[i] = seed + i; hi[i] = b ^ i; x[i] = (seed + threadIdx.x) * (2 * i + 3) + 1; }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            b += 0x9e3779b9u;
            // two independent 4-product carry chains
            asm volatile(
                "mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;\n\t"
                : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3])
                : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(b));
            asm volatile(
                "mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;\n\t"
                : "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
                : "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(b));
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= lo[i] ^ hi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

}  // namespace

namespace ecmb200 { void count_launch(); }

extern "C" int ecm_b200_measure_imad_peak(int device, double *products_per_sec, double *sm_clock_mhz)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ECM_B200_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return ECM_B200_ECUDA;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int threads = 1024, blocks = sms;
    uint32_t *out = nullptr; unsigned long long *cyc = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * 4) != cudaSuccess) return ECM_B200_ECUDA;
    if (cudaMalloc(&cyc, (size_t)blocks * 8) != cudaSuccess) { cudaFree(out); return ECM_B200_ECUDA; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0, clk = 0;
    for (int rep = 0; rep < 6; rep++) {          // rep 0 warms up
        cudaEventRecord(e0);
        k_imad_peak<<<blocks, threads>>>(out, 12345u + rep, cyc);
        ecmb200::count_launch();
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); cudaFree(cyc); return ECM_B200_ECUDA; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double prods = (double)kIters * 4 * 8 * threads * blocks;
        const double rate = prods / (ms * 1e-3);
        if (rep > 0 && rate > best) {
            best = rate;
            unsigned long long c0 = 0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
            clk = (double)c0 / (ms * 1e-3) / 1e6;
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out); cudaFree(cyc);
    if (products_per_sec) *products_per_sec = best;
    if (sm_clock_mhz) *sm_clock_mhz = clk;
    return ECM_B200_OK;
}
