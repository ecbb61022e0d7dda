// imad_peak.cu -- live integer-multiply roofline of the GPU this process runs on.
// Retires dependency-free chains of IMAD.WIDE.U32.X (mad.lo.cc/madc.hi.cc pairs, the exact
// instruction form mont_mul is built from; per-thread multiplicands in vector registers) and
// reports 32x32->64 products per second, plus the SM clock derived from clock64().
#include "../../include/ecm_b200.h"
#include <cuda_runtime.h>
#include <cstdint>
#include "mp.cuh"

namespace {

constexpr int kIters = 2048;

// UNIFORM_B: the multiplier is the same for the whole warp (like the limbs of N, which ptxas keeps in
// uniform registers); otherwise it is a per-thread value (like b[i] and the quotient digit m).
template <bool UNIFORM_B>
__global__ void __launch_bounds__(1024) k_imad_peak(uint32_t *out, uint32_t seed, unsigned long long *cyc)
{
    uint32_t lo[8], hi[8], x[8];
    uint32_t b = seed * 3u + (UNIFORM_B ? blockIdx.x : threadIdx.x) * 0x9e3779b9u;
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

// Third probe: the densest real IMAD.WIDE stream we know -- a register-resident 32-limb Montgomery
// multiply chained on itself (2*32^2+32 products per call, no memory traffic).  It sustains more than
// the synthetic chains above because ptxas interleaves ~6 carry chains; whichever probe is fastest
// defines the roof.
constexpr int kMontIters = 600;
__global__ void __launch_bounds__(384) k_mont_peak(const ecmb200::ModParams<32> P, uint32_t *out, uint32_t seed)
{
    uint32_t a[32], b[32];
#pragma unroll
    for (int k = 0; k < 32; k++) { a[k] = (seed + threadIdx.x) * (2 * k + 1); b[k] = (seed ^ blockIdx.x) * (2 * k + 3) + threadIdx.x; }
    a[31] &= 0x7fffffffu; b[31] &= 0x7fffffffu;            // operands below N (N has its top bit set)
#pragma unroll 1
    for (int it = 0; it < kMontIters; it++) ecmb200::mont_mul<32>(a, a, b, P);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 32; k++) s ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    for (int rep = 0; rep < 10; rep++) {         // reps 0,1 warm up; both multiplier forms, best wins
        cudaEventRecord(e0);
        if (rep & 1) k_imad_peak<true><<<blocks, threads>>>(out, 12345u + rep, cyc);
        else k_imad_peak<false><<<blocks, threads>>>(out, 12345u + rep, cyc);
        ecmb200::count_launch();
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); cudaFree(cyc); return ECM_B200_ECUDA; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double prods = (double)kIters * 4 * 8 * threads * blocks;
        const double rate = prods / (ms * 1e-3);
        if (rep > 1 && rate > best) {
            best = rate;
            unsigned long long c0 = 0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
            clk = (double)c0 / (ms * 1e-3) / 1e6;
        }
    }
    {
        ecmb200::ModParams<32> P;
        for (int k = 0; k < 32; k++) { P.n[k] = 0xffffffffu - 2u * (k == 0 ? 9u : (uint32_t)k); P.one[k] = P.r2[k] = P.r3[k] = P.rrefinv[k] = k + 1; }
        P.n[0] |= 1u; P.n[31] |= 0x80000000u;
        uint32_t inv = 1; for (int i = 0; i < 5; i++) inv *= 2 - P.n[0] * inv;
        P.m0inv = 0u - inv;
        const int mthreads = 384;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            k_mont_peak<<<blocks, mthreads>>>(P, out, 777u + rep);
            ecmb200::count_launch();
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); cudaFree(cyc); return ECM_B200_ECUDA; }
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            const double rate = (double)kMontIters * (2.0 * 32 * 32 + 32) * mthreads * blocks / (ms * 1e-3);
            if (rep > 0 && rate > best) best = rate;
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out); cudaFree(cyc);
    if (products_per_sec) *products_per_sec = best;
    if (sm_clock_mhz) *sm_clock_mhz = clk;
    return ECM_B200_OK;
}
