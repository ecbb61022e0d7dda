// geom.hpp -- how curve c maps into a [group][slot][limb][stride] array (shared by host and device).
#pragma once
#include <cstdint>
#include <cstddef>
#ifdef __CUDACC__
#define ECM_HD __host__ __device__
#else
#define ECM_HD
#endif
namespace ecmb200 {
struct Geom {
    uint32_t T, stride, nslot;     // curves per group, lane stride, slots per group
    uint32_t L = 1;                // lanes per curve: a value's NL limbs are striped over L adjacent lanes, NL/L limbs each
                                   // (warp-cooperative kernels, coop.cuh); 1 = one thread per curve
    ECM_HD size_t idx(uint32_t c, uint32_t slot, uint32_t k, uint32_t NL) const
    {
        const uint32_t M = NL / L;
        return (((size_t)(c / T) * nslot + slot) * M + k % M) * stride + (c % T) * L + k / M;
    }
};
}  // namespace ecmb200
