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
    ECM_HD size_t idx(uint32_t c, uint32_t slot, uint32_t k, uint32_t NL) const
    {
        return (((size_t)(c / T) * nslot + slot) * NL + k) * stride + (c % T);
    }
};
}  // namespace ecmb200
