// special_fold.hpp -- reduction of a double-length product modulo a special-form base number
//     M = 2^k - c  (c >= 1: Mersenne and pseudo-Mersenne)      or      M = 2^k + 1,
// the shift-and-fold step that replaces the reference's vecmulmod52_mersenne tail
// (vecarith52.c:770-1025: "reduce by adding hi to lo", lo + hi*c, lo - hi).  Unlike the reference's
// lazily reduced words the result is the canonical residue in [0, M), which is what its operators
// return for all but a 2^-52 fraction of operands (tests/test_ref_fieldops.py).
//
// The same source compiles for the host (the carry flag becomes a variable) so that
// tests/test_special_fold_cpu.py can check every form and bit position limb for limb against Python.
// The word index W = k / 32 of bit k is a template parameter so that every array index is a
// compile-time constant (register arrays); s = k % 32 is a run-time shift count.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ECM_SF_HD __host__ __device__ __forceinline__
#else
#define ECM_SF_HD inline
#endif

namespace ecmb200 {

// ---- carry-chain primitives: PTX on the device, a flag variable on the host (same call sequence) -------------
#ifdef __CUDA_ARCH__
#define ECM_SF_ASM2(name, ins) ECM_SF_HD uint32_t name(uint32_t a, uint32_t b) { uint32_t d; asm volatile(ins " %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
#define ECM_SF_ASM3(name, ins) ECM_SF_HD uint32_t name(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile(ins " %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
ECM_SF_ASM2(sf_add_cc, "add.cc.u32") ECM_SF_ASM2(sf_addc_cc, "addc.cc.u32") ECM_SF_ASM2(sf_addc, "addc.u32")
ECM_SF_ASM2(sf_sub_cc, "sub.cc.u32") ECM_SF_ASM2(sf_subc_cc, "subc.cc.u32") ECM_SF_ASM2(sf_subc, "subc.u32")
ECM_SF_ASM3(sf_mad_lo_cc, "mad.lo.cc.u32") ECM_SF_ASM3(sf_madc_lo_cc, "madc.lo.cc.u32")
ECM_SF_ASM3(sf_mad_hi_cc, "mad.hi.cc.u32") ECM_SF_ASM3(sf_madc_hi_cc, "madc.hi.cc.u32") ECM_SF_ASM3(sf_madc_hi, "madc.hi.u32")
#else
static thread_local uint32_t sf_cf = 0;            // the CC.CF flag: carry of an add, borrow of a sub
inline uint32_t sf_set(uint64_t t) { sf_cf = (uint32_t)(t >> 32) & 1u; return (uint32_t)t; }
inline uint32_t sf_add_cc(uint32_t a, uint32_t b) { return sf_set((uint64_t)a + b); }
inline uint32_t sf_addc_cc(uint32_t a, uint32_t b) { return sf_set((uint64_t)a + b + sf_cf); }
inline uint32_t sf_addc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a + b + sf_cf); }
inline uint32_t sf_sub_cc(uint32_t a, uint32_t b) { return sf_set((uint64_t)a - b); }
inline uint32_t sf_subc_cc(uint32_t a, uint32_t b) { return sf_set((uint64_t)a - b - sf_cf); }
inline uint32_t sf_subc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a - b - sf_cf); }
inline uint32_t sf_mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return sf_set((uint64_t)(uint32_t)((uint64_t)a * b) + c); }
inline uint32_t sf_madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return sf_set((uint64_t)(uint32_t)((uint64_t)a * b) + c + sf_cf); }
inline uint32_t sf_mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return sf_set((((uint64_t)a * b) >> 32) + c); }
inline uint32_t sf_madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return sf_set((((uint64_t)a * b) >> 32) + c + sf_cf); }
inline uint32_t sf_madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)((((uint64_t)a * b) >> 32) + c + sf_cf); }
#endif

// limb of (x >> s) from two neighbouring limbs
ECM_SF_HD uint32_t sf_shr(uint32_t lo, uint32_t hi, uint32_t s)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);                         // one SHF; s = 0 returns lo
#else
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}

// kind > 0: M = 2^k - c; kind < 0: M = 2^k + 1.  T < M^2 (2*NL limbs), 32*NL > k >= 64, c < 2^31.
// About 6 single-issue instructions per limb (4 for 2^k+1), every one a link of a carry chain.
template <int NL, int W>
ECM_SF_HD void special_fold_w(uint32_t (&r)[NL], const uint32_t (&T)[2 * NL], uint32_t s, int kind, uint32_t c)
{
    static_assert(W >= 2 && W < NL, "bit k must lie inside the NL limbs");
    const uint32_t mask = s ? ((1u << s) - 1u) : 0u;          // bits of limb W that belong to the low part
    uint32_t hi[W + 2], x[W + 2];
#pragma unroll
    for (int j = 0; j < W + 2; j++) {                          // hi = T >> k
        const uint32_t a = (W + j < 2 * NL) ? T[W + j] : 0u;
        const uint32_t b = (W + j + 1 < 2 * NL) ? T[W + j + 1] : 0u;
        hi[j] = sf_shr(a, b, s);
    }
    const uint32_t lotop = T[W] & mask;                        // lo = T mod 2^k: limbs T[0..W-1], lotop
    if (kind > 0) {
        // x = lo + c*hi  (hi < 2^k; x < 2^k (1 + c): W+2 limbs)
        if (c == 1) {
            x[0] = sf_add_cc(T[0], hi[0]);
#pragma unroll
            for (int j = 1; j < W; j++) x[j] = sf_addc_cc(T[j], hi[j]);
            x[W] = sf_addc(lotop, hi[W]);
            x[W + 1] = 0;
        } else {
            x[0] = sf_mad_lo_cc(c, hi[0], T[0]);
#pragma unroll
            for (int j = 1; j < W; j++) x[j] = sf_madc_lo_cc(c, hi[j], T[j]);
            x[W] = sf_madc_lo_cc(c, hi[W], lotop);
            x[W + 1] = sf_addc(0, 0);
            x[1] = sf_mad_hi_cc(c, hi[0], x[1]);
#pragma unroll
            for (int j = 2; j <= W; j++) x[j] = sf_madc_hi_cc(c, hi[j - 1], x[j]);
            x[W + 1] = sf_madc_hi(c, hi[W], x[W + 1]);
        }
        // second fold: h2 = x >> k <= c;  x = (x mod 2^k) + c*h2 < 2^k + 2^63
        const uint32_t h2 = sf_shr(x[W], x[W + 1], s);
        x[W] &= mask;
        const uint64_t p = (uint64_t)c * h2;
        x[0] = sf_add_cc(x[0], (uint32_t)p);
        x[1] = sf_addc_cc(x[1], (uint32_t)(p >> 32));
#pragma unroll
        for (int j = 2; j < W; j++) x[j] = sf_addc_cc(x[j], 0);
        if (W > 1) x[W] = sf_addc(x[W], 0);
        // canonical: x >= M  <=>  y = x + c >= 2^k, and then x - M = y - 2^k
        uint32_t y[W + 1];
        y[0] = sf_add_cc(x[0], c);
#pragma unroll
        for (int j = 1; j < W; j++) y[j] = sf_addc_cc(x[j], 0);
        y[W] = sf_addc(x[W], 0);
        const bool ge = (s ? (y[W] >> s) : y[W]) != 0;
#pragma unroll
        for (int j = 0; j < W; j++) x[j] = ge ? y[j] : x[j];
        x[W] = ge ? (y[W] & mask) : x[W];
    } else {
        // T = lo + 2^k mid + 2^2k top with 2^k = -1:  x = lo - mid + top; negative -> add M = 2^k + 1
        const uint32_t top = sf_shr(hi[W], hi[W + 1], s);      // 0 or 1 (1 only for T = 2^2k, where lo = mid = 0)
        x[0] = sf_sub_cc(T[0], hi[0]);
#pragma unroll
        for (int j = 1; j < W; j++) x[j] = sf_subc_cc(T[j], hi[j]);
        x[W] = sf_subc_cc(lotop, hi[W] & mask);
        const uint32_t neg = sf_subc(0, 0);                    // all-ones when lo < mid
        x[0] = sf_add_cc(x[0], top | (neg & 1u));
#pragma unroll
        for (int j = 1; j < W; j++) x[j] = sf_addc_cc(x[j], 0);
        x[W] = sf_addc(x[W], neg & (1u << s));
    }
#pragma unroll
    for (int j = 0; j < NL; j++) r[j] = (j <= W) ? x[j] : 0u;
}

// run-time W -> compile-time W.  LOW = lowest supported word index for this NL: bases shorter than that are
// served by the next smaller kernel set, so the build passes the previous compiled limb count as ECM_SP_LOW
// (a base of more than 32*prev bits has k/32 >= prev); without it a window of 9 positions is compiled.
#ifndef ECM_SP_LOW
#define ECM_SP_LOW 0
#endif
template <int NL> struct SpecialRange {
    static constexpr int LOW = (ECM_SP_LOW >= 2 && ECM_SP_LOW < NL) ? ECM_SP_LOW : ((NL > 10) ? NL - 9 : 2);
};

template <int NL, int W>
ECM_SF_HD void special_fold_dispatch(uint32_t (&r)[NL], const uint32_t (&T)[2 * NL], uint32_t w, uint32_t s, int kind, uint32_t c)
{
    if constexpr (W < SpecialRange<NL>::LOW) {
#pragma unroll
        for (int j = 0; j < NL; j++) r[j] = 0;               // unreachable: the host never selects such a base
    } else {
        if (w == (uint32_t)W) special_fold_w<NL, W>(r, T, s, kind, c);
        else special_fold_dispatch<NL, W - 1>(r, T, w, s, kind, c);
    }
}

template <int NL>
ECM_SF_HD void special_fold(uint32_t (&r)[NL], const uint32_t (&T)[2 * NL], uint32_t kbits, int kind, uint32_t c)
{
    special_fold_dispatch<NL, NL - 1>(r, T, kbits >> 5, kbits & 31u, kind, c);
}

}  // namespace ecmb200
