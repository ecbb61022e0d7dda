// mp.cuh -- multi-precision field arithmetic mod N for sm_100a, 32-bit limbs, one value
// per thread held in registers.  Replaces the reference's vector field ops
//   vecmulmod52 / vecsqrmod52   (vecarith52.c:2438-3074, 3317-4548)
//   vecaddmod52 / vecsubmod52 / vec_simul_addsub52 (vecarith52.c:4550-4611, 4684-4723, 4877-4968)
// Every routine returns the canonical representative in [0,N) exactly like the reference
// (vecarith52.c:3048-3070), so residues are identical whatever R is (R = 2^(32*NL) here).
//
// Multiply: word-serial Montgomery (CIOS) where each 32x32->64 product is ONE
// IMAD.WIDE.U32[.X]: PTX "mad.lo.cc / madc.hi.cc" pairs on an even-aligned register pair are
// fused by ptxas.  Because IMAD.WIDE accumulates into an aligned 64-bit pair, the running sum
// T is kept as two interleaved accumulators, T = E + O*2^32: products a[j]*b_i with j even
// go to E (pair j,j+1), with j odd go to O (pair j-1,j).  After the reduction row the
// division by 2^32 swaps the roles: T/2^32 = O + E[1] + (E>>64)*2^32.
#pragma once
#include <stdint.h>

namespace ecmb200 {

// ---- carry-chain primitives (CC flag lives between consecutive asm volatile statements) ----
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
__device__ __forceinline__ void add_cc(uint32_t &d, uint32_t a) { asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(d) : "r"(a)); }
__device__ __forceinline__ void addc_cc(uint32_t &d, uint32_t a) { asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(d) : "r"(a)); }
__device__ __forceinline__ void addc(uint32_t &d, uint32_t a) { asm volatile("addc.u32 %0, %0, %1;" : "+r"(d) : "r"(a)); }
__device__ __forceinline__ uint32_t add3_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t addc3_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t addc3(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t sub3_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t subc3_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t subc3(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ void mad_lo_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b)); }
__device__ __forceinline__ void madc_lo_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b)); }
__device__ __forceinline__ void madc_hi_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b)); }

// Modulus and Montgomery constants, passed by value as a kernel parameter so that the limbs
// are constant-bank operands of the IMADs (no registers spent on N).
template <int NL>
struct ModParams {
    uint32_t n[NL];      // modulus N (odd)
    uint32_t one[NL];    // R mod N          (monty.one, main.c:633-634)
    uint32_t r2[NL];     // R^2 mod N        (to enter Montgomery form)
    uint32_t r3[NL];     // R^3 mod N        (fix-up after a raw modular inverse)
    uint32_t rrefinv[NL];// (2^-MAXBITS_ref) * R mod N : Montgomery form of the reference's R^-1, used
                         // only on the inversion-failure path (see vm.cuh op INV)
    uint32_t m0inv;      // -N^-1 mod 2^32   (monty.vrho, main.c:637-640)
};

template <int NL> struct MontW { static constexpr int W = (NL % 2 == 0) ? NL + 2 : NL + 3; };

// acc += x[j]*y over the even (PAR=0) or odd (PAR=1) limbs j of x; pairs land on acc[j-PAR], acc[j-PAR+1].
// FIRST_CARRY: the first instruction consumes the CC flag left by the caller.
template <int NL, int W, int PAR, bool FIRST_CARRY, typename XT>
__device__ __forceinline__ void mad_row(uint32_t (&acc)[W], const XT &x, uint32_t y)
{
    constexpr int TOP = (PAR == 0) ? NL + 1 : NL;       // highest word this accumulator can reach
    int k = 0;
#pragma unroll
    for (int j = PAR; j < NL; j += 2) {
        k = j - PAR;
        if (j == PAR && !FIRST_CARRY) mad_lo_cc(acc[k], x[j], y);
        else madc_lo_cc(acc[k], x[j], y);
        madc_hi_cc(acc[k + 1], x[j], y);
    }
    constexpr int NPAIR = (NL - PAR + 1) / 2;
    constexpr int T0 = 2 * NPAIR;                        // first word after the product pairs
#pragma unroll
    for (int t = T0; t <= TOP; t++) {
        if (t == TOP) addc(acc[t], 0); else addc_cc(acc[t], 0);
    }
}

// r = a*b*R^-1 mod N, canonical.  a,b canonical (< N).
template <int NL>
__device__ __forceinline__ void mont_mul(uint32_t (&r)[NL], const uint32_t (&a)[NL], const uint32_t (&b)[NL],
                                         const ModParams<NL> &P)
{
    constexpr int W = MontW<NL>::W;
    uint32_t X[W], Y[W];
#pragma unroll
    for (int k = 0; k < W; k++) { X[k] = 0; Y[k] = 0; }

    // one row: (Eo,Oo) are the current E/O roles; afterwards the roles are swapped.
    auto row = [&](uint32_t (&Eo)[W], uint32_t (&Oo)[W], uint32_t bi) {
        uint32_t e1 = Eo[1];
#pragma unroll
        for (int k = 0; k < W - 2; k++) Eo[k] = Eo[k + 2];     // O' = E >> 64 (register renaming)
        Eo[W - 2] = 0; Eo[W - 1] = 0;
        // E' = O + e1 ; the carry of that add enters the O' chain at word 0
        add_cc(Oo[0], e1);
        if (NL > 1) mad_row<NL, W, 1, true>(Eo, a, bi);
        else { addc_cc(Eo[0], 0); addc(Eo[1], 0); }
        mad_row<NL, W, 0, false>(Oo, a, bi);
        uint32_t m = mul_lo(Oo[0], P.m0inv);
        if (NL > 1) mad_row<NL, W, 1, false>(Eo, P.n, m);
        mad_row<NL, W, 0, false>(Oo, P.n, m);
    };
#pragma unroll
    for (int i = 0; i < NL; i++) {
        if ((i & 1) == 0) row(X, Y, b[i]); else row(Y, X, b[i]);
    }
    // after NL rows: E role is X if NL even else Y.  T' = (E>>32) + O, NL+1 words.
    uint32_t t[NL + 1];
    {
        uint32_t (&E)[W] = (NL % 2 == 0) ? X : Y;
        uint32_t (&O)[W] = (NL % 2 == 0) ? Y : X;
        t[0] = add3_cc(E[1], O[0]);
#pragma unroll
        for (int k = 1; k < NL; k++) t[k] = addc3_cc(E[k + 1], O[k]);
        t[NL] = addc3(E[NL + 1], O[NL]);
    }
    // conditional subtraction: T' < 2N
    uint32_t d[NL];
    d[0] = sub3_cc(t[0], P.n[0]);
#pragma unroll
    for (int k = 1; k < NL; k++) d[k] = subc3_cc(t[k], P.n[k]);
    uint32_t nb = subc3(t[NL], 0);           // 0 or 1 if T' >= N, 0xffffffff if T' < N
    bool take = (nb != 0xffffffffu);
#pragma unroll
    for (int k = 0; k < NL; k++) r[k] = take ? d[k] : t[k];
}

// Two independent products at once, rows interleaved: (r0,r1) = (a0*b0, a1*b1) * R^-1 mod N.
// Same arithmetic as mont_mul; writing the rows alternately hands ptxas twice as many independent
// IMAD.WIDE carry chains (and two independent reduction tails), which is what hides the
// fixed-latency stalls when only ~3 warps share a sub-partition.
template <int NL>
__device__ __forceinline__ void mont_mul2(uint32_t (&r0)[NL], const uint32_t (&a0)[NL], const uint32_t (&b0)[NL],
                                          uint32_t (&r1)[NL], const uint32_t (&a1)[NL], const uint32_t (&b1)[NL],
                                          const ModParams<NL> &P)
{
    constexpr int W = MontW<NL>::W;
    uint32_t X0[W], Y0[W], X1[W], Y1[W];
#pragma unroll
    for (int k = 0; k < W; k++) { X0[k] = 0; Y0[k] = 0; X1[k] = 0; Y1[k] = 0; }
    auto row = [&](uint32_t (&Eo)[W], uint32_t (&Oo)[W], const uint32_t (&a)[NL], uint32_t bi) {
        uint32_t e1 = Eo[1];
#pragma unroll
        for (int k = 0; k < W - 2; k++) Eo[k] = Eo[k + 2];
        Eo[W - 2] = 0; Eo[W - 1] = 0;
        add_cc(Oo[0], e1);
        if (NL > 1) mad_row<NL, W, 1, true>(Eo, a, bi);
        else { addc_cc(Eo[0], 0); addc(Eo[1], 0); }
        mad_row<NL, W, 0, false>(Oo, a, bi);
        uint32_t m = mul_lo(Oo[0], P.m0inv);
        if (NL > 1) mad_row<NL, W, 1, false>(Eo, P.n, m);
        mad_row<NL, W, 0, false>(Oo, P.n, m);
    };
#pragma unroll
    for (int i = 0; i < NL; i++) {
        if ((i & 1) == 0) { row(X0, Y0, a0, b0[i]); row(X1, Y1, a1, b1[i]); }
        else { row(Y0, X0, a0, b0[i]); row(Y1, X1, a1, b1[i]); }
    }
    auto finish = [&](uint32_t (&r)[NL], uint32_t (&X)[W], uint32_t (&Y)[W]) {
        uint32_t t[NL + 1];
        uint32_t (&E)[W] = (NL % 2 == 0) ? X : Y;
        uint32_t (&O)[W] = (NL % 2 == 0) ? Y : X;
        t[0] = add3_cc(E[1], O[0]);
#pragma unroll
        for (int k = 1; k < NL; k++) t[k] = addc3_cc(E[k + 1], O[k]);
        t[NL] = addc3(E[NL + 1], O[NL]);
        uint32_t d[NL];
        d[0] = sub3_cc(t[0], P.n[0]);
#pragma unroll
        for (int k = 1; k < NL; k++) d[k] = subc3_cc(t[k], P.n[k]);
        uint32_t nb = subc3(t[NL], 0);
        bool take = (nb != 0xffffffffu);
#pragma unroll
        for (int k = 0; k < NL; k++) r[k] = take ? d[k] : t[k];
    };
    finish(r0, X0, Y0);
    finish(r1, X1, Y1);
}

template <int NL>
__device__ __forceinline__ void mont_sqr(uint32_t (&r)[NL], const uint32_t (&a)[NL], const ModParams<NL> &P)
{
    mont_mul<NL>(r, a, a, P);
}

// r = (a+b) mod N, canonical   (vecaddmod52, vecarith52.c:4550-4611)
template <int NL>
__device__ __forceinline__ void mod_add(uint32_t (&r)[NL], const uint32_t (&a)[NL], const uint32_t (&b)[NL],
                                        const ModParams<NL> &P)
{
    uint32_t t[NL], d[NL];
    t[0] = add3_cc(a[0], b[0]);
#pragma unroll
    for (int k = 1; k < NL; k++) t[k] = addc3_cc(a[k], b[k]);
    uint32_t c = addc3(0, 0);
    d[0] = sub3_cc(t[0], P.n[0]);
#pragma unroll
    for (int k = 1; k < NL; k++) d[k] = subc3_cc(t[k], P.n[k]);
    uint32_t nb = subc3(c, 0);
    bool take = (nb != 0xffffffffu);
#pragma unroll
    for (int k = 0; k < NL; k++) r[k] = take ? d[k] : t[k];
}

// r = (a-b) mod N, canonical   (vecsubmod52, vecarith52.c:4684-4723)
template <int NL>
__device__ __forceinline__ void mod_sub(uint32_t (&r)[NL], const uint32_t (&a)[NL], const uint32_t (&b)[NL],
                                        const ModParams<NL> &P)
{
    uint32_t t[NL];
    t[0] = sub3_cc(a[0], b[0]);
#pragma unroll
    for (int k = 1; k < NL; k++) t[k] = subc3_cc(a[k], b[k]);
    uint32_t bo = subc3(0, 0);               // 0xffffffff if a < b
    r[0] = add3_cc(t[0], P.n[0] & bo);
#pragma unroll
    for (int k = 1; k < NL; k++) r[k] = (k == NL - 1) ? addc3(t[k], P.n[k] & bo) : addc3_cc(t[k], P.n[k] & bo);
    if (NL == 1) { /* add3_cc already final */ }
}

// 1 if a == 0
template <int NL>
__device__ __forceinline__ bool is_zero(const uint32_t (&a)[NL])
{
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < NL; k++) o |= a[k];
    return o == 0;
}

}  // namespace ecmb200
