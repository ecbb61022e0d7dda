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
#include "special_fold.hpp"

// Kernel variant of this translation unit.  ECM_SPECIAL = 1: every modular product is a plain double-length
// product followed by the shift-and-fold reduction of special_fold.hpp (bases 2^k-c, 2^k+1; residues are
// plain, i.e. "Montgomery form with R = 1": one = r2 = r3 = 1).  The variants live in different inline
// namespaces so that their kernels and engine classes are distinct symbols in libecm_b200.so.
#ifndef ECM_SPECIAL
#define ECM_SPECIAL 0
#endif
#if ECM_SPECIAL
#define ECM_VNS sp
#else
#define ECM_VNS gen
#endif

namespace ecmb200 {
inline namespace ECM_VNS {

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
    uint32_t rref[NL];   // the reference's R = 2^MAXBITS_ref mod N as a plain integer (1 for special-form inputs): only for
                         // the stale-high-words emulation of the stage-2 inversion (kernels.cuh)
    uint32_t m0inv;      // -N^-1 mod 2^32   (monty.vrho, main.c:637-640)
    // special-form base (only read by the ECM_SPECIAL kernels): N = 2^kbits - cval (kind > 0) or 2^kbits + 1 (kind < 0)
    int32_t kind;
    uint32_t kbits, cval;
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

// T = a*b, all 2*NL limbs: the rows of mont_mul without the reduction half.  The low word of each row is
// final (nothing is added below it afterwards) and is retired to T[i]; the window that is left after NL
// rows is the high half.
template <int NL>
__device__ __forceinline__ void full_mul(uint32_t (&T)[2 * NL], const uint32_t (&a)[NL], const uint32_t (&b)[NL])
{
    constexpr int W = MontW<NL>::W;
    uint32_t X[W], Y[W];
#pragma unroll
    for (int k = 0; k < W; k++) { X[k] = 0; Y[k] = 0; }
    auto row = [&](uint32_t (&Eo)[W], uint32_t (&Oo)[W], uint32_t bi, uint32_t &low) {
        uint32_t e1 = Eo[1];
#pragma unroll
        for (int k = 0; k < W - 2; k++) Eo[k] = Eo[k + 2];
        Eo[W - 2] = 0; Eo[W - 1] = 0;
        add_cc(Oo[0], e1);
        if (NL > 1) mad_row<NL, W, 1, true>(Eo, a, bi);
        else { addc_cc(Eo[0], 0); addc(Eo[1], 0); }
        mad_row<NL, W, 0, false>(Oo, a, bi);
        low = Oo[0];
    };
#pragma unroll
    for (int i = 0; i < NL; i++) {
        if ((i & 1) == 0) row(X, Y, b[i], T[i]); else row(Y, X, b[i], T[i]);
    }
    uint32_t (&E)[W] = (NL % 2 == 0) ? X : Y;
    uint32_t (&O)[W] = (NL % 2 == 0) ? Y : X;
    T[NL] = add3_cc(E[1], O[0]);
#pragma unroll
    for (int k = 1; k < NL - 1; k++) T[NL + k] = addc3_cc(E[k + 1], O[k]);
    T[2 * NL - 1] = addc3(E[NL], O[NL - 1]);
}

// r = a*b mod N for a special-form N (P.kind/kbits/cval), canonical
template <int NL>
__device__ __forceinline__ void special_mul(uint32_t (&r)[NL], const uint32_t (&a)[NL], const uint32_t (&b)[NL],
                                            const ModParams<NL> &P)
{
    uint32_t T[2 * NL];
    full_mul<NL>(T, a, b);
    special_fold<NL>(r, T, P.kbits, P.kind, P.cval);
}

// r = a*b*R^-1 mod N, canonical.  a,b canonical (< N).
template <int NL>
__device__ __forceinline__ void mont_mul(uint32_t (&r)[NL], const uint32_t (&a)[NL], const uint32_t (&b)[NL],
                                         const ModParams<NL> &P)
{
#if ECM_SPECIAL
    special_mul<NL>(r, a, b, P);
#else
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
#endif
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
#if ECM_SPECIAL
    {
        // (interleaving the rows of the two double-length products was tried and changed nothing: 88.5 k curves/s at
        // 2^415-1 either way -- the fold kernels are bound by the dependent chains of the fold itself)
        uint32_t t0[NL];                                   // r0 may alias a1/b1
        special_mul<NL>(t0, a0, b0, P);
        special_mul<NL>(r1, a1, b1, P);
#pragma unroll
        for (int k = 0; k < NL; k++) r0[k] = t0[k];
        return;
    }
#endif
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

// ---- dedicated squaring -------------------------------------------------------------------------
// r_k = a_k^2 * R^-1 mod N for K independent operands (K = 1 or 2, rows interleaved like mont_mul2).
// Replaces vecsqrmod52 (vecarith52.c:3317-4548), which likewise doubles the symmetric terms.
//   1. off-diagonal products a_i*a_j (i<j) once: NL(NL-1)/2 IMAD.WIDE, even positions i+j into SE,
//      odd positions into SO (same aligned-pair trick as the multiply);
//   2. S = 2*(SE + SO<<32) + sum a_i^2 * 2^(64 i)  (NL more products);
//   3. Montgomery reduction of the 2NL-limb square through an (NL+2)-limb window kept in the same
//      E/O form as the multiply: each row adds m*N (NL products), the division by 2^32 swaps the
//      roles, and one more limb of S enters at the top (its carry-out waits in `pend`).
// 1.5 NL^2 + 2.5 NL products instead of 2 NL^2 + NL.  The word-level algorithm was checked first in
// tools/models/mont_sqr_model.py (every carry asserted) and transcribed from there.
template <int NL, int W, int PAR, bool FIRST_CARRY, typename XT>
__device__ __forceinline__ uint32_t mad_row_capture(uint32_t (&acc)[W], const XT &x, uint32_t y)
{   // mad_row whose carry out of the top word is returned instead of being impossible
    static_assert(PAR == 0, "only the E role can overflow its top word");
    constexpr int TOP = NL + 1;
#pragma unroll
    for (int j = 0; j < NL; j += 2) {
        if (j == 0 && !FIRST_CARRY) mad_lo_cc(acc[j], x[j], y);
        else madc_lo_cc(acc[j], x[j], y);
        madc_hi_cc(acc[j + 1], x[j], y);
    }
    constexpr int T0 = 2 * ((NL + 1) / 2);
#pragma unroll
    for (int t = T0; t <= TOP; t++) addc_cc(acc[t], 0);
    return addc3(0, 0);
}

template <int NL, int K>
__device__ __forceinline__ void mont_sqr_k(uint32_t (&r)[K][NL], const uint32_t (&a)[K][NL], const ModParams<NL> &P)
{
    constexpr int W2 = 2 * NL + 2;
    constexpr int W = MontW<NL>::W;
    uint32_t SE[K][W2], SO[K][W2];
#pragma unroll
    for (int q = 0; q < K; q++)
#pragma unroll
        for (int k = 0; k < W2; k++) { SE[q][k] = 0; SO[q][k] = 0; }
    // 1. off-diagonal products
#pragma unroll
    for (int i = 0; i < NL - 1; i++) {
#pragma unroll
        for (int q = 0; q < K; q++) {
            if (i + 2 < NL) {                               // even positions: j = i+2, i+4, ...
#pragma unroll
                for (int j = i + 2; j < NL; j += 2) {
                    if (j == i + 2) mad_lo_cc(SE[q][i + j], a[q][j], a[q][i]); else madc_lo_cc(SE[q][i + j], a[q][j], a[q][i]);
                    madc_hi_cc(SE[q][i + j + 1], a[q][j], a[q][i]);
                }
                addc(SE[q][i + (i + 2 + 2 * ((NL - 1 - (i + 2)) / 2)) + 2], 0);
            }
            {                                               // odd positions: j = i+1, i+3, ...
#pragma unroll
                for (int j = i + 1; j < NL; j += 2) {
                    if (j == i + 1) mad_lo_cc(SO[q][i + j - 1], a[q][j], a[q][i]); else madc_lo_cc(SO[q][i + j - 1], a[q][j], a[q][i]);
                    madc_hi_cc(SO[q][i + j], a[q][j], a[q][i]);
                }
                addc(SO[q][i + (i + 1 + 2 * ((NL - 1 - (i + 1)) / 2)) - 1 + 2], 0);
            }
        }
    }
    // 2. S = 2*(SE + SO<<32) + diagonal, in place in SE[0..2NL)
#pragma unroll
    for (int q = 0; q < K; q++) {
        add_cc(SE[q][1], SO[q][0]);
#pragma unroll
        for (int k = 2; k < 2 * NL; k++) { if (k < 2 * NL - 1) addc_cc(SE[q][k], SO[q][k - 1]); else addc(SE[q][k], SO[q][k - 1]); }
#pragma unroll
        for (int k = 2 * NL - 1; k >= 1; k--) SE[q][k] = __funnelshift_l(SE[q][k - 1], SE[q][k], 1);
        SE[q][0] <<= 1;
#pragma unroll
        for (int i = 0; i < NL; i++) {
            if (i == 0) mad_lo_cc(SE[q][0], a[q][0], a[q][0]); else madc_lo_cc(SE[q][2 * i], a[q][i], a[q][i]);
            madc_hi_cc(SE[q][2 * i + 1], a[q][i], a[q][i]);      // the last carry-out is zero: the square fits 2NL limbs
        }
    }
    // 3. windowed Montgomery reduction
    uint32_t X[K][W], Y[K][W], pend[K];
#pragma unroll
    for (int q = 0; q < K; q++) {
        pend[q] = 0;
#pragma unroll
        for (int k = 0; k < W; k++) { X[q][k] = (k < NL + 2) ? SE[q][k] : 0; Y[q][k] = 0; }
    }
    auto row = [&](uint32_t (&Eo)[W], uint32_t (&Oo)[W], uint32_t &pd, uint32_t tin, bool shift) {
        // on entry (Eo,Oo) are the roles of the previous row; with shift they swap
        if (shift) {
            uint32_t e1 = Eo[1];
#pragma unroll
            for (int k = 0; k < W - 2; k++) Eo[k] = Eo[k + 2];
            Eo[W - 2] = 0; Eo[W - 1] = 0;
            Oo[NL + 1] = add3_cc(tin, pd); pd = addc3(0, 0);      // next limb of the square enters the new E
            add_cc(Oo[0], e1);
            uint32_t m = mul_lo(Oo[0], P.m0inv);
            if (NL > 1) mad_row<NL, W, 1, true>(Eo, P.n, m); else { addc_cc(Eo[0], 0); addc(Eo[1], 0); }
            pd += mad_row_capture<NL, W, 0, false>(Oo, P.n, m);
        } else {
            uint32_t m = mul_lo(Eo[0], P.m0inv);
            if (NL > 1) mad_row<NL, W, 1, false>(Oo, P.n, m);
            pd += mad_row_capture<NL, W, 0, false>(Eo, P.n, m);
        }
    };
#pragma unroll
    for (int i = 0; i < NL; i++) {
#pragma unroll
        for (int q = 0; q < K; q++) {
            const uint32_t tin = (i >= 1 && NL + 1 + i < 2 * NL) ? SE[q][NL + 1 + i] : 0;
            if (i == 0) row(X[q], Y[q], pend[q], 0, false);
            else if ((i & 1) == 1) row(X[q], Y[q], pend[q], tin, true);      // roles before: E=X -> after: E=Y
            else row(Y[q], X[q], pend[q], tin, true);
        }
    }
    // after NL rows the E role is X when NL is odd (row 0 does not swap), Y when NL is even
#pragma unroll
    for (int q = 0; q < K; q++) {
        uint32_t (&E)[W] = (NL % 2 == 1) ? X[q] : Y[q];
        uint32_t (&O)[W] = (NL % 2 == 1) ? Y[q] : X[q];
        uint32_t t[NL + 1], d[NL];
        t[0] = add3_cc(E[1], O[0]);
#pragma unroll
        for (int k = 1; k < NL; k++) t[k] = addc3_cc(E[k + 1], O[k]);
        t[NL] = addc3(E[NL + 1], O[NL]);
        d[0] = sub3_cc(t[0], P.n[0]);
#pragma unroll
        for (int k = 1; k < NL; k++) d[k] = subc3_cc(t[k], P.n[k]);
        uint32_t nb = subc3(t[NL], 0);
        bool take = (nb != 0xffffffffu);
#pragma unroll
        for (int k = 0; k < NL; k++) r[q][k] = take ? d[k] : t[k];
    }
}

// Measured on B200: the dedicated squaring wins from ~20 limbs up (1024-bit: +6 % in stage 1); at 13 limbs
// its extra shifts/adds and short chains cost more than the 22 % fewer products save (7.6 -> 6.9 Tprod/s),
// and beyond 32 limbs the 2NL-limb square no longer fits the register file.
template <int NL> struct UseSqr { static constexpr bool value = (NL >= 20 && NL <= 32); };

template <int NL>
__device__ __forceinline__ void mont_sqr(uint32_t (&r)[NL], const uint32_t (&a)[NL], const ModParams<NL> &P)
{
    if (!UseSqr<NL>::value || ECM_SPECIAL) { mont_mul<NL>(r, a, a, P); return; }
    uint32_t rr[1][NL], aa[1][NL];
#pragma unroll
    for (int k = 0; k < NL; k++) aa[0][k] = a[k];
    mont_sqr_k<NL, 1>(rr, aa, P);
#pragma unroll
    for (int k = 0; k < NL; k++) r[k] = rr[0][k];
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

// ---- wide moduli (NL > 32): operands streamed from shared memory ------------------------------------
// Holding a (NL), b (NL) and the two accumulators (2NL+4) in registers no longer fits at 48/64 limbs
// (the compiler spills and the multiply drops to 0.4 of the roof).  Here only the accumulators are
// register-resident; a[j] and b[i] are read from the slot file when needed (NL^2 + NL shared loads per
// multiply against 2NL^2 IMAD.WIDE: the LSU is ~50 % busy) and the reduction tail works in place.
template <int STRIDE>
struct SmemLimbs {
    const uint32_t *p;
    __device__ __forceinline__ uint32_t operator[](int j) const { return p[j * STRIDE]; }
};

// canonical result of T' = (E>>32) + O (< 2N), written limb by limb through `put(k, value)`
template <int NL, int W, class PUT>
__device__ __forceinline__ void mont_tail_inplace(uint32_t (&E)[W], uint32_t (&O)[W], const ModParams<NL> &P, PUT put)
{
    add_cc(O[0], E[1]);
#pragma unroll
    for (int k = 1; k < NL; k++) addc_cc(O[k], E[k + 1]);
    addc(O[NL], E[NL + 1]);
    // does T' >= N ?  (borrow of T' - N, results discarded)
    (void)sub3_cc(O[0], P.n[0]);
#pragma unroll
    for (int k = 1; k < NL; k++) (void)subc3_cc(O[k], P.n[k]);
    const uint32_t nb = subc3(O[NL], 0);
    const uint32_t mask = (nb != 0xffffffffu) ? 0xffffffffu : 0u;
    uint32_t v = sub3_cc(O[0], P.n[0] & mask);
    put(0, v);
#pragma unroll
    for (int k = 1; k < NL; k++) { v = (k == NL - 1) ? subc3(O[k], P.n[k] & mask) : subc3_cc(O[k], P.n[k] & mask); put(k, v); }
}

template <int NL, class AT, class BT, class PUT>
__device__ __forceinline__ void mont_mul_stream(const AT &a, const BT &b, const ModParams<NL> &P, PUT put)
{
    constexpr int W = MontW<NL>::W;
    uint32_t X[W], Y[W];
#pragma unroll
    for (int k = 0; k < W; k++) { X[k] = 0; Y[k] = 0; }
    auto row = [&](uint32_t (&Eo)[W], uint32_t (&Oo)[W], uint32_t bi) {
        uint32_t e1 = Eo[1];
#pragma unroll
        for (int k = 0; k < W - 2; k++) Eo[k] = Eo[k + 2];
        Eo[W - 2] = 0; Eo[W - 1] = 0;
        add_cc(Oo[0], e1);
        mad_row<NL, W, 1, true>(Eo, a, bi);
        mad_row<NL, W, 0, false>(Oo, a, bi);
        uint32_t m = mul_lo(Oo[0], P.m0inv);
        mad_row<NL, W, 1, false>(Eo, P.n, m);
        mad_row<NL, W, 0, false>(Oo, P.n, m);
    };
#pragma unroll
    for (int i = 0; i < NL; i++) {
        if ((i & 1) == 0) row(X, Y, b[i]); else row(Y, X, b[i]);
    }
    if (NL % 2 == 0) mont_tail_inplace<NL, W>(X, Y, P, put); else mont_tail_inplace<NL, W>(Y, X, P, put);
}

// (a+b) mod N and (a-b) mod N with streamed operands, canonical
template <int NL, class AT, class BT, class PUT>
__device__ __forceinline__ void mod_add_stream(const AT &a, const BT &b, const ModParams<NL> &P, PUT put)
{
    uint32_t t[NL];
    t[0] = add3_cc(a[0], b[0]);
#pragma unroll
    for (int k = 1; k < NL; k++) t[k] = addc3_cc(a[k], b[k]);
    const uint32_t c = addc3(0, 0);
    (void)sub3_cc(t[0], P.n[0]);
#pragma unroll
    for (int k = 1; k < NL; k++) (void)subc3_cc(t[k], P.n[k]);
    const uint32_t nb = subc3(c, 0);
    const uint32_t mask = (nb != 0xffffffffu) ? 0xffffffffu : 0u;
    uint32_t v = sub3_cc(t[0], P.n[0] & mask);
    put(0, v);
#pragma unroll
    for (int k = 1; k < NL; k++) { v = (k == NL - 1) ? subc3(t[k], P.n[k] & mask) : subc3_cc(t[k], P.n[k] & mask); put(k, v); }
}
template <int NL, class AT, class BT, class PUT>
__device__ __forceinline__ void mod_sub_stream(const AT &a, const BT &b, const ModParams<NL> &P, PUT put)
{
    uint32_t t[NL];
    t[0] = sub3_cc(a[0], b[0]);
#pragma unroll
    for (int k = 1; k < NL; k++) t[k] = subc3_cc(a[k], b[k]);
    const uint32_t bo = subc3(0, 0);
    uint32_t v = add3_cc(t[0], P.n[0] & bo);
    put(0, v);
#pragma unroll
    for (int k = 1; k < NL; k++) { v = (k == NL - 1) ? addc3(t[k], P.n[k] & bo) : addc3_cc(t[k], P.n[k] & bo); put(k, v); }
}

}  // inline namespace ECM_VNS
}  // namespace ecmb200
