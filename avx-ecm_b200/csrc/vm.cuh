// vm.cuh -- the curve-arithmetic "field-op machine".
//
// One thread owns one curve.  All curves of a batch execute the same op stream (the PRAC
// chain for a prime does not depend on the curve), so the host compiles stage 1 / stage 2 into
// a compact stream of macro-ops once and every thread of the grid interprets it with zero
// divergence.  A macro-op expands (from a table in constant memory) into field micro-ops on
// numbered slots (HybridSlots below: points in the L2-resident global state, hot scratch values in
// shared memory as [slot][limb][thread], bank = thread, conflict free); operands are pulled into
// registers for the multiply.  One copy each of the unrolled single and dual Montgomery multiply
// exists in the kernel, which keeps the instruction footprint small.
//
// Reference semantics reproduced (Appendix A of SURVEY.md):
//   vec_add        ecm.c:407-443      vec_duplicate  ecm.c:445-457
//   prac() body    ecm.c:603-873      ecm_stage1     ecm.c:1806-1854
// Pointer swaps of the reference (ecm.c:624-629, 704-711) cost nothing here: the host tracks
// which physical point slot currently plays A, B, C, T and encodes that permutation in the
// macro-op byte.
#pragma once
#include "mp.cuh"

namespace ecmb200 {
inline namespace ECM_VNS {

// ---- micro-ops --------------------------------------------------------------------------
enum : uint32_t { U_MUL = 0, U_SQR = 1, U_ADD = 2, U_SUB = 3, U_ADDSUB = 4, U_COPY = 5, U_MUL2 = 6, U_END = 15 };
// symbolic operands: 0..7 point coordinates resolved through the permutation, 8.. fixed slots
enum : uint32_t { AX = 0, AZ, BX, BZ, CX, CZ, TX, TZ, S1 = 8, D1, S2, D2, SP, NSLOT_S1 };
#define UOP(op, d, d2, x, y) ((uint32_t)(op) | ((uint32_t)(d) << 4) | ((uint32_t)(d2) << 8) | ((uint32_t)(x) << 12) | ((uint32_t)(y) << 16))
// two independent multiplies in one micro-op: d = x*y and e = u*v (operands are all read before either result is written)
#define UOP2(d, x, y, e, u, v) ((uint32_t)U_MUL2 | ((uint32_t)(d) << 4) | ((uint32_t)(x) << 12) | ((uint32_t)(y) << 16) | \
                                ((uint32_t)(e) << 8) | ((uint32_t)(u) << 20) | ((uint32_t)(v) << 24))

// vec_add(Pin -> Pout) with the current s1,d1,s2,d2; temporaries reuse d1 / s1 (dead after
// their first use):  d1 = d1*s2 ; s1 = s1*d2 ; (d1,s1) = (d1+s1, d1-s1) ; d1 = d1^2 ; s1 = s1^2 ;
// Pout.X = d1*Pin.Z ; Pout.Z = s1*Pin.X                                   (ecm.c:417-439)
#define PROG_ADD(PinX, PinZ, PoutX, PoutZ)                                                     \
    UOP2(D1, D1, S2, S1, S1, D2), UOP(U_ADDSUB, D1, S1, D1, S1), UOP2(D1, D1, D1, S1, S1, S1),    \
    UOP2(PoutX, D1, PinZ, PoutZ, S1, PinX)
// vec_duplicate(s,d -> P), scratch = a dead sum slot:  d = d^2 ; s = s^2 ; P.X = d*s ;
// tmp = s-d ; s = tmp*sp ; s = s+d ; P.Z = s*tmp                           (ecm.c:447-454)
#define PROG_DUP(s, d, tmp, PX, PZ)                                                            \
    UOP2(d, d, d, s, s, s), UOP(U_SUB, tmp, 0, s, d), UOP2(PX, d, s, s, tmp, SP),                  \
    UOP(U_ADD, s, 0, s, d), UOP(U_MUL, PZ, 0, s, tmp)
#define PROG_SUMS(PX, PZ, s, d) UOP(U_ADDSUB, s, d, PX, PZ)

// ---- stage-1 macro-ops (low 3 bits of the stream byte; high 5 bits = permutation index) ----
enum : uint32_t { M_DBL = 0, M_INIT = 1, M_C3 = 2, M_C4 = 3, M_C5 = 4, M_C9 = 5, M_FINAL = 6, M_NOP = 7 };
#define MAXPROG 20
static __constant__ uint32_t c_prog_s1[8][MAXPROG] = {
    // M_DBL: P (held in logical T) doubled in place                      (ecm.c:1816-1822)
    { PROG_SUMS(TX, TZ, S1, D1), PROG_DUP(S1, D1, S2, TX, TZ), UOP(U_END, 0, 0, 0, 0) },
    // M_INIT: P is in logical B (the host renamed it); C = P ; A = 2P      (ecm.c:603-613)
    { UOP(U_COPY, CX, 0, BX, 0), UOP(U_COPY, CZ, 0, BZ, 0), PROG_SUMS(BX, BZ, S1, D1),
      PROG_DUP(S1, D1, S2, AX, AZ), UOP(U_END, 0, 0, 0, 0) },
    // M_C3: T = B + A (C) ; host then rotates (B,T,C)                     (ecm.c:683-713)
    { PROG_SUMS(BX, BZ, S1, D1), PROG_SUMS(AX, AZ, S2, D2), PROG_ADD(CX, CZ, TX, TZ), UOP(U_END, 0, 0, 0, 0) },
    // M_C4: B = B + A (C) ; A = 2A                                        (ecm.c:714-726)
    { PROG_SUMS(BX, BZ, S1, D1), PROG_SUMS(AX, AZ, S2, D2), PROG_ADD(CX, CZ, BX, BZ),
      PROG_DUP(S2, D2, S1, AX, AZ), UOP(U_END, 0, 0, 0, 0) },
    // M_C5: C = C + A (B) ; A = 2A                                        (ecm.c:728-740)
    { PROG_SUMS(CX, CZ, S1, D1), PROG_SUMS(AX, AZ, S2, D2), PROG_ADD(BX, BZ, CX, CZ),
      PROG_DUP(S2, D2, S1, AX, AZ), UOP(U_END, 0, 0, 0, 0) },
    // M_C9: C = C + B (A) ; B = 2B                                        (ecm.c:853-865)
    { PROG_SUMS(CX, CZ, S1, D1), PROG_SUMS(BX, BZ, S2, D2), PROG_ADD(AX, AZ, CX, CZ),
      PROG_DUP(S2, D2, S1, BX, BZ), UOP(U_END, 0, 0, 0, 0) },
    // M_FINAL: P = A + B (C), written to logical T                        (ecm.c:868-873)
    { PROG_SUMS(AX, AZ, S1, D1), PROG_SUMS(BX, BZ, S2, D2), PROG_ADD(CX, CZ, TX, TZ), UOP(U_END, 0, 0, 0, 0) },
    { UOP(U_END, 0, 0, 0, 0) },
};

// permutation index -> physical point slot of (A,B,C,T), 2 bits each (A lowest)
static __constant__ uint8_t c_perm[24] = {
    0xE4, 0xB4, 0xD8, 0x78, 0x9C, 0x6C, 0xE1, 0xB1, 0xC9, 0x39, 0x8D, 0x2D,
    0xD2, 0x72, 0xC6, 0x36, 0x4E, 0x1E, 0x93, 0x63, 0x87, 0x27, 0x4B, 0x1B };

__device__ __forceinline__ uint32_t resolve(uint32_t sym, uint32_t permbits)
{
    return (sym < 8) ? (2u * ((permbits >> (2u * (sym >> 1))) & 3u) + (sym & 1u)) : sym;
}

// Hybrid slot file of the stage-1 machine: the four points (slots 0..7) stay in the batch state
// in global memory -- L2-resident, [group][slot][limb][STRIDE] so that limb offsets are immediates
// and a warp reads 128 contiguous bytes -- while the five hot scratch slots (s1,d1,s2,d2,sp) live in
// shared memory.  A multiply takes thousands of cycles, so the L2 latency of the point reads is
// hidden by the other resident warps, and the small shared footprint (5 slots) is what lets 14-16
// warps share an SM instead of 10.
template <int NL, int STRIDE, int NGLOBAL = 8, int NSMEM = 5>
struct HybridSlots {
    uint32_t *gl;     // state of this block's group + threadIdx.x
    uint32_t *sm;     // smem + threadIdx.x
    // wide moduli: pointer to a slot's limbs (stride STRIDE words); global slots are first copied into
    // one of two shared staging slots (shared indices 5 and 6)
    __device__ __forceinline__ uint32_t *ptr(uint32_t slot) const
    {
        return slot < NGLOBAL ? gl + slot * (NL * STRIDE) : sm + (slot - NGLOBAL) * (NL * STRIDE);
    }
    __device__ __forceinline__ const uint32_t *operand(uint32_t slot, int staging) const
    {
        if (slot >= NGLOBAL) return sm + (slot - NGLOBAL) * (NL * STRIDE);
        uint32_t *dst = sm + (NSMEM + staging) * (NL * STRIDE);
        const uint32_t *src = gl + slot * (NL * STRIDE);
#pragma unroll 8
        for (int k = 0; k < NL; k++) dst[k * STRIDE] = src[k * STRIDE];
        return dst;
    }
    __device__ __forceinline__ void load(uint32_t (&r)[NL], uint32_t slot) const
    {
        if (slot < NGLOBAL) {
            const uint32_t *p = gl + slot * (NL * STRIDE);
#pragma unroll
            for (int k = 0; k < NL; k++) r[k] = p[k * STRIDE];
        } else {
            const uint32_t *p = sm + (slot - NGLOBAL) * (NL * STRIDE);
#pragma unroll
            for (int k = 0; k < NL; k++) r[k] = p[k * STRIDE];
        }
    }
    __device__ __forceinline__ void store(uint32_t slot, const uint32_t (&r)[NL]) const
    {
        if (slot < NGLOBAL) {
            uint32_t *p = gl + slot * (NL * STRIDE);
#pragma unroll
            for (int k = 0; k < NL; k++) p[k * STRIDE] = r[k];
        } else {
            uint32_t *p = sm + (slot - NGLOBAL) * (NL * STRIDE);
#pragma unroll
            for (int k = 0; k < NL; k++) p[k * STRIDE] = r[k];
        }
    }
};

template <int NL, int STRIDE>
__device__ __forceinline__ void exec_uop_stream(const HybridSlots<NL, STRIDE> &S, uint32_t u, uint32_t op, uint32_t d,
                                                uint32_t x, uint32_t y, uint32_t permbits, const ModParams<NL> &P)
{
    if (op == U_MUL2 || op == U_MUL || op == U_SQR) {
        const int n = (op == U_MUL2) ? 2 : 1;
#pragma unroll 1
        for (int h = 0; h < n; h++) {
            const uint32_t xs = h ? resolve((u >> 20) & 15u, permbits) : x;
            const uint32_t ys = h ? resolve((u >> 24) & 15u, permbits) : y;
            const uint32_t ds = h ? resolve((u >> 8) & 15u, permbits) : d;
            // a stays in registers for the whole product, b[i] is fetched once per row
            const uint32_t *pa = S.ptr(xs);
            uint32_t areg[NL];
#pragma unroll
            for (int k = 0; k < NL; k++) areg[k] = pa[k * STRIDE];
            SmemLimbs<STRIDE> B{S.operand(ys, 1)};
            uint32_t *dst = S.ptr(ds);
            mont_mul_stream<NL>(areg, B, P, [&](int k, uint32_t val) { dst[k * STRIDE] = val; });
        }
    } else if (op == U_ADDSUB || op == U_ADD || op == U_SUB) {
        SmemLimbs<STRIDE> A{S.operand(x, 0)}, B{S.operand(y, 1)};
        // Each routine reads every source limb before it writes its first result limb, so a destination
        // may alias a source.  For the in-place pair (d1,s1) = (d1+s1, d1-s1) the sum waits in registers
        // until the difference has been formed from the untouched sources.
        if (op == U_ADDSUB) {
            uint32_t rs[NL];
            mod_add_stream<NL>(A, B, P, [&](int k, uint32_t val) { rs[k] = val; });
            uint32_t *dst2 = S.ptr(resolve((u >> 8) & 15u, permbits));
            mod_sub_stream<NL>(A, B, P, [&](int k, uint32_t val) { dst2[k * STRIDE] = val; });
            uint32_t *dst = S.ptr(d);
#pragma unroll
            for (int k = 0; k < NL; k++) dst[k * STRIDE] = rs[k];
        } else if (op == U_ADD) {
            uint32_t *dst = S.ptr(d);
            mod_add_stream<NL>(A, B, P, [&](int k, uint32_t val) { dst[k * STRIDE] = val; });
        } else {
            uint32_t *dst = S.ptr(d);
            mod_sub_stream<NL>(A, B, P, [&](int k, uint32_t val) { dst[k * STRIDE] = val; });
        }
    } else {  // U_COPY
        const uint32_t *src = S.ptr(x);
        uint32_t *dst = S.ptr(d);
#pragma unroll 8
        for (int k = 0; k < NL; k++) dst[k * STRIDE] = src[k * STRIDE];
    }
}

// Execute one micro-op on the slot file.
template <int NL, class SlotsT>
__device__ __forceinline__ void exec_uop(const SlotsT &S, uint32_t u, uint32_t permbits,
                                         const ModParams<NL> &P)
{
    const uint32_t op = u & 15u;
    const uint32_t d = resolve((u >> 4) & 15u, permbits);
    const uint32_t x = resolve((u >> 12) & 15u, permbits);
    const uint32_t y = resolve((u >> 16) & 15u, permbits);
    if constexpr (NL > 32) {
        // wide moduli: operands stay in shared memory and are streamed into the multiply
        exec_uop_stream<NL>(S, u, op, d, x, y, permbits, P);
        return;
    }
    if (op == U_MUL2) {
        const uint32_t e = resolve((u >> 8) & 15u, permbits);
        const uint32_t x2 = resolve((u >> 20) & 15u, permbits);
        const uint32_t y2 = resolve((u >> 24) & 15u, permbits);
        if (NL <= 16) {                          // both products in flight: registers allow it
            uint32_t a0[NL], b0[NL], a1[NL], b1[NL], r0[NL], r1[NL];
            S.load(a0, x); S.load(b0, y); S.load(a1, x2); S.load(b1, y2);
            mont_mul2<NL>(r0, a0, b0, r1, a1, b1, P);
            S.store(d, r0); S.store(e, r1);
        } else {                                 // wide operands: one after the other through the same body
            uint32_t a0[NL], b0[NL], r0[NL];
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const uint32_t xs = h ? x2 : x, ys = h ? y2 : y;
                S.load(a0, xs);
                if (xs == ys && UseSqr<NL>::value) mont_sqr<NL>(r0, a0, P);
                else { S.load(b0, ys); mont_mul<NL>(r0, a0, b0, P); }
                S.store(h ? e : d, r0);
            }
        }
        return;
    }
    uint32_t a[NL], b[NL], r[NL];
    S.load(a, x);
    if (op == U_MUL || op == U_SQR) {
        S.load(b, y);
        mont_mul<NL>(r, a, b, P);
        S.store(d, r);
    } else if (op == U_ADDSUB) {
        const uint32_t d2 = resolve((u >> 8) & 15u, permbits);
        S.load(b, y);
        mod_add<NL>(r, a, b, P);
        S.store(d, r);
        mod_sub<NL>(r, a, b, P);
        S.store(d2, r);
    } else if (op == U_ADD) {
        S.load(b, y);
        mod_add<NL>(r, a, b, P);
        S.store(d, r);
    } else if (op == U_SUB) {
        S.load(b, y);
        mod_sub<NL>(r, a, b, P);
        S.store(d, r);
    } else {  // U_COPY
        S.store(d, a);
    }
}

}  // inline namespace ECM_VNS
}  // namespace ecmb200
