// rv.cuh -- stage 1 as a register-resident macro-op machine (second generation of the field-op machine in vm.cuh).
//
// vm.cuh interprets vec_add / vec_duplicate as field micro-ops on a slot file: every product loads its operands from
// slots and stores its result back (~103 LD/ST per modular product at 13 limbs), and every micro-op pays for operand
// decoding.  ncu showed those ~360 non-multiply instructions per product on the critical path of each warp (3 warps per
// scheduler): fmaheavy 83 % busy.  Here the unit of interpretation is a PHASE -- "prepare four operands, one dual
// product, dispose of the results" -- and the four operand values live in registers across the phases of a point
// operation:
//     vec_add (ecm.c:407-443)        A1: (A0,A1) = (X1-Z1, X1+Z1), (B0,B1) = (X2+Z2, X2-Z2) ; A0*=B0 | A1*=B1
//                                    A2: (A0,A1) = (A0+A1, A0-A1) ; squares
//                                    A3: B0 = Pin.Z, B1 = Pin.X ; A0*=B0 | A1*=B1 -> Pout
//     vec_duplicate (ecm.c:445-457)  D1: (A0,A1) = (s,d) ; squares
//                                    D2: B0 = d^2, B1 = s^2-d^2, A1 = (A+2)/4 ; A0*=B0 (= X) | A1*=B1
//                                    D3: X -> P ; A0 = A1 + B0 ; A0*=B1 (= Z) -> P
// Only point coordinates travel: 6 loads + 2 stores of a value per addition, 1-3 loads + 2 stores per doubling; the sums
// (s2,d2) that a following doubling re-uses (PRAC rules 4, 5, 9) are parked in shared memory.  One body each of the dual
// and the single product exists in the kernel, as before.
//
// The machine is written against a FIELD policy, which is where one-thread-per-curve and warp-cooperative layouts
// differ: SoloField<NL> (mp.cuh: a value is NL registers of one thread) and CoopField<M,L> (coop.cuh: a value is striped
// over L lanes, M registers each).  Same op stream, same state slots, same results.
#pragma once
#include "vm.cuh"
#include "coop.cuh"
#include "rv_prog.hpp"

namespace ecmb200 {
inline namespace ECM_VNS {

// ---- field policies ---------------------------------------------------------------------------------------------
template <int NL>
struct SoloField {
    static constexpr int M = NL, L = 1, LIMBS = NL;
    const ModParams<NL> &P;
    __device__ __forceinline__ explicit SoloField(const ModParams<NL> &p) : P(p) {}
    // (a0, a1) <- (a0*b0, a1*b1)
    __device__ __forceinline__ void mul2(uint32_t (&a0)[M], const uint32_t (&b0)[M], uint32_t (&a1)[M], const uint32_t (&b1)[M]) const
    {
        if (NL <= 16) mont_mul2<NL>(a0, a0, b0, a1, a1, b1, P);
        else { mont_mul<NL>(a0, a0, b0, P); mont_mul<NL>(a1, a1, b1, P); }
    }
    // (a0, a1) <- (a0^2, a1^2): the dedicated dual squaring (1.5n^2+2.5n products each) where it is compiled in
#ifndef ECM_RV_SQR
#define ECM_RV_SQR 0
#endif
    static constexpr bool HAS_SQR2 = ECM_RV_SQR && !ECM_SPECIAL && NL <= 16 && NL >= 4;
    __device__ __forceinline__ void sqr2(uint32_t (&a0)[M], uint32_t (&a1)[M]) const
    {
        uint32_t aa[2][NL], rr[2][NL];
#pragma unroll
        for (int k = 0; k < NL; k++) { aa[0][k] = a0[k]; aa[1][k] = a1[k]; }
        mont_sqr_k<NL, 2>(rr, aa, P);
#pragma unroll
        for (int k = 0; k < NL; k++) { a0[k] = rr[0][k]; a1[k] = rr[1][k]; }
    }
    __device__ __forceinline__ void mul(uint32_t (&a)[M], const uint32_t (&b)[M]) const { mont_mul<NL>(a, a, b, P); }
    __device__ __forceinline__ void add(uint32_t (&r)[M], const uint32_t (&a)[M], const uint32_t (&b)[M]) const { mod_add<NL>(r, a, b, P); }
    __device__ __forceinline__ void sub(uint32_t (&r)[M], const uint32_t (&a)[M], const uint32_t (&b)[M]) const { mod_sub<NL>(r, a, b, P); }
};

template <int M_, int L_>
struct CoopField {
    static constexpr int M = M_, L = L_, LIMBS = M_ * L_;
    uint32_t n[M];
    uint32_t m0inv;
    coop::WarpComm<L> cm;
    static constexpr bool HAS_SQR2 = false;
    __device__ __forceinline__ void sqr2(uint32_t (&)[M], uint32_t (&)[M]) const {}
    __device__ __forceinline__ explicit CoopField(const ModParams<LIMBS> &p)
    {
        m0inv = p.m0inv;
#pragma unroll
        for (int k = 0; k < M; k++) n[k] = p.n[cm.part * M + k];
    }
    __device__ __forceinline__ void mul2(uint32_t (&a0)[M], const uint32_t (&b0)[M], uint32_t (&a1)[M], const uint32_t (&b1)[M]) const
    {
        coop::mont_mul<M, L>(a0, a0, b0, n, m0inv, cm);
        coop::mont_mul<M, L>(a1, a1, b1, n, m0inv, cm);
    }
    __device__ __forceinline__ void mul(uint32_t (&a)[M], const uint32_t (&b)[M]) const { coop::mont_mul<M, L>(a, a, b, n, m0inv, cm); }
    __device__ __forceinline__ void add(uint32_t (&r)[M], const uint32_t (&a)[M], const uint32_t (&b)[M]) const { coop::mod_add<M, L>(r, a, b, n, cm); }
    __device__ __forceinline__ void sub(uint32_t (&r)[M], const uint32_t (&a)[M], const uint32_t (&b)[M]) const { coop::mod_sub<M, L>(r, a, b, n, cm); }
};

#ifdef __CUDACC__
static __constant__ uint32_t c_prog_rv[8][RV_MAXPROG] = { RV_PROGRAMS };
#endif

template <int M, int STRIDE>
__device__ __forceinline__ void rv_load(uint32_t (&r)[M], const uint32_t *p)
{
#pragma unroll
    for (int k = 0; k < M; k++) r[k] = p[k * STRIDE];
}
template <int M, int STRIDE>
__device__ __forceinline__ void rv_store(uint32_t *p, const uint32_t (&r)[M])
{
#pragma unroll
    for (int k = 0; k < M; k++) p[k * STRIDE] = r[k];
}

// State of a group: the same 13 slots of M*STRIDE words as the vm.cuh kernels ([slot][limb][lane], lane = thread of the
// block; a cooperative value has its L*M limbs at limb index k % M of lane curve*L + k / M).  Slots 0..7 are the four
// points, SP the curve constant; the scratch slots 8..11 are not used by this machine.
template <class F, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
k_stage1_rv(const ModParams<F::LIMBS> P, uint32_t *__restrict__ state, const uint8_t *__restrict__ ops,
            uint64_t nops, uint32_t chunk_len, uint32_t groups, uint64_t item0)
{
    constexpr int M = F::M, STRIDE = MAXT, SLOTW = M * STRIDE;
    extern __shared__ uint32_t smem[];                    // the parked sums: [2][M][STRIDE]
    const uint64_t item = item0 + blockIdx.x;
    const uint32_t g = (uint32_t)(item % groups);
    const uint64_t chunk = item / groups;
    uint32_t *gl = state + (size_t)g * (NSLOT_S1 * SLOTW) + threadIdx.x;
    uint32_t *park = smem + threadIdx.x;
    const F field(P);

    uint32_t A0[M], A1[M], B0[M], B1[M];
    uint64_t i = chunk * chunk_len;
    const uint64_t end = (i + chunk_len < nops) ? i + chunk_len : nops;
    const uint32_t *ops32 = reinterpret_cast<const uint32_t *>(ops);
    uint32_t cur = 0;
#pragma unroll 1
    for (; i < end; i++) {
        if ((i & 3) == 0) cur = __ldg(ops32 + (i >> 2));
        const uint32_t byte = cur & 0xffu;
        cur >>= 8;
        const uint32_t permbits = c_perm[byte >> 3];
        const uint32_t *prog = c_prog_rv[byte & 7u];
#pragma unroll 1
        for (int k = 0;; k++) {
            const uint32_t u = prog[k];
            const uint32_t kind = u & 15u;
            if (kind == PH_END) break;
            // physical point slots of the phase's operands: word offset of the X coordinate
            uint32_t *px = gl + ((permbits >> (2u * ((u >> 4) & 3u))) & 3u) * (2 * SLOTW);
            uint32_t *py = gl + ((permbits >> (2u * ((u >> 8) & 3u))) & 3u) * (2 * SLOTW);
            if (kind == PH_A1) {
                uint32_t x[M], z[M];
                rv_load<M, STRIDE>(x, px); rv_load<M, STRIDE>(z, px + SLOTW);
                field.sub(A0, x, z); field.add(A1, x, z);                    // d1, s1
                rv_load<M, STRIDE>(x, py); rv_load<M, STRIDE>(z, py + SLOTW);
                field.add(B0, x, z); field.sub(B1, x, z);                    // s2, d2
                if ((u >> 16) & 1u) { rv_store<M, STRIDE>(park, B0); rv_store<M, STRIDE>(park + SLOTW, B1); }
            } else if (kind == PH_A2) {
                uint32_t t[M];
                field.add(t, A0, A1); field.sub(A1, A0, A1);
#pragma unroll
                for (int j = 0; j < M; j++) { A0[j] = t[j]; B0[j] = t[j]; B1[j] = A1[j]; }
            } else if (kind == PH_A3) {
                rv_load<M, STRIDE>(B0, px + SLOTW); rv_load<M, STRIDE>(B1, px);     // Pin.Z, Pin.X
            } else if (kind == PH_D1L || kind == PH_D1P) {
                if (kind == PH_D1L) {
                    uint32_t x[M], z[M];
                    rv_load<M, STRIDE>(x, px); rv_load<M, STRIDE>(z, px + SLOTW);
                    field.add(A0, x, z); field.sub(A1, x, z);                // s, d
                } else {
                    rv_load<M, STRIDE>(A0, park); rv_load<M, STRIDE>(A1, park + SLOTW);
                }
#pragma unroll
                for (int j = 0; j < M; j++) { B0[j] = A0[j]; B1[j] = A1[j]; }
            } else if (kind == PH_D2) {                                      // A0 = s^2, A1 = d^2
#pragma unroll
                for (int j = 0; j < M; j++) B0[j] = A1[j];
                field.sub(B1, A0, A1);                                       // s^2 - d^2
                rv_load<M, STRIDE>(A1, gl + SP * SLOTW);                     // (A+2)/4
            } else if (kind == PH_D3) {                                      // A0 = X, A1 = sp*t, B0 = d^2, B1 = t
                rv_store<M, STRIDE>(px, A0);
                field.add(A0, A1, B0);
            } else {  // PH_COPY
                rv_load<M, STRIDE>(A0, px); rv_load<M, STRIDE>(A1, px + SLOTW);
                rv_store<M, STRIDE>(py, A0); rv_store<M, STRIDE>(py + SLOTW, A1);
                continue;
            }
            if constexpr (F::L > 1) {
                // one body of the cooperative product in the kernel, looped once (D3) or twice with the operand pairs
                // swapped in between: three inlined copies (5 120 SASS instructions) stalled every fourth issue slot on
                // instruction fetch once all 148 SMs ran the kernel (ncu: no_instruction 1.13 per issue)
                const int np = (kind == PH_D3) ? 1 : 2;
                if (kind == PH_D3) {
#pragma unroll
                    for (int j = 0; j < M; j++) B0[j] = B1[j];
                }
#pragma unroll 1
                for (int h = 0; h < np; h++) {
                    field.mul(A0, B0);
                    if (np == 2) {
#pragma unroll
                        for (int j = 0; j < M; j++) {
                            uint32_t t = A0[j]; A0[j] = A1[j]; A1[j] = t;
                            t = B0[j]; B0[j] = B1[j]; B1[j] = t;
                        }
                    }
                }
                if (kind == PH_D3) rv_store<M, STRIDE>(px + SLOTW, A0);
                else if (kind == PH_A3) { rv_store<M, STRIDE>(py, A0); rv_store<M, STRIDE>(py + SLOTW, A1); }
            } else if (kind == PH_D3) {
                field.mul(A0, B1);
                rv_store<M, STRIDE>(px + SLOTW, A0);
            } else if (F::HAS_SQR2 && (kind == PH_A2 || kind == PH_D1L || kind == PH_D1P)) {
                field.sqr2(A0, A1);
            } else {
                field.mul2(A0, B0, A1, B1);
                if (kind == PH_A3) { rv_store<M, STRIDE>(py, A0); rv_store<M, STRIDE>(py + SLOTW, A1); }
            }
        }
    }
}

}  // inline namespace ECM_VNS
}  // namespace ecmb200
