// coop.cuh -- warp-cooperative field arithmetic for wide moduli: one value is striped over L lanes of a warp.
//
// Above ~1024 bits one thread cannot hold a product's operands and accumulators in registers any more (the
// one-thread-per-curve kernels stream b from shared memory and run one warp per scheduler, 0.57 of the IMAD roof at
// 2048 bits).  Here lane p (0 <= p < L) of a group of L adjacent lanes owns limbs [p*M, (p+1)*M) of every value, of N
// and of the running sum, so a 64-limb value costs 16 registers per lane and the macro-op machine of rv.cuh keeps all
// operands of a point addition in registers.  This is the B200 counterpart of the reference's BLOCKWORDS = 4 column
// blocks (vecarith52.c:2438-3074), with the blocks spread over lanes instead of being walked by one thread.
//
// Montgomery product (word-serial CIOS, rows i = 0 .. n-1, n = L*M):
//   * every lane keeps a PRIVATE running sum V_p in the E/O form of mp.cuh (V = E + O*2^32), the global sum being
//     T = sum_p V_p * 2^(32*M*p) in REDUNDANT form: V_p may exceed 2^(32M), nothing flows upwards inside the loop;
//   * per row three warp shuffles: b_i from its owner, the quotient digit m_i from lane 0 (computed EARLY from the low
//     word only, so its latency hides behind the 2M multiplies of the row), and -- the division by 2^32 -- the low word
//     each lane drops, which its lower neighbour adds at limb M-1;
//   * after the last row the lanes' overflow words ripple upwards once (L-1 steps) and the conditional subtraction of N
//     is decided with two warp votes (first differing lane from the top wins).
// Results are canonical in [0, N) like everything else in the engine, hence independent of how the value is striped.
//
// The source compiles for the host as well (carry flag = a variable, shuffles and votes = a lock-step lane emulation),
// which is how tests/test_coop_cpu.py checks every routine against Python integers without a GPU.
#pragma once
#include <stdint.h>
#include "special_fold.hpp"      // sf_* carry-chain primitives: PTX on the device, a flag variable on the host

#ifdef __CUDACC__
#define ECM_CO_HD __host__ __device__ __forceinline__
#else
#define ECM_CO_HD inline
#endif

namespace ecmb200 {
namespace coop {

template <int M> struct AccW { static constexpr int W = (M % 2 == 0) ? M + 2 : M + 3; };

#ifdef __CUDACC__
// the L lanes of a group are adjacent lanes of one warp; all 32 lanes execute every call
template <int L>
struct WarpComm {
    uint32_t part, base;
    __device__ __forceinline__ WarpComm() { const uint32_t lane = threadIdx.x & 31u; part = lane & (L - 1); base = lane & ~(uint32_t)(L - 1); }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, uint32_t src_part) const { return __shfl_sync(0xffffffffu, v, base + src_part); }
    __device__ __forceinline__ uint32_t from_above(uint32_t v) const      // value of lane part+1; 0 for the top lane
    {
        const uint32_t t = __shfl_down_sync(0xffffffffu, v, 1);
        return part == L - 1 ? 0u : t;
    }
    __device__ __forceinline__ uint32_t from_below(uint32_t v) const      // value of lane part-1; 0 for lane 0
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, 1);
        return part == 0 ? 0u : t;
    }
    __device__ __forceinline__ uint32_t vote(bool p) const { return (__ballot_sync(0xffffffffu, p) >> base) & ((1u << L) - 1u); }
};
#endif

ECM_CO_HD uint32_t mul_lo32(uint32_t a, uint32_t b) { return a * b; }

// acc += x[j]*y over the even (PAR = 0) or odd (PAR = 1) limbs j of x; pairs land on acc[j-PAR], acc[j-PAR+1]
template <int M, int W, int PAR, bool FIRST_CARRY>
ECM_CO_HD void mad_row(uint32_t (&acc)[W], const uint32_t (&x)[M], uint32_t y)
{
    constexpr int TOP = (PAR == 0) ? M + 1 : M;
#pragma unroll
    for (int j = PAR; j < M; j += 2) {
        const int k = j - PAR;
        if (j == PAR && !FIRST_CARRY) acc[k] = sf_mad_lo_cc(x[j], y, acc[k]);
        else acc[k] = sf_madc_lo_cc(x[j], y, acc[k]);
        acc[k + 1] = sf_madc_hi_cc(x[j], y, acc[k + 1]);
    }
    constexpr int NPAIR = (M - PAR + 1) / 2;
    constexpr int T0 = 2 * NPAIR;
#pragma unroll
    for (int t = T0; t <= TOP; t++) acc[t] = (t == TOP) ? sf_addc(acc[t], 0) : sf_addc_cc(acc[t], 0);
}

// carry / borrow look-ahead over the L lanes of a group: bit p of G = lane p generates, of P = lane p propagates.
// Returns the mask of lanes that RECEIVE a carry (bit p = carry into lane p); *out = carry out of the top lane.
template <int L>
ECM_CO_HD uint32_t lookahead(uint32_t G, uint32_t P, uint32_t *out)
{
    uint32_t C = 0, c = 0;
#pragma unroll
    for (int p = 0; p < L; p++) {
        C |= c << p;
        c = ((G >> p) | ((P >> p) & c)) & 1u;
    }
    *out = c;
    return C;
}

template <int M>
ECM_CO_HD uint32_t all_ones(const uint32_t (&v)[M])
{
    uint32_t t = v[0];
#pragma unroll
    for (int k = 1; k < M; k++) t &= v[k];
    return t == 0xffffffffu;
}
template <int M>
ECM_CO_HD uint32_t all_zero(const uint32_t (&v)[M])
{
    uint32_t t = v[0];
#pragma unroll
    for (int k = 1; k < M; k++) t |= v[k];
    return t == 0u;
}

// r = T - N if T >= N else T, for T = sum_p t_p * 2^(32Mp) + top * 2^(32n) < 2N (top: only the top lane's is read)
template <int M, int L, class Comm>
ECM_CO_HD void cond_sub_n(uint32_t (&r)[M], const uint32_t (&t)[M], uint32_t top, const uint32_t (&n)[M], const Comm &cm)
{
    uint32_t d[M];
    d[0] = sf_sub_cc(t[0], n[0]);
#pragma unroll
    for (int k = 1; k < M; k++) d[k] = sf_subc_cc(t[k], n[k]);
    const uint32_t bo = sf_subc(0, 0);                      // 0xffffffff: this lane's slice of T is below its slice of N
    bool lt = bo != 0;
    bool gt = !lt && !all_zero<M>(d);
    if (cm.part == L - 1 && top) { gt = true; lt = false; }
    const uint32_t LT = cm.vote(lt), GT = cm.vote(gt);
    const bool take = GT >= LT;                              // the highest lane that differs decides; all equal: T = N
    const uint32_t low = (1u << cm.part) - 1u;
    const uint32_t bin = ((LT & low) > (GT & low)) ? 1u : 0u;   // borrow arriving from the lanes below
    // r = take ? d - bin : t
    uint32_t e[M];
    e[0] = sf_sub_cc(d[0], bin);
#pragma unroll
    for (int k = 1; k < M; k++) e[k] = (k == M - 1) ? sf_subc(d[k], 0) : sf_subc_cc(d[k], 0);
#pragma unroll
    for (int k = 0; k < M; k++) r[k] = take ? e[k] : t[k];
}

// r = a*b*R^-1 mod N (R = 2^(32*L*M)), canonical; a, b canonical.  n = this lane's limbs of N.  M must be even
// (the E/O roles of the accumulators return to the same arrays after the M rows of one b-block).
template <int M, int L, class Comm>
ECM_CO_HD void mont_mul(uint32_t (&r)[M], const uint32_t (&a)[M], const uint32_t (&b)[M], const uint32_t (&n)[M],
                        uint32_t m0inv, const Comm &cm)
{
    static_assert(M % 2 == 0, "limbs per lane must be even");
    constexpr int W = AccW<M>::W;
    uint32_t X[W], Y[W];
#pragma unroll
    for (int k = 0; k < W; k++) { X[k] = 0; Y[k] = 0; }
    // one row: (Eo, Oo) are the E/O roles before the division by 2^32; afterwards they are swapped
    auto row = [&](uint32_t (&Eo)[W], uint32_t (&Oo)[W], uint32_t bi) {
        const uint32_t w = Eo[0], e1 = Eo[1];
        const uint32_t up = cm.from_above(w);                // the upper neighbour's dropped word joins at limb M-1
        // quotient digit, early: only the low word of lane 0's new sum matters
        const uint32_t m = cm.shfl(mul_lo32(Oo[0] + e1 + mul_lo32(a[0], bi), m0inv), 0);
#pragma unroll
        for (int k = 0; k < W - 2; k++) Eo[k] = Eo[k + 2];   // O' = E >> 64 (register renaming)
        Eo[W - 2] = 0; Eo[W - 1] = 0;
        Oo[0] = sf_add_cc(Oo[0], e1);                        // E' = O + e1; its carry enters the O' chain at word 0
        mad_row<M, W, 1, true>(Eo, a, bi);
        mad_row<M, W, 0, false>(Oo, a, bi);
        Oo[M - 1] = sf_add_cc(Oo[M - 1], up);
        Oo[M] = sf_addc_cc(Oo[M], 0);
        Oo[M + 1] = sf_addc(Oo[M + 1], 0);
        mad_row<M, W, 1, false>(Eo, n, m);
        mad_row<M, W, 0, false>(Oo, n, m);
    };
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int q = 0; q < L; q++) {
#pragma unroll
        for (int k = 0; k < M; k += 2) {
            row(X, Y, cm.shfl(b[k], q));
            row(Y, X, cm.shfl(b[k + 1], q));
        }
    }
    // pending division by 2^32: t = (E >> 32) + O, the dropped word goes to the lower neighbour's limb M-1
    uint32_t t[M], hi;
    {
        const uint32_t up = cm.from_above(X[0]);
        t[0] = sf_add_cc(X[1], Y[0]);
#pragma unroll
        for (int k = 1; k < M; k++) t[k] = sf_addc_cc(X[k + 1], Y[k]);
        hi = sf_addc(X[M + 1], Y[M]);
        t[M - 1] = sf_add_cc(t[M - 1], up);
        hi = sf_addc(hi, 0);
    }
    // the lanes' overflow words ripple upwards: lane p adds what lane p-1 holds above its M limbs, once lane p-1 is final
#pragma unroll
    for (int p = 1; p < L; p++) {
        uint32_t inc = cm.from_below(hi);
        inc = (cm.part == p) ? inc : 0u;
        t[0] = sf_add_cc(t[0], inc);
#pragma unroll
        for (int k = 1; k < M; k++) t[k] = sf_addc_cc(t[k], 0);
        hi = sf_addc(hi, 0);
    }
    cond_sub_n<M, L>(r, t, hi, n, cm);
}

// r = (a + b) mod N, canonical
template <int M, int L, class Comm>
ECM_CO_HD void mod_add(uint32_t (&r)[M], const uint32_t (&a)[M], const uint32_t (&b)[M], const uint32_t (&n)[M], const Comm &cm)
{
    uint32_t s[M];
    s[0] = sf_add_cc(a[0], b[0]);
#pragma unroll
    for (int k = 1; k < M; k++) s[k] = sf_addc_cc(a[k], b[k]);
    const uint32_t c = sf_addc(0, 0);
    const uint32_t G = cm.vote(c != 0), P = cm.vote(all_ones<M>(s) != 0);
    uint32_t top;
    const uint32_t cin = (lookahead<L>(G, P, &top) >> cm.part) & 1u;
    s[0] = sf_add_cc(s[0], cin);
#pragma unroll
    for (int k = 1; k < M; k++) s[k] = (k == M - 1) ? sf_addc(s[k], 0) : sf_addc_cc(s[k], 0);
    cond_sub_n<M, L>(r, s, top, n, cm);
}

// r = (a - b) mod N, canonical
template <int M, int L, class Comm>
ECM_CO_HD void mod_sub(uint32_t (&r)[M], const uint32_t (&a)[M], const uint32_t (&b)[M], const uint32_t (&n)[M], const Comm &cm)
{
    uint32_t d[M];
    d[0] = sf_sub_cc(a[0], b[0]);
#pragma unroll
    for (int k = 1; k < M; k++) d[k] = sf_subc_cc(a[k], b[k]);
    const uint32_t bo = sf_subc(0, 0);
    uint32_t G = cm.vote(bo != 0), P = cm.vote(all_zero<M>(d) != 0), neg;
    const uint32_t bin = (lookahead<L>(G, P, &neg) >> cm.part) & 1u;
    d[0] = sf_sub_cc(d[0], bin);
#pragma unroll
    for (int k = 1; k < M; k++) d[k] = (k == M - 1) ? sf_subc(d[k], 0) : sf_subc_cc(d[k], 0);
    // a < b: add N back (the sum wraps past 2^(32n), which is the point)
    const uint32_t mask = neg ? 0xffffffffu : 0u;
    uint32_t s[M];
    s[0] = sf_add_cc(d[0], n[0] & mask);
#pragma unroll
    for (int k = 1; k < M; k++) s[k] = sf_addc_cc(d[k], n[k] & mask);
    const uint32_t c = sf_addc(0, 0);
    G = cm.vote(c != 0); P = cm.vote(all_ones<M>(s) != 0);
    uint32_t top;
    const uint32_t cin = (lookahead<L>(G, P, &top) >> cm.part) & 1u;
    r[0] = sf_add_cc(s[0], cin);
#pragma unroll
    for (int k = 1; k < M; k++) r[k] = (k == M - 1) ? sf_addc(s[k], 0) : sf_addc_cc(s[k], 0);
}

}  // namespace coop
}  // namespace ecmb200
