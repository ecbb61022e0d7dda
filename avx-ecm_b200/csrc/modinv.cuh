// modinv.cuh -- per-thread modular inverse / gcd for odd N (binary extended Euclid, right-shift
// form), operands in registers.  Replaces the host-side GMP calls on the hot path:
//   mpz_invert in batch_invert_pt_* (ecm.c:1925, 2060) and build_one_curve (ecm.c:1745,1759),
//   mpz_gcd in check_factor (ecm.c:2545) and on inversion failure (ecm.c:1932, 2067).
// Inverse and gcd are unique integers, so any exact algorithm reproduces GMP's results.
#pragma once
#include "mp.cuh"

namespace ecmb200 {
inline namespace ECM_VNS {

// in : x in [0,N)
// out: g = gcd(x,N) (g = N for x = 0); if g == 1, inv = x^-1 mod N in [0,N).  Returns g == 1.
// Invariants: A*x == u, C*x == v (mod N); u,v >= 0; A,C in [0,N).
// `n` points at the NL limbs of the modulus (any address space); for WANT_INV = false x may exceed N.
template <int NL, bool WANT_INV>
__device__ __noinline__ bool mod_inverse(uint32_t (&inv)[NL], uint32_t (&g)[NL], const uint32_t (&x)[NL],
                                         const uint32_t *n)
{
    uint32_t u[NL], v[NL], A[NL], C[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) { u[k] = x[k]; v[k] = n[k]; A[k] = (k == 0); C[k] = 0; }

    for (;;) {
        uint32_t nz = 0;
#pragma unroll
        for (int k = 0; k < NL; k++) nz |= u[k];
        if (nz == 0) break;

        const uint32_t odd = 0u - (u[0] & 1u);                 // all-ones if u odd
        // d = u - v ; borrow -> u < v
        uint32_t d[NL];
        d[0] = sub3_cc(u[0], v[0]);
#pragma unroll
        for (int k = 1; k < NL; k++) d[k] = subc3_cc(u[k], v[k]);
        const uint32_t lt = subc3(0, 0);                       // all-ones if u < v
        const uint32_t sw = odd & lt;                          // odd and u < v: swap roles
        // if sw: (u,v) = (v - u, u) else if odd: u = u - v   ; then u >>= 1
        // v - u = -(d): negate d
        uint32_t nd[NL];
        nd[0] = sub3_cc(0, d[0]);
#pragma unroll
        for (int k = 1; k < NL; k++) nd[k] = (k == NL - 1) ? subc3(0, d[k]) : subc3_cc(0, d[k]);
#pragma unroll
        for (int k = 0; k < NL; k++) {
            const uint32_t uo = u[k];
            u[k] = sw ? nd[k] : (odd ? d[k] : uo);
            v[k] = sw ? uo : v[k];
        }
#pragma unroll
        for (int k = 0; k < NL; k++) u[k] = (k == NL - 1) ? (u[k] >> 1) : __funnelshift_r(u[k], u[k + 1], 1);

        if (WANT_INV) {
            // cofactors: if sw: (A,C) = (C - A, A) else if odd: A = A - C ; then A = A/2 mod N
            uint32_t t[NL];
            // t = A - C mod N
            t[0] = sub3_cc(A[0], C[0]);
#pragma unroll
            for (int k = 1; k < NL; k++) t[k] = subc3_cc(A[k], C[k]);
            const uint32_t bo = subc3(0, 0);
            // tn = C - A mod N = -(A-C) mod N : if A-C borrowed (A<C): C-A = -(t) plain; else N - t (or 0)
            // compute both variants cheaply: m1 = t + (N & bo)  == (A - C) mod N
            uint32_t m1[NL];
            m1[0] = add3_cc(t[0], n[0] & bo);
#pragma unroll
            for (int k = 1; k < NL; k++) m1[k] = (k == NL - 1) ? addc3(t[k], n[k] & bo) : addc3_cc(t[k], n[k] & bo);
            // m2 = (C - A) mod N = (N - m1) if m1 != 0 else 0
            uint32_t m1nz = 0;
#pragma unroll
            for (int k = 0; k < NL; k++) m1nz |= m1[k];
            const uint32_t nzmask = m1nz ? 0xffffffffu : 0u;
            uint32_t m2[NL];
            m2[0] = sub3_cc(n[0] & nzmask, m1[0]);
#pragma unroll
            for (int k = 1; k < NL; k++) m2[k] = (k == NL - 1) ? subc3(n[k] & nzmask, m1[k]) : subc3_cc(n[k] & nzmask, m1[k]);
#pragma unroll
            for (int k = 0; k < NL; k++) {
                const uint32_t ao = A[k];
                A[k] = sw ? m2[k] : (odd ? m1[k] : ao);
                C[k] = sw ? ao : C[k];
            }
            // A = A/2 mod N : if A odd add N first (may carry out of NL limbs)
            const uint32_t ao = 0u - (A[0] & 1u);
            A[0] = add3_cc(A[0], n[0] & ao);
#pragma unroll
            for (int k = 1; k < NL; k++) A[k] = addc3_cc(A[k], n[k] & ao);
            const uint32_t top = addc3(0, 0);
#pragma unroll
            for (int k = 0; k < NL; k++) A[k] = (k == NL - 1) ? __funnelshift_r(A[k], top, 1) : __funnelshift_r(A[k], A[k + 1], 1);
        }
    }
    // u == 0: v = gcd
    uint32_t rest = 0;
#pragma unroll
    for (int k = 0; k < NL; k++) { g[k] = v[k]; if (k) rest |= v[k]; }
    const bool ok = (rest == 0) && (v[0] == 1);
    if (WANT_INV) {
#pragma unroll
        for (int k = 0; k < NL; k++) inv[k] = C[k];
    }
    return ok;
}

}  // inline namespace ECM_VNS
}  // namespace ecmb200
