// modinv_fast.hpp -- modular inverse / gcd for odd N in word-sized batches (replaces the bit-serial binary Euclid of
// modinv.cuh on the hot path: one inversion per giant-step window shift per curve, ecm.c:1925, 2060).
//
// Binary extended Euclid does one shift/subtract of the full-length operands per bit: ~64*NL iterations of ~18*NL
// instructions.  Here 30 iterations at a time run on 62-bit APPROXIMATIONS of (a, b) -- their low 30 bits, which decide
// every parity test exactly, and their top 32 bits, which decide the comparisons almost always -- while the transformation
// is only accumulated as a 2x2 matrix of small factors; the matrix is then applied once to the full-length (a, b) and
// (u, v): 2.1*NL rounds of ~30*25 + ~56*NL instructions (T. Pornin, "Optimized binary GCD for modular inversion", 2020,
// the k = 31 instance, restated here for 32-bit limbs).  A wrong comparison only makes a difference negative, which is
// repaired by negating it; every step keeps gcd(a, b), b odd and a = u*y, b = v*y (mod N), so the loop simply runs until
// a = 0 (the paper's bound of ceil((2*len-1)/30) rounds is what it takes; a cap turns anything else into a failure).
// Inverse and gcd are unique integers: results are identical to GMP's mpz_invert / mpz_gcd and to modinv.cuh.
//
// Plain 64-bit C++ (no PTX): the same source runs on the host, where tests/test_modinv_cpu.py checks it against Python.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ECM_MI_HD __host__ __device__
#else
#define ECM_MI_HD
#endif

namespace ecmb200 {
namespace fastinv {

template <int NL>
ECM_MI_HD inline int bitlen(const uint32_t *x)
{
    for (int k = NL - 1; k >= 0; k--)
        if (x[k]) {
            uint32_t v = x[k];
            int b = 0;
            while (v) { b++; v >>= 1; }
            return 32 * k + b;
        }
    return 0;
}

// bits [pos, pos+32) of x (pos may be negative or reach past the top: missing bits are zero)
template <int NL>
ECM_MI_HD inline uint64_t bits32(const uint32_t *x, int pos)
{
    if (pos <= 0) return pos <= -32 ? 0 : ((uint64_t)x[0] << (-pos)) & 0xffffffffull;
    const int w = pos >> 5, s = pos & 31;
    const uint64_t lo = w < NL ? x[w] : 0, hi = w + 1 < NL ? x[w + 1] : 0;
    return ((lo | (hi << 32)) >> s) & 0xffffffffull;
}

// r = (f*x + g*y) / 2^30 for |f|, |g| <= 2^30 and x, y >= 0; the combination must be divisible by 2^30.  Returns the
// sign (true: the exact result is negative and r holds its magnitude).
template <int NL>
ECM_MI_HD inline bool lincomb_shift(uint32_t *r, int64_t f, const uint32_t *x, int64_t g, const uint32_t *y)
{
    uint32_t t[NL + 1];
    int64_t c = 0;
    for (int k = 0; k < NL; k++) {
        // |f*x_k| < 2^62, twice that plus a carry below 2^32 in magnitude stays inside int64
        const int64_t v = f * (int64_t)x[k] + g * (int64_t)y[k] + c;
        t[k] = (uint32_t)v;
        c = v >> 32;                                   // arithmetic shift: floor division
    }
    t[NL] = (uint32_t)c;                               // sign-extended top word
    const bool neg = c < 0;
    if (neg) {                                         // two's complement -> magnitude
        uint64_t b = 1;
        for (int k = 0; k <= NL; k++) { b += (uint32_t)~t[k]; t[k] = (uint32_t)b; b >>= 32; }
    }
    for (int k = 0; k < NL; k++) r[k] = (t[k] >> 30) | (t[k + 1] << 2);
    return neg;
}

// r = (f*u + g*v) / 2^30 mod n, canonical, for u, v in [0, n): a multiple of n makes the combination divisible by 2^30
// (ninv30 = -n^-1 mod 2^30), the quotient lies in (-2n, 2n) and is brought into [0, n).
template <int NL>
ECM_MI_HD inline void lincomb_mod(uint32_t *r, int64_t f, const uint32_t *u, int64_t g, const uint32_t *v, const uint32_t *n, uint32_t ninv30)
{
    const uint32_t low = (uint32_t)((uint64_t)f * u[0] + (uint64_t)g * v[0]);         // combination mod 2^32
    const int64_t q = (int64_t)((low * ninv30) & 0x3fffffffu);                         // t + q*n == 0 (mod 2^30), 0 <= q < 2^30
    uint32_t t[NL + 2];
    int64_t c = 0;
    for (int k = 0; k < NL; k++) {
        // three terms below 2^62 in magnitude: split the carry so that nothing overflows
        const int64_t a = f * (int64_t)u[k] + g * (int64_t)v[k];                       // |a| < 2^63
        const int64_t b = q * (int64_t)n[k] + c;                                       // 0 <= q*n_k < 2^62, |c| < 2^33
        const uint64_t lo = (uint64_t)(uint32_t)a + (uint64_t)(uint32_t)b;
        t[k] = (uint32_t)lo;
        c = (a >> 32) + (b >> 32) + (int64_t)(lo >> 32);
    }
    t[NL] = (uint32_t)c;
    t[NL + 1] = (uint32_t)(c >> 32);
    bool neg = c < 0;
    // shift right by 30 (exact), keeping the sign in two's complement over NL+1 words
    uint32_t s[NL + 1];
    for (int k = 0; k <= NL; k++) s[k] = (t[k] >> 30) | (t[k + 1] << 2);
    // s in (-2n, 2n): add n while negative, subtract n while >= n
    for (int pass = 0; pass < 2 && neg; pass++) {
        uint64_t cy = 0;
        for (int k = 0; k < NL; k++) { cy += (uint64_t)s[k] + n[k]; s[k] = (uint32_t)cy; cy >>= 32; }
        s[NL] = (uint32_t)(s[NL] + cy);
        neg = (s[NL] >> 31) != 0;
    }
    for (int pass = 0; pass < 2; pass++) {
        uint32_t d[NL];
        int64_t bw = 0;
        for (int k = 0; k < NL; k++) { bw += (int64_t)s[k] - n[k]; d[k] = (uint32_t)bw; bw >>= 32; }
        bw += (int64_t)s[NL];
        if (bw < 0) break;                                                             // s < n
        for (int k = 0; k < NL; k++) s[k] = d[k];
        s[NL] = (uint32_t)bw;
    }
    for (int k = 0; k < NL; k++) r[k] = s[k];
}

// in : y (any value below 2^(32 NL); for WANT_INV it must be in [0, n)), n odd
// out: g = gcd(y, n) (g = n for y = 0); if g == 1 and WANT_INV, inv = y^-1 mod n in [0, n).  Returns g == 1.
template <int NL, bool WANT_INV>
ECM_MI_HD inline bool mod_inverse(uint32_t *inv, uint32_t *g, const uint32_t *y, const uint32_t *n)
{
    uint32_t a[NL], b[NL], u[NL], v[NL], ta[NL], tb[NL];
    for (int k = 0; k < NL; k++) { a[k] = y[k]; b[k] = n[k]; u[k] = (k == 0); v[k] = 0; }
    uint32_t ni = 1;
    for (int i = 0; i < 5; i++) ni *= 2 - n[0] * ni;                                   // n^-1 mod 2^32
    const uint32_t ninv30 = (0u - ni) & 0x3fffffffu;
    for (int round = 0; round < 5 * NL + 8; round++) {
        uint32_t nz = 0;
        for (int k = 0; k < NL; k++) nz |= a[k];
        if (nz == 0) break;
        int len = bitlen<NL>(a);
        const int lb = bitlen<NL>(b);
        if (lb > len) len = lb;
        if (len < 62) len = 62;
        // approximations: low 30 bits, high 32 bits
        uint64_t xa = (a[0] & 0x3fffffffu) | (bits32<NL>(a, len - 32) << 30);
        uint64_t xb = (b[0] & 0x3fffffffu) | (bits32<NL>(b, len - 32) << 30);
        int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
        for (int j = 0; j < 30; j++) {
            if (xa & 1) {
                if (xa < xb) {
                    uint64_t t = xa; xa = xb; xb = t;
                    int64_t s = f0; f0 = f1; f1 = s;
                    s = g0; g0 = g1; g1 = s;
                }
                xa -= xb; f0 -= f1; g0 -= g1;
            }
            xa >>= 1; f1 <<= 1; g1 <<= 1;
        }
        const bool na = lincomb_shift<NL>(ta, f0, a, g0, b);
        const bool nb = lincomb_shift<NL>(tb, f1, a, g1, b);
        for (int k = 0; k < NL; k++) { a[k] = ta[k]; b[k] = tb[k]; }
        if (WANT_INV) {
            if (na) { f0 = -f0; g0 = -g0; }
            if (nb) { f1 = -f1; g1 = -g1; }
            lincomb_mod<NL>(ta, f0, u, g0, v, n, ninv30);
            lincomb_mod<NL>(tb, f1, u, g1, v, n, ninv30);
            for (int k = 0; k < NL; k++) { u[k] = ta[k]; v[k] = tb[k]; }
        }
    }
    uint32_t rest = 0, anz = 0;
    for (int k = 0; k < NL; k++) { g[k] = b[k]; if (k) rest |= b[k]; anz |= a[k]; }
    const bool ok = anz == 0 && rest == 0 && b[0] == 1;
    if (WANT_INV) for (int k = 0; k < NL; k++) inv[k] = v[k];
    return ok;
}

}  // namespace fastinv
}  // namespace ecmb200
