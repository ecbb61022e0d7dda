// special.cpp -- see special.hpp.  Citations are file:line into the reference tree.
#include "special.hpp"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

// The factor of 2^k + sign (sign = -1 or +1) that survives after removing the factors it shares with
// 2^d + sign for the proper divisors d = k/q (q odd prime): with m = k / (distinct odd primes of k),
// the alternating product over subsets T of those primes of (2^(m*prod T) + sign)^(+-1).  This is what
// find_primitive_factor (main.c:187-358) computes for base 2; like it, at most three distinct odd primes.
static bool primitive_part(mpz_t out, int k, int sign)
{
    std::vector<int> odd;
    int e = k;
    for (int q = 2; q < 1000 && e > 1; q++) {
        bool hit = false;
        while (e % q == 0) { e /= q; hit = true; }
        if (hit && (q & 1)) odd.push_back(q);
    }
    if (odd.size() > 3) { printf("gen: too many distinct odd factors in exponent!\n"); return false; }
    int m = k;
    for (int q : odd) m /= q;
    mpz_t num, den, term;
    mpz_init(num); mpz_init(den); mpz_init(term);
    mpz_set_ui(num, 1); mpz_set_ui(den, 1);
    for (unsigned mask = 0; mask < (1u << odd.size()); mask++) {
        int t = 1, bits = 0;
        for (size_t i = 0; i < odd.size(); i++) if (mask >> i & 1) { t *= odd[i]; bits++; }
        mpz_set_ui(term, 1); mpz_mul_2exp(term, term, (mp_bitcnt_t)m * t);
        if (sign < 0) mpz_sub_ui(term, term, 1); else mpz_add_ui(term, term, 1);
        if (((int)odd.size() - bits) % 2 == 0) mpz_mul(num, num, term); else mpz_mul(den, den, term);
    }
    mpz_tdiv_q(out, num, den);
    mpz_clear(num); mpz_clear(den); mpz_clear(term);
    return true;
}

static unsigned ref_maxbits(size_t bits)            // main.c:465-499, DIGITBITS = 52
{
    unsigned mb = 208;
    while (mb <= bits) mb += 208;
    return mb;
}

SpecialForm classify_input(mpz_t n, mpz_t base)
{
    SpecialForm f;
    mpz_t r, g;
    mpz_init(r); mpz_init(g);
    int size_n = (int)mpz_sizeinbase(n, 2);
    for (int i = size_n - 1; i < 2048; i++) {                     // main.c:408-441
        mpz_set_ui(r, 1); mpz_mul_2exp(r, r, i); mpz_mod(g, r, n);           // 2^i mod n
        if (mpz_cmp_ui(g, 1) == 0 || (i == 0 && mpz_cmp_ui(n, 1) == 0)) { f.kind = 1; f.k = i; break; }
        mpz_add_ui(r, g, 1);
        if (mpz_cmp(r, n) == 0) { f.kind = -1; f.k = i; break; }
        // the reference keeps c in an int (main.c:391,437); larger c would overflow there and are left to REDC
        if (mpz_sizeinbase(g, 2) < 52) { if (mpz_sizeinbase(g, 2) < 32) { f.kind = (long)mpz_get_ui(g); f.k = i; } break; }
    }
    if (f.kind == 1 || f.kind == -1) {                            // main.c:444-457
        if (!primitive_part(g, f.k, f.kind == 1 ? -1 : 1)) exit(1);         // like main.c:229-233
        {
            mpz_tdiv_q(r, n, g);
            // the reference prints the quotient input/primitive (0 when the input is a proper divisor)
            gmp_printf("removing algebraic %s%d factor %Zd\n", mpz_probab_prime_p(g, 3) ? "PRP" : "C",
                       (int)mpz_sizeinbase(r, 10), r);
            mpz_gcd(n, n, g);
        }
    }
    if (f.kind) {                                                 // main.c:505-521
        const double nwords = ref_maxbits(mpz_sizeinbase(n, 2)) / 52, mwords = ref_maxbits((size_t)f.k) / 52;
        if (nwords / mwords < 0.7) {
            printf("Mersenne input 2^%d %c %ld determined to be faster by REDC\n", f.k, f.kind > 0 ? '-' : '+', f.kind);
            f.kind = 0;
        }
    }
    if (f.kind) {
        mpz_set_ui(base, 1); mpz_mul_2exp(base, base, f.k);
        if (f.kind > 0) mpz_sub_ui(base, base, (unsigned long)f.kind); else mpz_add_ui(base, base, 1);
        if (f.kind > 1) printf("Using special pseudo-Mersenne mod for factor of: 2^%d-%ld\n", f.k, f.kind);
        else printf("Using special Mersenne mod for factor of: 2^%d%c1\n", f.k, f.kind > 0 ? '-' : '+');
    } else {
        f.k = (int)mpz_sizeinbase(n, 2);
    }
    mpz_clear(r); mpz_clear(g);
    return f;
}
