// special.hpp -- input classification of the command-line driver: Mersenne-like inputs
// (N | 2^k-1, N | 2^k+1, N = 2^k-c with small c), the job main.c:405-521 does before any curve runs.
#pragma once
#include "gmp.h"

struct SpecialForm {
    long kind = 0;        // 0 generic (REDC on N); 1: 2^k-1; -1: 2^k+1; c > 1: 2^k-c
    int k = 0;            // exponent of the base number
};

// Classifies n, divides out the algebraic factors of 2^k+-1 (n is updated in place, main.c:444-457) and
// decides between the special base and plain REDC (main.c:505-521).  base receives 2^k-c / 2^k+1 when
// the special form is used.  Prints the reference's messages.
SpecialForm classify_input(mpz_t n, mpz_t base);
