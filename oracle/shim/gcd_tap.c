/* TEST INFRASTRUCTURE ONLY.
 * LD_PRELOAD interposer used when generating golden vectors from the compiled
 * reference (oracle/_ref): logs the operands of every mpz_gcd() call, which is
 * how check_factor (ecm.c:2542-2557) sees the stage-1 Z and the stage-2
 * accumulator of every lane (both still in Montgomery form).  The reference
 * binary itself is unmodified.  Log file: $GCD_TAP_FILE (append). */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include "gmp.h"
#undef mpz_gcd
void __gmpz_gcd(mpz_ptr g, mpz_srcptr a, mpz_srcptr b)
{
    static void (*real)(mpz_ptr, mpz_srcptr, mpz_srcptr) = NULL;
    const char *fn = getenv("GCD_TAP_FILE");
    if (!real) real = (void (*)(mpz_ptr, mpz_srcptr, mpz_srcptr))dlsym(RTLD_NEXT, "__gmpz_gcd");
    if (fn) {
        FILE *f = fopen(fn, "a");
        if (f) { gmp_fprintf(f, "gcd %Zx %Zx\n", a, b); fclose(f); }
    }
    real(g, a, b);
}
