/* Declarations-only stand-in for <gmp.h> (host CLI of the B200 engine).
 *
 * The dev container and the
 * GPU boxes ship the GMP 6.3.0 runtime (/usr/lib/x86_64-linux-gnu/libgmp.so.10)
 * but no development header.  This file declares the subset of GMP's public,
 * documented C API that the reference (bbuhrow/avx-ecm) and oracle/ use, so
 * they can be compiled and linked against that runtime library.  It contains
 * no GMP code: every function below resolves to the exported __gmpz_* / __gmp_*
 * symbol of libgmp.so.10.
 */
#ifndef ECM_B200_CLI_GMP_H
#define ECM_B200_CLI_GMP_H

#include <stdio.h>
#include <stddef.h>
#include <stdarg.h>

#ifdef __cplusplus
extern "C" {
#endif

#define __GNU_MP_VERSION 6
#define __GNU_MP_VERSION_MINOR 3
#define __GNU_MP_VERSION_PATCHLEVEL 0
#define GMP_LIMB_BITS 64
#define GMP_NUMB_BITS 64

typedef unsigned long int mp_limb_t;
typedef long int mp_limb_signed_t;
typedef unsigned long int mp_bitcnt_t;
typedef long int mp_size_t;

typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t *_mp_d;
} __mpz_struct;

typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

/* opaque, generously sized (the real struct is 32 bytes on LP64) */
typedef struct {
    mpz_t _mp_seed;
    int _mp_alg;
    union { void *_mp_lc; } _mp_algdata;
    unsigned char _pad[32];
} __gmp_randstate_struct;
typedef __gmp_randstate_struct gmp_randstate_t[1];

#define mpz_sgn(Z) ((Z)->_mp_size < 0 ? -1 : (Z)->_mp_size > 0)
#define mpz_odd_p(z) (((z)->_mp_size != 0) & (int)(z)->_mp_d[0])
#define mpz_even_p(z) (!mpz_odd_p(z))

#define mpz_init __gmpz_init
void mpz_init(mpz_ptr);
#define mpz_init2 __gmpz_init2
void mpz_init2(mpz_ptr, mp_bitcnt_t);
#define mpz_clear __gmpz_clear
void mpz_clear(mpz_ptr);
#define mpz_set __gmpz_set
void mpz_set(mpz_ptr, mpz_srcptr);
#define mpz_set_ui __gmpz_set_ui
void mpz_set_ui(mpz_ptr, unsigned long);
#define mpz_set_si __gmpz_set_si
void mpz_set_si(mpz_ptr, long);
#define mpz_set_str __gmpz_set_str
int mpz_set_str(mpz_ptr, const char *, int);
#define mpz_get_ui __gmpz_get_ui
unsigned long mpz_get_ui(mpz_srcptr);
#define mpz_get_si __gmpz_get_si
long mpz_get_si(mpz_srcptr);
#define mpz_get_str __gmpz_get_str
char *mpz_get_str(char *, int, mpz_srcptr);
#define mpz_sizeinbase __gmpz_sizeinbase
size_t mpz_sizeinbase(mpz_srcptr, int);
#define mpz_cmp __gmpz_cmp
int mpz_cmp(mpz_srcptr, mpz_srcptr);
#define mpz_cmp_ui __gmpz_cmp_ui
int mpz_cmp_ui(mpz_srcptr, unsigned long);
#define mpz_cmp_si __gmpz_cmp_si
int mpz_cmp_si(mpz_srcptr, long);

#define mpz_add __gmpz_add
void mpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_add_ui __gmpz_add_ui
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_sub __gmpz_sub
void mpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_sub_ui __gmpz_sub_ui
void mpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_mul __gmpz_mul
void mpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_mul_ui __gmpz_mul_ui
void mpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_mul_2exp __gmpz_mul_2exp
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_neg __gmpz_neg
void mpz_neg(mpz_ptr, mpz_srcptr);
#define mpz_abs __gmpz_abs
void mpz_abs(mpz_ptr, mpz_srcptr);
#define mpz_tdiv_q __gmpz_tdiv_q
void mpz_tdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_tdiv_r __gmpz_tdiv_r
void mpz_tdiv_r(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_tdiv_qr __gmpz_tdiv_qr
void mpz_tdiv_qr(mpz_ptr, mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_tdiv_ui __gmpz_tdiv_ui
unsigned long mpz_tdiv_ui(mpz_srcptr, unsigned long);
#define mpz_tdiv_q_2exp __gmpz_tdiv_q_2exp
void mpz_tdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_tdiv_r_2exp __gmpz_tdiv_r_2exp
void mpz_tdiv_r_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_mod __gmpz_mod
void mpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
void mpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_pow_ui __gmpz_pow_ui
void mpz_pow_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_powm __gmpz_powm
void mpz_powm(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
#define mpz_sqrt __gmpz_sqrt
void mpz_sqrt(mpz_ptr, mpz_srcptr);
#define mpz_root __gmpz_root
int mpz_root(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_gcd __gmpz_gcd
void mpz_gcd(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_invert __gmpz_invert
int mpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_jacobi __gmpz_jacobi
int mpz_jacobi(mpz_srcptr, mpz_srcptr);
#define mpz_probab_prime_p __gmpz_probab_prime_p
int mpz_probab_prime_p(mpz_srcptr, int);
#define mpz_nextprime __gmpz_nextprime
void mpz_nextprime(mpz_ptr, mpz_srcptr);
#define mpz_fac_ui __gmpz_fac_ui
void mpz_fac_ui(mpz_ptr, unsigned long);
#define mpz_primorial_ui __gmpz_primorial_ui
void mpz_primorial_ui(mpz_ptr, unsigned long);
#define mpz_fib_ui __gmpz_fib_ui
void mpz_fib_ui(mpz_ptr, unsigned long);
#define mpz_lucnum_ui __gmpz_lucnum_ui
void mpz_lucnum_ui(mpz_ptr, unsigned long);
#define mpz_and __gmpz_and
void mpz_and(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_ior __gmpz_ior
void mpz_ior(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_xor __gmpz_xor
void mpz_xor(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_com __gmpz_com
void mpz_com(mpz_ptr, mpz_srcptr);
#define mpz_urandomb __gmpz_urandomb
void mpz_urandomb(mpz_ptr, gmp_randstate_t, mp_bitcnt_t);
#define mpz_urandomm __gmpz_urandomm
void mpz_urandomm(mpz_ptr, gmp_randstate_t, mpz_srcptr);
#define mpz_import __gmpz_import
void mpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
#define mpz_export __gmpz_export
void *mpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);

#define gmp_randinit_default __gmp_randinit_default
void gmp_randinit_default(gmp_randstate_t);
#define gmp_randclear __gmp_randclear
void gmp_randclear(gmp_randstate_t);
#define gmp_randseed_ui __gmp_randseed_ui
void gmp_randseed_ui(gmp_randstate_t, unsigned long);

#define gmp_printf __gmp_printf
int gmp_printf(const char *, ...);
#define gmp_fprintf __gmp_fprintf
int gmp_fprintf(FILE *, const char *, ...);
#define gmp_sprintf __gmp_sprintf
int gmp_sprintf(char *, const char *, ...);
#define gmp_snprintf __gmp_snprintf
int gmp_snprintf(char *, size_t, const char *, ...);

#ifdef __cplusplus
}
#endif
#endif
