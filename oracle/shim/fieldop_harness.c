/* fieldop_harness.c -- TEST INFRASTRUCTURE ONLY.
 * Drives the UNMODIFIED reference's special-form field operators (vecarith52.c:284-2436,
 * 4613-4801, 4970-5090) one call at a time, so that the oracle's claim "lazily reduced
 * representative == canonical residue mod 2^k+-c" can be checked operator by operator.
 * Built by oracle/build_ref.sh into oracle/_ref/fieldop-ref from the reference sources where
 * they lie (main.c is compiled with -Dmain=ref_main; nothing is copied).
 *
 * usage: fieldop-ref nbits c      (c = 1: 2^nbits-1, c = -1: 2^nbits+1, c > 1: 2^nbits-c)
 * stdin lines:  mul A B | sqr A | add A B | sub A B | addsub A B     (hex operands, broadcast to all lanes)
 *               mul8 A0 B0 .. A7 B7      eight independent lanes, eight results
 *               dup X0 Z0 S0 .. X7 Z7 S7 the doubling of ecm_stage1 (ecm.c:1819-1821 + vec_duplicate) on eight
 *                                        lanes with the reference's buffer reuse; prints lane 0's X Z
 * stdout: result(s) in hex, one line per op.                                              */
#include "avx_ecm.h"

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: fieldop-ref nbits c\n"); return 1; }
    int nbits = atoi(argv[1]); long c = atol(argv[2]);
    MAXBITS = 208; while (MAXBITS <= (uint32_t)nbits) MAXBITS += 208;      /* main.c:485-499 */
    NWORDS = MAXBITS / DIGITBITS; NBLOCKS = NWORDS / BLOCKWORDS;
    monty *md = monty_alloc();
    mpz_t m, t, u; mpz_init(m); mpz_init(t); mpz_init(u);
    mpz_set_ui(m, 1); mpz_mul_2exp(m, m, nbits);
    if (c > 0) mpz_sub_ui(m, m, (unsigned long)c); else mpz_add_ui(m, m, 1);
    md->isMersenne = (int)c; md->nbits = nbits;                            /* main.c:597-616 */
    broadcast_mpz_to_vec(md->n, m);
    mpz_set_ui(t, 1); broadcast_mpz_to_vec(md->one, t);
    bignum *a = vecInit(), *b = vecInit(), *r = vecInit(), *r2 = vecInit(), *s = vecInit();
    char op[16], A[4096], B[4096];
    while (scanf("%15s", op) == 1) {
        if (!strcmp(op, "mul8")) {      /* 8 independent lanes: mul8 A0 B0 ... A7 B7 -> 8 results */
            memset(a->data, 0, VECLEN * (2 * NWORDS + 4) * sizeof(base_t));
            memset(b->data, 0, VECLEN * (2 * NWORDS + 4) * sizeof(base_t));
            for (int l = 0; l < VECLEN; l++) {
                if (scanf("%4095s %4095s", A, B) != 2) return 1;
                mpz_set_str(t, A, 16); insert_mpz_to_vec(a, t, l);
                mpz_set_str(t, B, 16); insert_mpz_to_vec(b, t, l);
            }
            vecmulmod52_mersenne(a, b, r, md->n, s, md);
            for (int l = 0; l < VECLEN; l++) { extract_bignum_from_vec_to_mpz(t, r, l, NWORDS); gmp_printf("%Zx ", t); }
            printf("\n");
            continue;
        }
        if (!strcmp(op, "dup")) {       /* dup X Z S: ecm.c:1819-1821 + vec_duplicate (ecm.c:445-457), with its buffer reuse */
            char S[4096];
            bignum *PX = vecInit(), *PZ = vecInit(), *ws = vecInit(), *d1 = vecInit(), *s1 = vecInit();
            bignum *tt1 = vecInit(), *tt2 = vecInit(), *tt3 = vecInit(), *tt4 = vecInit();
            for (int l = 0; l < VECLEN; l++) {
            if (scanf("%4095s %4095s %4095s", A, B, S) != 3) break;
            mpz_set_str(t, A, 16); insert_mpz_to_vec(PX, t, l);
            mpz_set_str(t, B, 16); insert_mpz_to_vec(PZ, t, l);
            mpz_set_str(t, S, 16); insert_mpz_to_vec(ws, t, l);
            }
            vecsubmod52_mersenne(PX, PZ, d1, md);
            vecaddmod52_mersenne(PX, PZ, s1, md);
            vecsqrmod52_mersenne(d1, tt1, md->n, tt4, md);
            vecsqrmod52_mersenne(s1, tt2, md->n, tt4, md);
            vecmulmod52_mersenne(tt1, tt2, PX, md->n, tt4, md);
            vecsubmod52_mersenne(tt2, tt1, tt3, md);
            vecmulmod52_mersenne(tt3, ws, tt2, md->n, tt4, md);
            vecaddmod52_mersenne(tt2, tt1, tt2, md);
            vecmulmod52_mersenne(tt2, tt3, PZ, md->n, tt4, md);
            extract_bignum_from_vec_to_mpz(t, PX, 0, NWORDS); gmp_printf("%Zx", t);
            extract_bignum_from_vec_to_mpz(t, PZ, 0, NWORDS); gmp_printf(" %Zx\n", t);
            continue;
        }
        int two = strcmp(op, "sqr") != 0;
        if (scanf("%4095s", A) != 1) break;
        if (two && scanf("%4095s", B) != 1) break;
        mpz_set_str(t, A, 16); if (two) mpz_set_str(u, B, 16);
        memset(a->data, 0, VECLEN * (2 * NWORDS + 4) * sizeof(base_t));
        memset(b->data, 0, VECLEN * (2 * NWORDS + 4) * sizeof(base_t));
        broadcast_mpz_to_vec(a, t); if (two) broadcast_mpz_to_vec(b, u);
        if (!strcmp(op, "mul")) vecmulmod52_mersenne(a, b, r, md->n, s, md);
        else if (!strcmp(op, "sqr")) vecsqrmod52_mersenne(a, r, md->n, s, md);
        else if (!strcmp(op, "add")) vecaddmod52_mersenne(a, b, r, md);
        else if (!strcmp(op, "sub")) vecsubmod52_mersenne(a, b, r, md);
        else if (!strcmp(op, "addsub")) vec_simul_addsub52_mersenne(a, b, r, r2, md);
        else { fprintf(stderr, "bad op %s\n", op); return 1; }
        extract_bignum_from_vec_to_mpz(t, r, 0, NWORDS); gmp_printf("%Zx", t);
        if (!strcmp(op, "addsub")) { extract_bignum_from_vec_to_mpz(t, r2, 0, NWORDS); gmp_printf(" %Zx", t); }
        printf("\n");
    }
    return 0;
}
