/*
 * ecm_oracle.c -- CPU restatement of avx-ecm's stage-1 / stage-2 algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under avx-ecm_b200/ may include, link or
 * execute this file; it is the checker for tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg.
 *
 * What it is: a scalar (one curve at a time) re-statement of the reference's
 * algorithm with every field operation written as plain big-integer arithmetic
 * mod N (GMP runtime, declarations from oracle/shim/gmp.h).  The reference
 * keeps residues in Montgomery form x*R and every one of its field ops returns
 * the canonical representative in [0,N) (vecarith52.c:3048-3070, 4576-4609,
 * 4712-4721), so the sequence of *true* residues is independent of R and of
 * the limb width; this file tracks the true residues directly.
 *
 * Parity: PINNED.  tests/test_oracle_vs_golden.py checks it byte-for-byte against
 * save_b1.txt produced by the compiled reference (oracle/_ref, built by
 * oracle/build_ref.sh) and against the factors/sigmas in the reference's
 * test_inputs.txt / test.csh; tests/golden/ holds the committed vectors.
 *
 * Citations are file:line into the reference tree.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <inttypes.h>
#include "gmp.h"

/* ------------------------------------------------------------------ */
/* field ops: canonical residues (avx_ecm.h:205-209 operator table)    */
/* ------------------------------------------------------------------ */
typedef struct { mpz_t X, Z; } opt;

typedef struct {
    mpz_t n, s;                       /* modulus, (A+2)/4   (ecm_work.s)      */
    mpz_t sum1, diff1, sum2, diff2;   /* ecm_work sum/diff scratch            */
    mpz_t tt1, tt2, tt3, tt4;         /* ecm_work temporaries                 */
    opt pt1, pt2, pt3, pt4;           /* PRAC A, B, C, T                      */
    opt *Pb;  mpz_t *Pbprod;          /* stage-2 baby steps + scratch         */
    opt *Pa;  mpz_t *Pa_inv, *Paprod; /* stage-2 giant-step window            */
    opt Pad, Pd;
    mpz_t acc;
    uint32_t *map;
    uint32_t D, U, L, R;
    uint32_t amin;
    uint64_t A;
    uint32_t ptadds, ptdups, numinv, paired;
    uint32_t npb;
    int found_inv;                    /* any inversion failure seen           */
    mpz_t rref_inv;                   /* 2^-MAXBITS mod N, MAXBITS of the 52-bit build (main.c:465-483) */
    unsigned long maxbits_ref;        /* that MAXBITS; 0 for special-form inputs (plain residues) */
    mpz_t nchk;                       /* modulus of the factor checks: n, or the input cofactor (nhat) in Mersenne mode */
} owork;

static void fmul(owork *w, mpz_t c, const mpz_t a, const mpz_t b)   /* vecmulmod_ptr */
{ mpz_mul(c, a, b); mpz_tdiv_r(c, c, w->n); }
static void fsqr(owork *w, mpz_t c, const mpz_t a)                  /* vecsqrmod_ptr */
{ mpz_mul(c, a, a); mpz_tdiv_r(c, c, w->n); }
static void fadd(owork *w, mpz_t c, const mpz_t a, const mpz_t b)   /* vecaddmod_ptr */
{ mpz_add(c, a, b); if (mpz_cmp(c, w->n) >= 0) mpz_sub(c, c, w->n); }
static void fsub(owork *w, mpz_t c, const mpz_t a, const mpz_t b)   /* vecsubmod_ptr */
{ mpz_sub(c, a, b); if (mpz_sgn(c) < 0) mpz_add(c, c, w->n); }
/* vecaddsubmod_ptr(a,b,sum,diff); outputs may alias nothing in our uses */
static void faddsub(owork *w, const mpz_t a, const mpz_t b, mpz_t sum, mpz_t diff)
{
    mpz_t t; mpz_init(t);
    mpz_add(t, a, b); if (mpz_cmp(t, w->n) >= 0) mpz_sub(t, t, w->n);
    mpz_sub(diff, a, b); if (mpz_sgn(diff) < 0) mpz_add(diff, diff, w->n);
    mpz_set(sum, t); mpz_clear(t);
}

static void pt_init(opt *p) { mpz_init(p->X); mpz_init(p->Z); }
static void pt_clear(opt *p) { mpz_clear(p->X); mpz_clear(p->Z); }
static void pt_set(opt *d, const opt *s) { mpz_set(d->X, s->X); mpz_set(d->Z, s->Z); }
static void pt_swap(opt *a, opt *b) { opt t = *a; *a = *b; *b = t; }

/* ecm.c:407-443.  The in==out pointer-swap branch produces the same values. */
static void vec_add(owork *w, const opt *Pin, opt *Pout)
{
    mpz_t ox, oz; mpz_init(ox); mpz_init(oz);
    fmul(w, w->tt1, w->diff1, w->sum2);
    fmul(w, w->tt2, w->sum1, w->diff2);
    faddsub(w, w->tt1, w->tt2, w->tt3, w->tt4);
    fsqr(w, w->tt1, w->tt3);
    fsqr(w, w->tt2, w->tt4);
    fmul(w, ox, w->tt1, Pin->Z);
    fmul(w, oz, w->tt2, Pin->X);
    mpz_set(Pout->X, ox); mpz_set(Pout->Z, oz);
    mpz_clear(ox); mpz_clear(oz);
    w->ptadds++;
}

/* ecm.c:445-457 */
static void vec_duplicate(owork *w, const mpz_t insum, const mpz_t indiff, opt *P)
{
    fsqr(w, w->tt1, indiff);
    fsqr(w, w->tt2, insum);
    fmul(w, P->X, w->tt1, w->tt2);
    fsub(w, w->tt3, w->tt2, w->tt1);
    fmul(w, w->tt2, w->tt3, w->s);
    fadd(w, w->tt2, w->tt2, w->tt1);
    fmul(w, P->Z, w->tt2, w->tt3);
    w->ptdups++;
}

/* ------------------------------------------------------------------ */
/* PRAC (ecm.c:459-884), ORIG_PRAC undefined                           */
/* ------------------------------------------------------------------ */
#define ADD 5.5
#define DUP 4.5
#define NV 10
static const double val[NV] = {
    0.61803398874989485, 0.72360679774997897, 0.58017872829546410,
    0.63283980608870629, 0.61242994950949500, 0.62018198080741576,
    0.61721461653440386, 0.61834711965622806, 0.61791440652881789,
    0.61807966846989581 };

/* ecm.c:479-563 */
static double lucas_cost(uint64_t n, double v)
{
    uint64_t d, e, r;
    double c;
    d = n;
    r = (uint64_t)((double)d * v + 0.5);
    if (r >= n) return (ADD * (double)n);
    d = n - r;
    e = 2 * r - n;
    c = DUP + ADD;
    while (d != e) {
        if (d < e) { r = d; d = e; e = r; }
        if ((d + 3) / 4 <= e) { d -= e; c += ADD; }
        else if ((d + e) % 2 == 0) { d = (d - e) / 2; c += ADD + DUP; }
        else if (d % 2 == 0) { d /= 2; c += ADD + DUP; }
        else { e /= 2; c += ADD + DUP; }
    }
    if (d != 1) return 999999999.;
    return c;
}

/* optional trace of executed chain steps (for planner tests):
 * 'I' init-dup, 'S' swap, '3','4','5','9' conditions, 'F' final add */
static uint8_t *g_trace = NULL; static uint64_t g_trace_len = 0, g_trace_cap = 0;
static void tr(uint8_t c) { if (g_trace && g_trace_len < g_trace_cap) g_trace[g_trace_len] = c; if (g_trace) g_trace_len++; }

/* ecm.c:565-884 */
static int prac(owork *w, opt *P, uint64_t c)
{
    uint64_t d, e, r;
    double cmin, cost;
    int i;

    for (i = d = 0, cmin = ADD * (double)c; d < NV; d++) {
        cost = lucas_cost(c, val[d]);
        if (cost < cmin) { cmin = cost; i = d; }
    }
    d = c;
    r = (uint64_t)((double)d * val[i] + 0.5);
    d = c - r;
    e = 2 * r - c;

    pt_set(&w->pt1, P); pt_set(&w->pt2, P); pt_set(&w->pt3, P);
    fsub(w, w->diff1, w->pt1.X, w->pt1.Z);
    fadd(w, w->sum1, w->pt1.X, w->pt1.Z);
    vec_duplicate(w, w->sum1, w->diff1, &w->pt1);
    tr('I');

    while (d != e) {
        if (d < e) {
            r = d; d = e; e = r;
            pt_swap(&w->pt1, &w->pt2);
            tr('S');
        }
        if ((d + 3) / 4 <= e) {            /* condition 3, ecm.c:683-713 */
            d -= e;
            faddsub(w, w->pt2.X, w->pt2.Z, w->sum1, w->diff1);
            faddsub(w, w->pt1.X, w->pt1.Z, w->sum2, w->diff2);
            vec_add(w, &w->pt3, &w->pt4);
            { opt t = w->pt2; w->pt2 = w->pt4; w->pt4 = w->pt3; w->pt3 = t; }
            tr('3');
        } else if ((d + e) % 2 == 0) {     /* condition 4, ecm.c:714-726 */
            d = (d - e) / 2;
            faddsub(w, w->pt2.X, w->pt2.Z, w->sum1, w->diff1);
            faddsub(w, w->pt1.X, w->pt1.Z, w->sum2, w->diff2);
            vec_add(w, &w->pt3, &w->pt2);
            vec_duplicate(w, w->sum2, w->diff2, &w->pt1);
            tr('4');
        } else if (d % 2 == 0) {           /* condition 5, ecm.c:728-740 */
            d /= 2;
            faddsub(w, w->pt3.X, w->pt3.Z, w->sum1, w->diff1);
            faddsub(w, w->pt1.X, w->pt1.Z, w->sum2, w->diff2);
            vec_add(w, &w->pt2, &w->pt3);
            vec_duplicate(w, w->sum2, w->diff2, &w->pt1);
            tr('5');
        } else {                           /* condition 9, ecm.c:853-865 */
            e /= 2;
            faddsub(w, w->pt3.X, w->pt3.Z, w->sum1, w->diff1);
            faddsub(w, w->pt2.X, w->pt2.Z, w->sum2, w->diff2);
            vec_add(w, &w->pt1, &w->pt3);
            vec_duplicate(w, w->sum2, w->diff2, &w->pt2);
            tr('9');
        }
    }
    fsub(w, w->diff1, w->pt1.X, w->pt1.Z);
    fadd(w, w->sum1, w->pt1.X, w->pt1.Z);
    fsub(w, w->diff2, w->pt2.X, w->pt2.Z);
    fadd(w, w->sum2, w->pt2.X, w->pt2.Z);
    vec_add(w, &w->pt3, P);
    tr('F');
    return d == 1;
}

/* ecm.c:886-976, binary Montgomery ladder */
static void next_pt_vec(owork *w, opt *P, uint64_t c)
{
    uint64_t mask, d, e;
    opt *x1 = &w->pt1, *x2 = &w->pt2;
    if (c == 1) return;
    pt_set(x1, P);
    fsub(w, w->diff1, P->X, P->Z);
    fadd(w, w->sum1, P->X, P->Z);
    vec_duplicate(w, w->sum1, w->diff1, x2);
    if (c == 2) { pt_set(P, x2); return; }
    mask = 1ULL << (64 - __builtin_clzll(c) - 2);
    d = 1; e = 2;
    while (mask > 0) {
        faddsub(w, x2->X, x2->Z, w->sum2, w->diff2);
        faddsub(w, x1->X, x1->Z, w->sum1, w->diff1);
        if (c & mask) {
            vec_add(w, P, x1);
            vec_duplicate(w, w->sum2, w->diff2, x2);
            d = d + e; e *= 2;
        } else {
            vec_add(w, P, x2);
            vec_duplicate(w, w->sum1, w->diff1, x1);
            e = e + d; d *= 2;
        }
        mask >>= 1;
    }
    if (d != c) { fprintf(stderr, "oracle: ladder mismatch\n"); abort(); }
    pt_set(P, x1);
}

/* ------------------------------------------------------------------ */
/* primes: plain odd sieve (the reference uses YAFU's SoE; any exact   */
/* prime list is equivalent)                                           */
/* ------------------------------------------------------------------ */
static uint64_t *sieve_range(uint64_t lo, uint64_t hi, uint64_t *count)
{   /* all primes p with lo <= p <= hi */
    uint64_t span = hi - lo + 1, i, p, n = 0;
    uint8_t *comp = (uint8_t *)calloc(span, 1);
    uint32_t root = 1; while ((uint64_t)root * root <= hi) root++;
    uint8_t *small = (uint8_t *)calloc(root + 1, 1);
    for (p = 2; p <= root; p++) {
        if (small[p]) continue;
        for (i = p * p; i <= root; i += p) small[i] = 1;
        uint64_t start = (lo + p - 1) / p * p; if (start < p * p) start = p * p;
        for (i = start; i <= hi; i += p) comp[i - lo] = 1;
    }
    for (i = 0; i < span; i++) if (!comp[i] && lo + i >= 2) n++;
    uint64_t *out = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    n = 0;
    for (i = 0; i < span; i++) if (!comp[i] && lo + i >= 2) out[n++] = lo + i;
    free(comp); free(small);
    *count = n;
    return out;
}

/* ------------------------------------------------------------------ */
/* curve construction, ecm.c:1548-1803 (true residues: no <<R step)    */
/* ------------------------------------------------------------------ */
static void build_one_curve(owork *w, mpz_t X, mpz_t Z, mpz_t A, uint64_t sigma)
{
    mpz_t u, v, t1, t2, t3, t4;
    mpz_init(u); mpz_init(v); mpz_init(t1); mpz_init(t2); mpz_init(t3); mpz_init(t4);
    mpz_set_ui(v, sigma); mpz_mul_2exp(v, v, 2);
    mpz_set_ui(u, sigma); mpz_mul(u, u, u); mpz_sub_ui(u, u, 5);
    mpz_mul(X, u, u); mpz_mul(X, X, u); mpz_tdiv_r(X, X, w->n);
    mpz_mul(Z, v, v); mpz_mul(Z, Z, v); mpz_tdiv_r(Z, Z, w->n);
    if (mpz_cmp(u, v) > 0) { mpz_sub(t1, v, u); mpz_add(t1, t1, w->n); }
    else mpz_sub(t1, v, u);
    mpz_mul(t2, t1, t1); mpz_tdiv_r(t2, t2, w->n);
    mpz_mul(t4, t2, t1); mpz_tdiv_r(t4, t4, w->n);
    mpz_mul_ui(t1, u, 3); mpz_add(t3, t1, v); mpz_tdiv_r(t3, t3, w->n);
    mpz_mul(t1, t3, t4); mpz_tdiv_r(t1, t1, w->n);
    mpz_mul_ui(t2, X, 16); mpz_mul(t4, t2, v); mpz_tdiv_r(t4, t4, w->n);
    mpz_invert(t2, t4, w->n);
    mpz_mul(A, t1, t2); mpz_tdiv_r(A, A, w->n);
    mpz_invert(t1, Z, w->n);
    mpz_mul(X, X, t1);
    mpz_set_ui(Z, 1);
    mpz_tdiv_r(X, X, w->n); mpz_tdiv_r(A, A, w->n);
    /* tdiv_r truncates toward zero: a negative intermediate (only possible for
     * N < 2^128 with a 64-bit sigma) would stay negative in the
     * reference before its <<R; canonicalise like the Montgomery ops would. */
    if (mpz_sgn(X) < 0) mpz_add(X, X, w->n);
    if (mpz_sgn(A) < 0) mpz_add(A, A, w->n);
    mpz_clear(u); mpz_clear(v); mpz_clear(t1); mpz_clear(t2); mpz_clear(t3); mpz_clear(t4);
}

/* ecm.c:1806-1854; PRIMES = all primes of [0, ...], PRIMES[0] = 2 is skipped */
static void ecm_stage1(owork *w, opt *P, uint64_t b1, const uint64_t *primes, uint64_t nump)
{
    uint64_t q = 2, i;
    while (q < b1) {
        fsub(w, w->diff1, P->X, P->Z);
        fadd(w, w->sum1, P->X, P->Z);
        vec_duplicate(w, w->sum1, w->diff1, P);
        tr('D');
        q *= 2;
    }
    for (i = 1; i < nump && primes[i] < b1; i++) {
        uint64_t c = 1;
        q = primes[i];
        do { prac(w, P, q); c *= q; } while ((c * q) < b1);
    }
}

/* The stage-1 driver loop of vececm(), ecm.c:1134-1140 and 1207-1234: primes come in ranges of PRIME_RANGE = 1e8
 * and ecm_stage1 is called once per range with the global PRIMES array replaced.  Beyond the first range this
 * has two visible consequences that a bit-exact engine has to copy: every call repeats the doublings for
 * q = 2,4,.. < B1 (ecm.c:1815-1822), and every call starts at PRIMES[1], so the first prime of each later range
 * (100000007, 200000033, ...) is never used.  stop_after_ranges > 0 stops early: the state the reference writes
 * to checkpoint.txt after that many ranges (ecm.c:1237-1311).                                                  */
static uint64_t STAGE1_PRIME_RANGE = 100000000ULL;      /* PRIME_RANGE, main.c:585 */
/* test hook: small ranges exercise the range-by-range driver at small B1 (the 1e8 value is pinned by the golden
 * vector syn206_b1_1.1e8_two_stage1_ranges) */
void oracle_set_prime_range(uint64_t r) { STAGE1_PRIME_RANGE = r ? r : 100000000ULL; }
static void stage1_all_ranges(owork *w, opt *P, uint64_t b1, uint64_t b2, int stop_after_ranges)
{
    uint64_t p, rangemin = 0, rangemax, nump, *primes;
    int done = 0;
    rangemax = (b2 + 1000 < STAGE1_PRIME_RANGE) ? b2 + 1000 : STAGE1_PRIME_RANGE;
    for (p = 0; p < b1; p += STAGE1_PRIME_RANGE) {
        if (p >= rangemax) {
            rangemin = rangemax;
            rangemax = (b2 + 1000 < rangemin + STAGE1_PRIME_RANGE) ? b2 + 1000 : rangemin + STAGE1_PRIME_RANGE;
        }
        primes = sieve_range(rangemin, rangemax, &nump);
        ecm_stage1(w, P, b1, primes, nump);
        free(primes);
        if (stop_after_ranges > 0 && ++done == stop_after_ranges) break;
    }
}

/* ------------------------------------------------------------------ */
/* stage 2                                                             */
/* ------------------------------------------------------------------ */
static uint32_t gcd32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

/* main.c:834-882 D table; U=16/L=32 is what the reference binary selects
 * (main.c:912 reads an uninitialised paircost). */
static uint32_t stage2_D(uint64_t b1)
{
    uint32_t D = 2310;
    if (b1 <= 4096) D = 1155;
    if (b1 <= 2048) D = 385;
    if (b1 <= 512) D = 210;
    if (b1 <= 256) D = 120;
    if (b1 <= 128) D = 60;
    if (b1 <= 60) D = 30;
    return D;
}

/* ecm.c:248-340 (map construction ecm.c:301-329) */
static void work_init_stage2(owork *w, uint64_t b1)
{
    uint32_t i, j, m, D, U, L, R;
    D = w->D = stage2_D(b1);
    for (j = 0, i = 0; i < 2 * D; i++) if (gcd32(i, 2 * D) == 1) j++;
    R = w->R = j + 3;
    U = w->U = 16; L = w->L = 2 * U;
    w->npb = U * (R + 1);
    w->Pb = (opt *)malloc(w->npb * sizeof(opt));
    w->Pbprod = (mpz_t *)malloc(w->npb * sizeof(mpz_t));
    for (i = 0; i < w->npb; i++) { pt_init(&w->Pb[i]); mpz_init(w->Pbprod[i]); }
    w->Pa = (opt *)malloc(2 * L * sizeof(opt));
    w->Pa_inv = (mpz_t *)malloc(2 * L * sizeof(mpz_t));
    w->Paprod = (mpz_t *)malloc(2 * L * sizeof(mpz_t));
    for (i = 0; i < 2 * L; i++) { pt_init(&w->Pa[i]); mpz_init(w->Pa_inv[i]); mpz_init(w->Paprod[i]); }
    w->map = (uint32_t *)calloc(U * (D + 1) + 3, sizeof(uint32_t));
    w->map[0] = 0; w->map[1] = 1; w->map[2] = 2;
    m = 3;
    for (i = 0; i < U; i++) {
        j = (i == 0) ? 3 : 1;
        for (; j < D; j++) w->map[i * D + j] = (gcd32(j, D) == 1) ? m++ : 0;
        if (i == 0) w->map[i * D + j] = m++;
    }
}

/* ecm.c:1869-2001 / 2003-2136: Montgomery's simultaneous inversion of
 * Z[start..start+num) ; out[i] = X[i]/Z[i].  Values are true residues, so the
 * reference's leave/re-enter-Montgomery steps vanish.
 *
 * Non-invertible product a (a factor has been met): the reference writes the
 * raw integer g = gcd(a,N) into the Montgomery-domain accumulator and leaves in
 * B[num-1] whatever the vector lane held.  For lane 0 of a vector that is the
 * raw integer a itself: mpz_invert leaves its (freshly initialised, zero)
 * output untouched and insert_mpz_to_vec(0) writes no words (main.c:117-138),
 * so the lane keeps montmul(A,1) = a.  Read as Montgomery-domain values with
 * the reference's R = 2^MAXBITS these are the true residues g/R and a/R.  That
 * is what this oracle (and the GPU engine) reproduce: "lane-0 semantics of the
 * DIGITBITS=52 build".  Lanes 1..7 of the reference instead inherit the previous
 * lane's inverse (ecm.c:1925-1949) -- data that depends on the neighbouring
 * curve and on R, which no R-independent engine can reproduce; for those lanes
 * only the reported factor is compared. */
static int batch_invert(owork *w, opt *pts, mpz_t *out, mpz_t *A, int start, int num, int inplace)
{
    int i, found = 0;
    mpz_t *B = (mpz_t *)malloc(num * sizeof(mpz_t));
    mpz_t g;
    mpz_init(g);
    for (i = 0; i < num; i++) mpz_init(B[i]);
    w->numinv++;
    mpz_set(A[0], pts[start].Z);
    for (i = 1; i < num; i++) fmul(w, A[i], pts[start + i].Z, A[i - 1]);
    if (mpz_invert(B[num - 1], A[num - 1], w->n) == 0) {
        mpz_gcd(g, A[num - 1], w->n);
        fmul(w, w->acc, g, w->rref_inv);
        fmul(w, B[num - 1], A[num - 1], w->rref_inv);
        found = 1;
    } else {
        /* Stale high words (ecm.c:1913-1946 with insert_mpz_to_vec, main.c:117-138): the vector lane that receives the
         * inverse still holds the operand a = A[num-1] taken out of Montgomery form (ecm.c:1903-1911), and
         * insert_mpz_to_vec only writes the words of its source that are non-zero from the top -- the 52-bit words of
         * a above the length of the new value v = a^-1 * 2^MAXBITS mod N stay where they are.  The lane then holds
         * v + (a with its low 52k bits cleared), k = words of v, and the reference computes on with that residue.
         * With b bits in the top word of N this happens to about 2^-b of all inversions: never in practice for most
         * inputs, constantly when N is one bit longer than a multiple of 52 (test.csh lines 2 and 25: 729 and 417 bits). */
        size_t k;
        mpz_mul_2exp(g, B[num - 1], w->maxbits_ref);         /* v (special-form inputs: maxbits_ref = 0, plain values) */
        mpz_mod(g, g, w->n);
        k = (mpz_sizeinbase(g, 2) + 51) / 52;
        if (mpz_sgn(g) == 0) k = 0;
        {
            mpz_t stale; mpz_init(stale);
            mpz_tdiv_q_2exp(stale, A[num - 1], 52 * k);
            if (mpz_sgn(stale) != 0) {
                mpz_mul_2exp(stale, stale, 52 * k);
                mpz_add(g, g, stale);
                fmul(w, B[num - 1], g, w->rref_inv);
            }
            mpz_clear(stale);
        }
    }
    for (i = num - 2; i >= 0; i--) fmul(w, B[i], pts[start + i + 1].Z, B[i + 1]);
    for (i = 0; i < num; i++) {
        mpz_t *dst = inplace ? &pts[start + i].Z : &out[start + i];
        if (i == 0) mpz_set(*dst, B[0]); else fmul(w, *dst, B[i], A[i - 1]);
    }
    for (i = 0; i < num; i++) {
        if (inplace) fmul(w, pts[start + i].X, pts[start + i].X, pts[start + i].Z);
        else fmul(w, out[start + i], pts[start + i].X, out[start + i]);
    }
    for (i = 0; i < num; i++) mpz_clear(B[i]);
    free(B); mpz_clear(g);
    if (found) w->found_inv = 1;
    return found;
}

/* ecm.c:2201-2340 */
static void ecm_stage2_init(owork *w, opt *P, uint64_t b1)
{
    uint32_t wD = w->D, U = w->U, j;
    int lastMapID = 0;
    mpz_t ox, oz; mpz_init(ox); mpz_init(oz);
    w->amin = (uint32_t)((b1 + wD) / (2 * wD));
    w->paired = 0; w->ptadds = 0; w->ptdups = 0; w->numinv = 0;

    pt_set(&w->Pb[1], P);
    pt_set(&w->Pb[2], P);
    faddsub(w, P->X, P->Z, w->sum1, w->diff1);
    vec_duplicate(w, w->sum1, w->diff1, &w->Pb[2]);
    pt_set(&w->pt2, &w->Pb[1]);
    pt_set(&w->pt1, &w->Pb[2]);

    for (j = 3; j <= U * wD; j++) {
        opt *P1 = &w->pt1, *P2 = &w->Pb[1], *P3 = &w->pt2, *Pout = &w->Pb[w->map[j]];
        if (w->map[j] > 0) lastMapID = w->map[j];
        faddsub(w, P1->X, P1->Z, w->sum1, w->diff1);
        faddsub(w, P2->X, P2->Z, w->sum2, w->diff2);
        fmul(w, w->tt1, w->diff1, w->sum2);
        fmul(w, w->tt2, w->sum1, w->diff2);
        faddsub(w, w->tt1, w->tt2, ox, oz);
        fsqr(w, w->tt1, ox);
        fsqr(w, w->tt2, oz);
        fmul(w, ox, w->tt1, P3->Z);
        fmul(w, oz, w->tt2, P3->X);
        mpz_set(Pout->X, ox); mpz_set(Pout->Z, oz);
        w->ptadds++;
        pt_set(P3, P1);
        pt_set(P1, Pout);
    }
    mpz_set_ui(w->acc, 1);           /* = mdata->one, i.e. true residue 1 */
    /* batch_invert_pt_inplace(Pb, Pbprod, ..., lastMapID+1): entries 1..lastMapID */
    batch_invert(w, w->Pb, NULL, w->Pbprod, 1, lastMapID, 1);
    pt_set(&w->Pd, P);
    next_pt_vec(w, &w->Pd, wD);
    mpz_clear(ox); mpz_clear(oz);
}

/* simple FIFO with the capacity/ordering of queue.c:30-101 */
typedef struct { uint32_t *q; uint32_t sz, head, tail, len; } oq_t;
static void oq_push(oq_t *Q, uint32_t e)
{
    Q->q[Q->tail++] = e; Q->len++;
    if (Q->tail == Q->sz) Q->tail = 0;
    if (Q->len >= Q->sz) { fprintf(stderr, "oracle: Q overflowed\n"); abort(); }
}
static uint32_t oq_pop(oq_t *Q)
{
    uint32_t e;
    if (Q->len == 0) { fprintf(stderr, "oracle: dequeue from empty queue\n"); abort(); }
    e = Q->q[Q->head++]; if (Q->head == Q->sz) Q->head = 0; Q->len--;
    return e;
}

/* ecm.c:2559-2910 Montgomery's PAIR; Qmap/Qrmap built as in main.c:715-749 */
static uint32_t pair(uint32_t *pairmap_v, uint32_t *pairmap_u, uint32_t D, uint32_t U,
    const uint64_t *primes, uint64_t nump, uint64_t B1, uint64_t B2, uint32_t *amin_out,
    uint32_t *npairs_out)
{
    int64_t w = D, L = 2 * U, umax = w * U, q, mq;
    uint32_t *Qmap = (uint32_t *)malloc(2 * D * sizeof(uint32_t));
    uint32_t *Qrmap = (uint32_t *)malloc(2 * D * sizeof(uint32_t));
    uint32_t j, k, R, mapid = 0, pairs = 0;
    uint64_t pid = 0, amin = (B1 + w) / (2 * w), a, s, ap, u;
    oq_t *Q;
    int i;

    for (j = 0, k = 0; k < 2 * D; k++) {
        if (gcd32(k, 2 * D) == 1) { Qmap[k] = j; Qrmap[j++] = k; } else Qmap[k] = (uint32_t)-1;
    }
    R = j;
    for (k = j; k < 2 * D; k++) Qrmap[k] = (uint32_t)-1;
    Q = (oq_t *)calloc(R, sizeof(oq_t));
    for (k = 0; k < R; k++) { Q[k].q = (uint32_t *)malloc(D * sizeof(uint32_t)); Q[k].sz = D; }

    while (pid < nump && primes[pid] < B1) pid++;
    while (pid < nump && primes[pid] < B2) {
        s = primes[pid];
        a = (s + w) / (2 * w);
        while (a >= (amin + L)) {
            uint64_t oldmin = amin;
            amin = amin + L - U;
            for (i = 0; i < (int)R; i++) {
                int len = Q[i].len, jj;
                uint32_t uval = (Qrmap[i] > (uint32_t)w) ? (uint32_t)(2 * w - Qrmap[i]) : Qrmap[i];
                for (jj = 0; jj < len; jj++) {
                    ap = oq_pop(&Q[i]);
                    if ((uint32_t)ap < amin) {
                        pairmap_v[mapid] = (uint32_t)(2 * ap - oldmin);
                        pairmap_u[mapid] = uval;
                        mapid++; pairs++;
                    } else oq_push(&Q[i], (uint32_t)ap);
                }
            }
            pairmap_u[mapid] = 0; pairmap_v[mapid] = 0; mapid++;
        }
        q = (int64_t)s - 2 * (int64_t)a * w;
        if (q < 0) mq = -q; else mq = 2 * w - q;
        do {
            if (Q[Qmap[mq]].len > 0) {
                ap = oq_pop(&Q[Qmap[mq]]);
                if (q < 0) u = w * (a - ap) - (uint64_t)(-q); else u = w * (a - ap) + q;
                if (u > (uint64_t)umax) {
                    int64_t qq = (q < 0) ? -q : q;
                    if (q >= 0 && qq >= w) qq = 2 * w - qq;
                    pairmap_v[mapid] = (uint32_t)(2 * ap - amin);
                    pairmap_u[mapid] = (uint32_t)qq;
                    mapid++; pairs++;
                } else {
                    pairmap_v[mapid] = (uint32_t)(a + ap - amin);
                    pairmap_u[mapid] = (uint32_t)u;
                    mapid++; pairs++;
                }
            } else {
                if (q < 0) oq_push(&Q[Qmap[2 * w + q]], (uint32_t)a);
                else oq_push(&Q[Qmap[q]], (uint32_t)a);
                u = 0;
            }
        } while (u > (uint64_t)umax);
        pid++;
    }
    for (i = 0; i < (int)R; i++) {
        int len = Q[i].len, jj;
        uint32_t uval = (Qrmap[i] > (uint32_t)w) ? (uint32_t)(2 * w - Qrmap[i]) : Qrmap[i];
        for (jj = 0; jj < len; jj++) {
            ap = oq_pop(&Q[i]);
            pairmap_v[mapid] = (uint32_t)(2 * ap - amin);
            pairmap_u[mapid] = uval;
            mapid++; pairs++;
        }
    }
    for (k = 0; k < R; k++) free(Q[k].q);
    free(Q); free(Qmap); free(Qrmap);
    if (amin_out) *amin_out = (uint32_t)amin;
    if (npairs_out) *npairs_out = pairs;
    return mapid;
}

/* ecm.c:2342-2540 */
static void ecm_stage2_pair(owork *w, opt *P, uint32_t steps, const uint32_t *pm_v, const uint32_t *pm_u)
{
    uint32_t wD = w->D, U = w->U, L = w->L, amin = w->amin, mapid;
    int i;
    w->A = (uint64_t)amin * wD * 2;
    pt_set(&w->Pa[0], P);
    next_pt_vec(w, &w->Pa[0], w->A);
    pt_set(&w->Pad, P);
    next_pt_vec(w, &w->Pad, w->A - wD);
    fadd(w, w->sum1, w->Pa[0].X, w->Pa[0].Z);
    fadd(w, w->sum2, w->Pd.X, w->Pd.Z);
    fsub(w, w->diff1, w->Pa[0].X, w->Pa[0].Z);
    fsub(w, w->diff2, w->Pd.X, w->Pd.Z);
    vec_add(w, &w->Pad, &w->Pa[1]);
    w->A += wD;
    for (i = 2; i < (int)(2 * L); i++) {
        faddsub(w, w->Pa[i - 1].X, w->Pa[i - 1].Z, w->sum1, w->diff1);
        faddsub(w, w->Pd.X, w->Pd.Z, w->sum2, w->diff2);
        vec_add(w, &w->Pa[i - 2], &w->Pa[i]);
        w->A += wD;
    }
    batch_invert(w, w->Pa, w->Pa_inv, w->Paprod, 0, 2 * L, 0);
    w->numinv++;                                   /* ecm.c:2429 counts it twice */

    for (mapid = 0; mapid < steps; mapid++) {
        if (pm_u[mapid] == 0 && pm_v[mapid] == 0) {
            for (i = 0; i < (int)(2 * L - 2 * U); i++) {
                pt_set(&w->Pa[i], &w->Pa[i + 2 * U]);
                mpz_set(w->Pa_inv[i], w->Pa_inv[i + 2 * U]);
            }
            for (i = 2 * L - 2 * U; i < (int)(2 * L); i++) {
                faddsub(w, w->Pa[i - 1].X, w->Pa[i - 1].Z, w->sum1, w->diff1);
                faddsub(w, w->Pd.X, w->Pd.Z, w->sum2, w->diff2);
                vec_add(w, &w->Pa[i - 2], &w->Pa[i]);
                w->A += wD;
            }
            amin += U;
            batch_invert(w, w->Pa, w->Pa_inv, w->Paprod, 2 * L - 2 * U, 2 * U, 0);
        } else {
            int pa = (int)(pm_v[mapid] - amin), pb = (int)pm_u[mapid];
            if (pa >= (int)(2 * L) || pa < 0) { fprintf(stderr, "oracle: invalid A offset\n"); abort(); }
            fsub(w, w->tt1, w->Pa_inv[pa], w->Pb[w->map[pb]].X);     /* CROSS_PRODUCT_INV */
            fmul(w, w->acc, w->acc, w->tt1);
            w->paired++;
        }
    }
    w->amin = amin;
}

/* ------------------------------------------------------------------ */
/* work alloc                                                          */
/* ------------------------------------------------------------------ */
/* m_hex == NULL: generic Montgomery path, arithmetic and checks mod N.
 * m_hex != NULL: special-form path (main.c:405-457, 597-616): all arithmetic mod the base number
 * M = 2^k-1 / 2^k+1 / 2^k-c given in m_hex, residues are plain (no Montgomery form, one = 1), the
 * reference's lazily reduced representatives lie in [0,2^k) and equal the canonical residue mod M
 * except on the zero class (probability ~c/2^k per operation; vecarith52.c:880-1025, 4613-4801);
 * gcd checks use the original input N (ecm.c:1108-1119).                                        */
static owork *work_new2(const char *n_hex, const char *m_hex)
{
    owork *w = (owork *)calloc(1, sizeof(owork));
    mpz_init(w->n); mpz_init(w->s);
    mpz_init(w->sum1); mpz_init(w->diff1); mpz_init(w->sum2); mpz_init(w->diff2);
    mpz_init(w->tt1); mpz_init(w->tt2); mpz_init(w->tt3); mpz_init(w->tt4);
    pt_init(&w->pt1); pt_init(&w->pt2); pt_init(&w->pt3); pt_init(&w->pt4);
    pt_init(&w->Pad); pt_init(&w->Pd); mpz_init(w->acc);
    mpz_init(w->nchk);
    if (mpz_set_str(w->nchk, n_hex, 16) != 0) { free(w); return NULL; }
    if (m_hex) {
        if (mpz_set_str(w->n, m_hex, 16) != 0) { free(w); return NULL; }
        mpz_init(w->rref_inv); mpz_set_ui(w->rref_inv, 1);      /* failure lanes keep plain g and a (ecm.c:1903-1946) */
        w->maxbits_ref = 0;
    } else {   /* main.c:465-483: MAXBITS = smallest multiple of 208 strictly above bitlen(N) */
        unsigned long maxbits = 208;
        mpz_set(w->n, w->nchk);
        mpz_init(w->rref_inv);
        while (maxbits <= mpz_sizeinbase(w->n, 2)) maxbits += 208;
        w->maxbits_ref = maxbits;
        mpz_set_ui(w->rref_inv, 1); mpz_mul_2exp(w->rref_inv, w->rref_inv, maxbits);
        if (mpz_invert(w->rref_inv, w->rref_inv, w->n) == 0) mpz_set_ui(w->rref_inv, 0);
    }
    return w;
}
static owork *work_new(const char *n_hex) { return work_new2(n_hex, NULL); }
static void work_free(owork *w)
{
    uint32_t i;
    if (w->Pb) { for (i = 0; i < w->npb; i++) { pt_clear(&w->Pb[i]); mpz_clear(w->Pbprod[i]); } free(w->Pb); free(w->Pbprod); }
    if (w->Pa) { for (i = 0; i < 2 * w->L; i++) { pt_clear(&w->Pa[i]); mpz_clear(w->Pa_inv[i]); mpz_clear(w->Paprod[i]); } free(w->Pa); free(w->Pa_inv); free(w->Paprod); }
    free(w->map);
    mpz_clear(w->n); mpz_clear(w->s);
    mpz_clear(w->sum1); mpz_clear(w->diff1); mpz_clear(w->sum2); mpz_clear(w->diff2);
    mpz_clear(w->tt1); mpz_clear(w->tt2); mpz_clear(w->tt3); mpz_clear(w->tt4);
    pt_clear(&w->pt1); pt_clear(&w->pt2); pt_clear(&w->pt3); pt_clear(&w->pt4);
    pt_clear(&w->Pad); pt_clear(&w->Pd); mpz_clear(w->acc); mpz_clear(w->rref_inv); mpz_clear(w->nchk);
    free(w);
}

/* check_factor, ecm.c:2542-2557 */
static int check_factor(const mpz_t Z, const mpz_t n, mpz_t f)
{
    mpz_gcd(f, Z, n);
    if (mpz_cmp_ui(f, 1) > 0) {
        if (mpz_cmp(f, n) == 0) { mpz_set_ui(f, 0); return 0; }
        return 1;
    }
    return 0;
}

/* ================================================================== */
/* exported API (ctypes)                                               */
/* ================================================================== */
#define PRIME_RANGE 100000000ULL

/* Suyama curve for sigma: X (Z=1) and s=(A+2)/4 as hex true residues. */
int oracle_build_curve(const char *n_hex, uint64_t sigma, char *x_hex, char *s_hex)
{
    owork *w = work_new(n_hex); mpz_t X, Z, A;
    if (!w) return -1;
    mpz_init(X); mpz_init(Z); mpz_init(A);
    build_one_curve(w, X, Z, A, sigma);
    mpz_get_str(x_hex, 16, X); mpz_get_str(s_hex, 16, A);
    mpz_clear(X); mpz_clear(Z); mpz_clear(A); work_free(w);
    return 0;
}

/* Full run of one curve.  Outputs (caller buffers, >= hexlen(N)+2 bytes):
 *  x_hex,z_hex : stage-1 residues as written to save_b1.txt (ecm.c:1327-1380)
 *  f1_dec      : stage-1 factor ("0" if none)    (ecm.c:1335-1342)
 *  acc_hex     : stage-2 accumulator, true residue ("" if stage 2 off)
 *  f2_dec      : stage-2 factor ("0" if none)    (ecm.c:1485-1497)
 *  counters[8] : s1 ptadds, s1 ptdups, s2 ptadds, s2 numinv, s2 paired, pairmap steps, found_inv, last amin
 * do_stage2 follows main.c:543-552 (b2 <= b1 disables stage 2).               */
static int g_stop_after_ranges = 0;       /* oracle_set_checkpoint(): stop stage 1 after that many prime ranges */
void oracle_set_checkpoint(int ranges) { g_stop_after_ranges = ranges; }

static int ecm_curve(const char *n_hex, const char *m_hex, uint64_t b1, uint64_t b2, uint64_t sigma,
    char *x_hex, char *z_hex, char *f1_dec, char *acc_hex, char *f2_dec, uint32_t *counters)
{
    owork *w = work_new2(n_hex, m_hex);
    opt P; mpz_t A, f;
    uint64_t nump, *primes, p;
    int do2 = b2 > b1;
    if (!w) return -1;
    pt_init(&P); mpz_init(A); mpz_init(f);
    memset(counters, 0, 8 * sizeof(uint32_t));
    build_one_curve(w, P.X, P.Z, A, sigma);
    mpz_set(w->s, A);

    stage1_all_ranges(w, &P, b1, do2 ? b2 : b1, g_stop_after_ranges);
    counters[0] = w->ptadds; counters[1] = w->ptdups;
    mpz_get_str(x_hex, 16, P.X); mpz_get_str(z_hex, 16, P.Z);
    if (!check_factor(P.Z, w->nchk, f)) mpz_set_ui(f, 0);
    mpz_get_str(f1_dec, 10, f);
    acc_hex[0] = 0; strcpy(f2_dec, "0");

    if (do2) {
        uint32_t *pm_v, *pm_u;
        work_init_stage2(w, b1);
        ecm_stage2_init(w, &P, b1);
        for (p = b1; p < b2; p += PRIME_RANGE) {       /* ecm.c:1424-1476 */
            uint64_t hi = (p + PRIME_RANGE < b2) ? p + PRIME_RANGE : b2;
            uint32_t steps, amin, npairs;
            primes = sieve_range(p, hi + 1000, &nump);
            pm_v = (uint32_t *)malloc((nump + 2 * (hi - p) / w->D + 1024) * sizeof(uint32_t));
            pm_u = (uint32_t *)malloc((nump + 2 * (hi - p) / w->D + 1024) * sizeof(uint32_t));
            steps = pair(pm_v, pm_u, w->D, w->U, primes, nump, p, hi, &amin, &npairs);
            w->amin = (uint32_t)((p + w->D) / (2 * w->D));        /* pair() resets work->amin */
            ecm_stage2_pair(w, &P, steps, pm_v, pm_u);
            counters[5] += steps;
            free(pm_v); free(pm_u); free(primes);
        }
        counters[2] = w->ptadds; counters[3] = w->numinv; counters[4] = w->paired;
        counters[6] = w->found_inv; counters[7] = w->amin;
        mpz_get_str(acc_hex, 16, w->acc);
        if (!check_factor(w->acc, w->nchk, f)) mpz_set_ui(f, 0);
        mpz_get_str(f2_dec, 10, f);
    }
    pt_clear(&P); mpz_clear(A); mpz_clear(f); work_free(w);
    return 0;
}

int oracle_ecm_curve(const char *n_hex, uint64_t b1, uint64_t b2, uint64_t sigma,
    char *x_hex, char *z_hex, char *f1_dec, char *acc_hex, char *f2_dec, uint32_t *counters)
{ return ecm_curve(n_hex, NULL, b1, b2, sigma, x_hex, z_hex, f1_dec, acc_hex, f2_dec, counters); }

/* Special-form inputs: n_hex = the input cofactor after algebraic-factor removal (what the save file
 * prints as N), m_hex = the base number all arithmetic is done against.  Buffers must hold hexlen(M)+2. */
int oracle_ecm_curve_special(const char *n_hex, const char *m_hex, uint64_t b1, uint64_t b2, uint64_t sigma,
    char *x_hex, char *z_hex, char *f1_dec, char *acc_hex, char *f2_dec, uint32_t *counters)
{ return ecm_curve(n_hex, m_hex, b1, b2, sigma, x_hex, z_hex, f1_dec, acc_hex, f2_dec, counters); }

int oracle_build_curve_special(const char *m_hex, uint64_t sigma, char *x_hex, char *s_hex)
{ return oracle_build_curve(m_hex, sigma, x_hex, s_hex); }

/* The exact save_b1.txt line (ecm.c:1372-1380). */
int oracle_save_line(const char *n_hex, uint64_t b1, uint64_t sigma, const char *x_hex,
    const char *z_hex, char *out, size_t cap)
{
    mpz_t n; mpz_init(n); mpz_set_str(n, n_hex, 16);
    int r = gmp_snprintf(out, cap, "METHOD=ECM; SIGMA=%" PRIu64 "; B1=%" PRIu64 "; N=0x%Zx; X=0x%s; Z=0x%s; PROGRAM=AVX-ECM;\n",
        sigma, b1, n, x_hex, z_hex);
    mpz_clear(n);
    return r;
}

/* PAIR for one prime range [lo,hi) -> pairmap arrays. returns steps. */
uint32_t oracle_pair(uint64_t lo, uint64_t hi, uint32_t D, uint32_t U, uint32_t *pm_v, uint32_t *pm_u,
    uint32_t cap, uint32_t *amin_out, uint32_t *npairs_out)
{
    uint64_t nump, *primes = sieve_range(lo, hi + 1000, &nump);
    uint32_t *v = (uint32_t *)malloc((nump + 2 * (hi - lo) / D + 1024) * sizeof(uint32_t));
    uint32_t *u = (uint32_t *)malloc((nump + 2 * (hi - lo) / D + 1024) * sizeof(uint32_t));
    uint32_t steps = pair(v, u, D, U, primes, nump, lo, hi, amin_out, npairs_out);
    uint32_t ncopy = steps < cap ? steps : cap;
    memcpy(pm_v, v, ncopy * sizeof(uint32_t)); memcpy(pm_u, u, ncopy * sizeof(uint32_t));
    free(v); free(u); free(primes);
    return steps;
}

uint32_t oracle_stage2_D(uint64_t b1) { return stage2_D(b1); }

/* stage-2 index map (ecm.c:301-329); returns number of entries written */
uint32_t oracle_stage2_map(uint64_t b1, uint32_t *map, uint32_t cap)
{
    owork *w = work_new("3"); uint32_t n, i;
    work_init_stage2(w, b1);
    n = w->U * (w->D + 1) + 3;
    for (i = 0; i < n && i < cap; i++) map[i] = w->map[i];
    work_free(w);
    return n;
}

/* Trace of the stage-1 op sequence ('D' = power-of-two doubling, then per prime
 * 'I' [S]{3,4,5,9}* 'F'); returns the full length even if > cap. */
uint64_t oracle_stage1_trace(uint64_t b1, uint8_t *ops, uint64_t cap)
{
    owork *w = work_new("fffffffb"); opt P; mpz_t A; uint64_t len;
    pt_init(&P); mpz_init(A);
    build_one_curve(w, P.X, P.Z, A, 11); mpz_set(w->s, A);
    g_trace = ops ? ops : (uint8_t *)&len; g_trace_cap = ops ? cap : 0; g_trace_len = 0;
    stage1_all_ranges(w, &P, b1, b1, 0);
    len = g_trace_len; g_trace = NULL;
    pt_clear(&P); mpz_clear(A); work_free(w);
    return len;
}

/* PRAC multiplier choice for one prime (lucas_cost arg-min, ecm.c:574-584) */
int oracle_prac_best(uint64_t c)
{
    uint64_t d; int i; double cmin, cost;
    for (i = d = 0, cmin = ADD * (double)c; d < NV; d++) {
        cost = lucas_cost(c, val[d]);
        if (cost < cmin) { cmin = cost; i = (int)d; }
    }
    return i;
}

#ifdef ORACLE_MAIN
/* ecm_oracle N_dec curves B1 B2 sigma : prints save_b1 lines + factor lines */
int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: ecm_oracle N curves B1 B2 sigma\n"); return 1; }
    mpz_t n; mpz_init(n); mpz_set_str(n, argv[1], 10);
    char *nh = mpz_get_str(NULL, 16, n);
    size_t L = strlen(nh) + 8;
    uint64_t curves = strtoull(argv[2], 0, 10), b1 = strtoull(argv[3], 0, 10), b2 = strtoull(argv[4], 0, 10),
        sigma = strtoull(argv[5], 0, 10), i;
    char *x = malloc(L), *z = malloc(L), *f1 = malloc(L * 2), *acc = malloc(L), *f2 = malloc(L * 2), *line = malloc(4 * L + 256);
    uint32_t cnt[8];
    for (i = 0; i < curves; i++) {
        oracle_ecm_curve(nh, b1, b2, sigma + i, x, z, f1, acc, f2, cnt);
        oracle_save_line(nh, b1, sigma + i, x, z, line, 4 * L + 256);
        fputs(line, stdout);
        if (strcmp(f1, "0")) fprintf(stderr, "found factor %s in stage 1 sigma %" PRIu64 "\n", f1, sigma + i);
        if (strcmp(f2, "0")) fprintf(stderr, "found factor %s in stage 2 sigma %" PRIu64 "\n", f2, sigma + i);
    }
    fprintf(stderr, "s1 adds %u dups %u; s2 adds %u inv %u pairs %u steps %u\n", cnt[0], cnt[1], cnt[2], cnt[3], cnt[4], cnt[5]);
    return 0;
}
#endif
