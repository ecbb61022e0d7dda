/*
 * ecm_b200.h -- C ABI of the B200 ECM engine (libecm_b200.so).
 *
 * This is the drop-in boundary for avx-ecm's hot path.  The reference reaches its vector
 * arithmetic through four thread-pool work functions, one call per phase per batch
 * (ecm.c:1130-1133, 167-246): build curves, stage 1, stage-2 init, stage-2 pair.  A GPU cannot
 * be fed per field op (avx_ecm.h:205-209), so the boundary sits at that work-function level:
 * each entry point below replaces one of those phases for a whole batch of curves that all
 * share N.  Plain pointers and sizes only; all big numbers are little-endian arrays of
 * 32-bit limbs.  Batch buffers are limb-major with the curve index fastest,
 *        buf[limb * count + curve]
 * i.e. the layout of the reference's bignum.data[lane + word*VECLEN] (main.c:63-89,117-138).
 *
 * Return value: 0 on success, negative on error (ecm_b200_last_error() has the text).
 * There is no CPU fallback: every call fails with ECM_B200_ENODEV when no CUDA device is
 * usable.
 */
#ifndef ECM_B200_H
#define ECM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECM_B200_OK 0
#define ECM_B200_EINVAL -1
#define ECM_B200_ENODEV -2
#define ECM_B200_ECUDA -3
#define ECM_B200_ESTATE -4
#define ECM_B200_ENOMEM -5

typedef struct ecm_b200_ctx ecm_b200_ctx;

/* ---- context (replaces thread_init/monty_alloc, main.c:597-640, 834-970) -------------------
 * n: modulus, nlimbs 32-bit limbs, little endian, must be odd and > 1.  max_curves: capacity of
 * the batch.  The Montgomery constants (R = 2^(32*k), one = R mod N, -N^-1 mod 2^32) are derived
 * here; residues never depend on R.                                         */
int ecm_b200_create(ecm_b200_ctx **out, int device, const uint32_t *n, int nlimbs, uint32_t max_curves);
/* Special-form inputs (replaces the isMersenne set-up, main.c:405-457 and 597-616): n divides
 * base = 2^k-1, 2^k+1 or 2^k-c.  Like the reference, all curve arithmetic is then done modulo
 * `base` (ecm_b200_limbs() >= baselimbs), residues are reported modulo `base` (ecm.c:1327-1380
 * prints them that way), the factor checks of read_stage1/read_stage2 use gcd(., n)
 * (ecm.c:1108-1119) and a failed stage-2 inversion leaves plain values (ecm.c:1903-1946).
 * Detecting the form and removing algebraic factors is the caller's job (main.c:405-457).    */
int ecm_b200_create_special(ecm_b200_ctx **out, int device, const uint32_t *base, int baselimbs,
                            const uint32_t *n, int nlimbs, uint32_t max_curves);
/* 1 when the context computes with the shift-and-fold kernels (base of the shape 2^k-c, c < 2^31, or 2^k+1,
 * 64 <= k <= 1023, and kernels compiled for its length), 0 when it uses Montgomery products.  Results are
 * identical either way; the environment variable ECM_B200_NO_FOLD forces the Montgomery kernels.        */
int ecm_b200_uses_fold(const ecm_b200_ctx *ctx);
void ecm_b200_destroy(ecm_b200_ctx *ctx);
const char *ecm_b200_last_error(void);
/* number of 32-bit limbs the engine computes with (>= nlimbs of N; kernels exist for a fixed set) */
int ecm_b200_limbs(const ecm_b200_ctx *ctx);

/* ---- phase 0: curve construction (replaces build_one_curve, ecm.c:1548-1803, and
 * ecm_build_curve_work_fcn, ecm.c:201-246).  GMP-ECM "param 0" Suyama curves, curve i uses
 * sigma[i].  Computed on the device (one modular inversion per curve).                        */
int ecm_b200_build_curves(ecm_b200_ctx *ctx, uint32_t count, const uint64_t *sigma);
/* Same phase with host-built curves: plain residues X = x/z, s = (A+2)/4, Z is set to 1
 * (what insert_mpz_to_vec receives in ecm.c:222-228, minus the Montgomery shift).             */
int ecm_b200_load_curves(ecm_b200_ctx *ctx, uint32_t count, const uint32_t *x, const uint32_t *s);

/* ---- phase 1: stage 1 (replaces ecm_stage1, ecm.c:1806-1854) --------------------------------
 * Multiplies every curve's point by 2^e * prod p^k for p^k < B1.  The PRAC chains are planned on
 * the host (ecm_b200_plan_stage1) and cached in the context.                                   */
int ecm_b200_stage1(ecm_b200_ctx *ctx, uint64_t b1);
/* B1 above 1e8: the reference runs ecm_stage1 once per range of 1e8 primes-by-value and appends the
 * intermediate points to checkpoint.txt after every range but the last (vececm, ecm.c:1207-1311).  Every
 * such call repeats the doublings for the powers of two and starts at the second prime of its range
 * (ecm.c:1815-1824); ecm_b200_stage1 reproduces exactly that in one go (2 <= B1 <= 2e9).  To write the
 * checkpoints, run the ranges one at a time, in order, on freshly built curves: after range r (not the
 * last) ecm_b200_read_stage1 returns what the reference saves, with *last_prime the B1 it prints there.  */
int ecm_b200_stage1_ranges(uint64_t b1, uint32_t *count);
int ecm_b200_stage1_range(ecm_b200_ctx *ctx, uint64_t b1, uint32_t range, uint64_t *last_prime);
/* Asynchronous form: enqueue at most max_launches kernel launches of the stage-1 schedule and
 * return; *done is set to 1 when the whole stage has been enqueued.  Used for time slicing.   */
int ecm_b200_stage1_begin(ecm_b200_ctx *ctx, uint64_t b1);
int ecm_b200_stage1_step(ecm_b200_ctx *ctx, uint32_t max_launches, int *done);
int ecm_b200_stage1_launches(const ecm_b200_ctx *ctx, uint32_t *total, uint32_t *issued);
/* fraction of the stage-1 work (ops x curves) enqueued so far, in [0,1] */
int ecm_b200_stage1_progress(const ecm_b200_ctx *ctx, double *fraction);
int ecm_b200_sync(ecm_b200_ctx *ctx);
/* CUDA-event stopwatch on the context's stream: what = 0 record start, 1 record stop, 2 wait for
 * stop and return the elapsed device time in *ms.                                             */
int ecm_b200_timer(ecm_b200_ctx *ctx, int what, float *ms);
/* enqueue a 256 MiB memset on the context's stream (evicts L2 between timed steps) */
int ecm_b200_flush_l2(ecm_b200_ctx *ctx);

/* Results of stage 1 as the reference writes them to save_b1.txt (ecm.c:1327-1380): X and Z out
 * of Montgomery form, limb-major [limb*count+curve] with ecm_b200_limbs() limbs; factor_flag[i]
 * is 1 when gcd(Z,N) is a proper factor (check_factor, ecm.c:2542-2557) and then gcd_out holds it
 * (same layout).  Any output pointer may be NULL.                                              */
int ecm_b200_read_stage1(ecm_b200_ctx *ctx, uint32_t *x, uint32_t *z, uint8_t *factor_flag, uint32_t *gcd_out);

/* ---- phases 2+3: stage 2 (replaces ecm_stage2_init ecm.c:2201-2340 and ecm_stage2_pair
 * ecm.c:2342-2540 driven by pair() ecm.c:2559-2910).  Runs the pairing continuation over all
 * primes in [b1, b2) in the reference's 1e8 ranges, D and U as the reference selects them.     */
int ecm_b200_stage2(ecm_b200_ctx *ctx, uint64_t b1, uint64_t b2);
/* The same continuation at the reference's own granularity (ecm.c:67-72), for a host that keeps avx-ecm's driver loop
 * (ecm.c:1400-1476) and its pair():
 *   ecm_b200_stage2_init   replaces  int foundDuringInv = ecm_stage2_init(P, mdata, work, verbose)   (ecm.c:2201-2340)
 *   ecm_b200_stage2_range  replaces  ecm_stage2_pair(steps, pm_v, pm_u, P, mdata, work, verbose) with work->amin = amin
 *                                    (ecm.c:2342-2540)
 * init builds the baby-step table for the whole batch, which must fit the device in one wave (ECM_B200_ENOMEM
 * otherwise: ecm_b200_stage2 runs such batches in waves); the tables stay resident until the curves are rebuilt.
 * pm_v/pm_u are pair()'s pairmap_v/pairmap_u (ecm.c:2559-2910; ecm_b200_pair produces the same arrays): entry
 * (0,0) slides the giant-step window by U, any other entry multiplies the accumulator by Pa[pm_v-amin] - Pb[map[pm_u]].
 * A pairmap that would leave the tables is rejected with ECM_B200_EINVAL.  ecm_b200_read_stage2 may be called after
 * init and after every range.                                                                                        */
int ecm_b200_stage2_init(ecm_b200_ctx *ctx, uint64_t b1, int *found_inv);
int ecm_b200_stage2_range(ecm_b200_ctx *ctx, uint32_t amin, const uint32_t *pm_v, const uint32_t *pm_u, uint32_t steps);
/* acc: stage-2 accumulator out of Montgomery form; factor_flag/gcd_out as above for gcd(acc,N)
 * (ecm.c:1485-1497); inv_fail[i] = 1 when curve i met a non-invertible element.                */
int ecm_b200_read_stage2(ecm_b200_ctx *ctx, uint32_t *acc, uint8_t *factor_flag, uint32_t *gcd_out, uint8_t *inv_fail);

/* counters of the compiled stage-2 program, the numbers the reference prints (ecm.c:1481-1483) */
int ecm_b200_stage2_counters(const ecm_b200_ctx *ctx, uint64_t *ptadds, uint64_t *numinv, uint64_t *paired, uint64_t *steps);

/* ---- host-side planners (the scalar control flow the reference also runs on the host) ------
 * Stage-1 op stream for B1 (prac(), lucas_cost(), ecm.c:479-884, driven as in ecm.c:1815-1832).
 * Returns the number of stream bytes (also when ops==NULL or cap too small).  counts[0..1] =
 * point additions / doublings, the numbers the reference prints (ecm.c:1849).                  */
uint64_t ecm_b200_plan_stage1(uint64_t b1, uint8_t *ops, uint64_t cap, uint64_t *counts);
/* Montgomery's PAIR for the primes of [lo,hi) (pair(), ecm.c:2559-2910): fills pairmap_v/u,
 * returns the number of steps (also when the arrays are NULL / cap too small).                */
uint32_t ecm_b200_pair(uint64_t lo, uint64_t hi, uint32_t D, uint32_t U, uint32_t *pairmap_v,
                       uint32_t *pairmap_u, uint32_t cap, uint32_t *amin_final, uint32_t *npairs);
/* Compile the stage-2 program for (B1,B2) without running it; returns its length in instructions.
 * counts[0..5] = point additions, inversions, pair products, pairmap steps (ecm.c:1481-1483),
 * final amin, table entries per curve.                                                         */
uint64_t ecm_b200_plan_stage2(uint64_t b1, uint64_t b2, uint64_t *counts);
/* The compiled stage-2 instruction stream itself (for inspection and for CPU-side checking of the
 * compiler): which = -1 the ecm_stage2_init program, which = r >= 0 the program of the r-th 1e8 prime
 * range.  Instruction word: lo = op | d<<8 | x<<16 | y<<24 (V_MUL2: 4-bit fields d,x,y,e,u,v from bit 8),
 * hi = imm; op codes and slot numbers are listed in avx-ecm_b200/csrc/plan2.hpp.  layout[0..12] = npb,
 * table bases pbx, pbz, pba, pax, paz, pai, paa, qx, qz, pdx, pdz, total entries.  Returns the length. */
uint64_t ecm_b200_stage2_program(uint64_t b1, uint64_t b2, int which, uint64_t *out, uint64_t cap, uint32_t *layout);
/* The program ecm_b200_stage2_range would run for (amin, pairmap) with the geometry of b1; returns its length,
 * 0 when the pairmap is rejected (it would leave the tables).  For CPU-side checking, no GPU needed.      */
uint64_t ecm_b200_stage2_pairmap_program(uint64_t b1, uint32_t amin, const uint32_t *pm_v, const uint32_t *pm_u, uint32_t steps,
                                         uint64_t *out, uint64_t cap);
/* Phase program of one stage-1 macro-op (0 DBL, 1 INIT, 2..5 PRAC rules 3/4/5/9, 6 FINAL) of the register-resident
 * stage-1 kernel: phase word = kind | x<<4 | y<<8 | z<<12 | flag<<16 with kind 0 A1, 1 A2, 2 A3, 3 D1L, 4 D1P, 5 D2,
 * 6 D3, 7 COPY (avx-ecm_b200/csrc/rv_prog.hpp, rv.cuh).  Returns the number of phases.  For CPU-side checking.     */
int ecm_b200_rv_program(int macro_op, uint32_t *phases, int cap);
/* Stage-2 geometry chosen for B1 (thread_init, main.c:834-970): D, U, L, R.                    */
void ecm_b200_stage2_params(uint64_t b1, uint32_t *D, uint32_t *U, uint32_t *L, uint32_t *R);

/* ---- field-op hook (the reference's operator table vecmulmod_ptr ... avx_ecm.h:205-209) -----
 * op: 0 mul, 1 sqr, 2 add, 3 sub on count independent operand pairs, plain residues in/out
 * (the engine enters and leaves Montgomery form around the op).  For tests and micro-benchmarks. */
int ecm_b200_fieldop(ecm_b200_ctx *ctx, int op, uint32_t count, const uint32_t *a, const uint32_t *b, uint32_t *r, int repeat);

/* ---- instrumentation -------------------------------------------------------------------------
 * Kernel launches issued by this library in this process, and device time (ms) spent in them
 * as measured with CUDA events on the context's stream by the last stage call.                 */
uint64_t ecm_b200_launch_count(void);
int ecm_b200_last_timing(const ecm_b200_ctx *ctx, float *stage_ms, uint32_t *launches);
/* Live integer-multiply peak of this GPU: 32x32->64 products per second retired by dependent-free
 * IMAD.WIDE.U32 chains (the instruction the multiply is built from).                           */
int ecm_b200_measure_imad_peak(int device, double *products_per_sec, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif
