// avx-ecm-b200 -- command-line driver with avx-ecm's argument contract
//     avx-ecm-b200 $input $numcurves $B1 [$gpus] [$B2] [$sigma]
// (main.c:380-384; the 4th argument, the reference's thread count, is the number of GPUs here: one
// host thread per GPU, each with its own engine context and a disjoint sigma range).  Writes the
// reference's output files in its formats: save_b1.txt (GMP-ECM resume lines, ecm.c:1372-1380)
// and ecm_results.txt (ecm.c:1362-1366, 1517-1520).  All arithmetic on the hot path runs on the
// GPUs through include/ecm_b200.h; GMP is used for the expression, file output and PRP labels.
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/ecm_b200.h"
#include "calc.hpp"
#include "special.hpp"

static double now()
{
    struct timeval t; gettimeofday(&t, NULL);
    return t.tv_sec + t.tv_usec * 1e-6;
}

// Knuth's MMIX LCG, the generator the reference draws random sigmas from (main.c:993-998)
static uint64_t lcg_next(uint64_t *s) { *s = 6364136223846793005ULL * *s + 1442695040888963407ULL; return *s; }

static void limbs_to_mpz(mpz_t out, const std::vector<uint32_t> &buf, int L, uint32_t count, uint32_t i)
{
    std::vector<uint32_t> t(L);
    for (int k = 0; k < L; k++) t[k] = buf[(size_t)k * count + i];
    mpz_import(out, L, -1, 4, 0, 0, t.data());
}

static void report(FILE *res, mpz_t f, int stage, uint64_t bound, uint32_t curve, uint64_t sigma)
{
    char ftype[32];
    snprintf(ftype, sizeof ftype, "%s%d", mpz_probab_prime_p(f, 3) ? "PRP" : "C", (int)mpz_sizeinbase(f, 10));
    gmp_printf("\nfound %s factor %Zd in stage %d (B%d = %" PRIu64 "): thread %d, vec %d, sigma %" PRIu64 "\n",
               ftype, f, stage, stage, bound, 0, (int)(curve % 8), sigma);
    if (res)
        gmp_fprintf(res, "\nfound %s factor %Zd in stage %d (B%d = %" PRIu64 "): curve %d, thread %d, vec %d, sigma %" PRIu64 "\n",
                    ftype, f, stage, stage, bound, (int)curve, 0, (int)(curve % 8), sigma);
    fflush(stdout);
}

// Output files are appended to as soon as the data exists, like the reference does: checkpoint.txt after every
// stage-1 prime range but the last (ecm.c:1237-1311), save_b1.txt right after stage 1 and BEFORE stage 2 starts
// (ecm.c:1319-1388), so that a failure later on (stage 2 running out of memory on one GPU, a crash hours into a large
// B1) never costs finished stage-1 work.  The GPUs finish at different times; a ticket per file section keeps the
// lines in sigma order (= the reference's batch/thread/lane order with threads=1): shard g writes section k only
// after shards 0..g-1 have, and a shard that failed passes its turn without writing.
struct Output {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> turn;                 // per section: the shard whose turn it is
    mpz_t N;
    FILE *res = nullptr;                   // ecm_results.txt
    int found = 0;
    template <class F> void in_turn(size_t section, int shard, F &&write)
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return turn[section] == shard; });
        write();
        turn[section]++;
        cv.notify_all();
    }
};

struct Shard {
    int gpu = 0;
    uint32_t first = 0, count = 0;
    std::vector<uint64_t> sigma;
    int limbs = 0;
    double t_build = 0, t_s1 = 0, t_s2 = 0;
    std::string error;
};

// one section of resume lines for this shard: checkpoint.txt (B1 = last prime used) or save_b1.txt
static void write_lines(Output *out, const Shard *s, const char *file, uint64_t b1_printed, bool announce,
                        const std::vector<uint32_t> &X, const std::vector<uint32_t> &Z, const std::vector<uint8_t> &fl,
                        const std::vector<uint32_t> &G)
{
    FILE *fp = fopen(file, "a");
    if (!fp) printf("could not open %s for appending, Stage 1 data will not be saved\n", file);
    if (announce && s->gpu == 0) printf("Saving checkpoint after p=%" PRIu64 "\n", b1_printed);
    mpz_t x, z, f; mpz_init(x); mpz_init(z); mpz_init(f);
    for (uint32_t i = 0; i < s->count; i++) {
        limbs_to_mpz(x, X, s->limbs, s->count, i);
        limbs_to_mpz(z, Z, s->limbs, s->count, i);
        if (!announce && mpz_sgn(z) == 0) printf("something failed: curve %u has zero result\n", s->first + i);
        if (fl[i]) { limbs_to_mpz(f, G, s->limbs, s->count, i); report(out->res, f, 1, b1_printed, s->first + i, s->sigma[i]); out->found = 1; }
        if (fp) {
            fprintf(fp, "METHOD=ECM; SIGMA=%" PRIu64 "; B1=%" PRIu64 "; ", s->sigma[i], b1_printed);
            gmp_fprintf(fp, "N=0x%Zx; X=0x%Zx; Z=0x%Zx; PROGRAM=AVX-ECM;\n", out->N, x, z);
        }
    }
    mpz_clear(x); mpz_clear(z); mpz_clear(f);
    if (fp) fclose(fp);
    if (out->res) fflush(out->res);
}

// base32 empty: generic input.  Otherwise the curve arithmetic runs modulo the base number 2^k-c / 2^k+1 and
// only the factor checks use N (main.c:597-616, ecm.c:1108-1119).
static void run_shard(Shard *s, Output *out, const std::vector<uint32_t> *n32, const std::vector<uint32_t> *base32, uint64_t b1,
                      uint64_t b2, bool do2, uint32_t nranges)
{
    size_t section = 0;                                   // next file section this shard owes a turn for
    const size_t nsections = (size_t)nranges + (do2 ? 1 : 0);   // checkpoints, save_b1, stage-2 reports
    auto give_up = [&](const char *msg) {
        s->error = msg;
        for (; section < nsections; section++) out->in_turn(section, s->gpu, [] {});
    };
    ecm_b200_ctx *ctx = NULL;
    const int rc0 = base32->empty() ? ecm_b200_create(&ctx, s->gpu, n32->data(), (int)n32->size(), s->count)
                                    : ecm_b200_create_special(&ctx, s->gpu, base32->data(), (int)base32->size(), n32->data(),
                                                              (int)n32->size(), s->count);
    if (rc0) { give_up(ecm_b200_last_error()); return; }
    const int L = s->limbs = ecm_b200_limbs(ctx);
    const size_t words = (size_t)L * s->count;
    std::vector<uint32_t> X(words), Z(words), G(words);
    std::vector<uint8_t> fl(s->count);
    double t0 = now();
    int rc = ecm_b200_build_curves(ctx, s->count, s->sigma.data());
    s->t_build = now() - t0; t0 = now();
    if (!rc && nranges <= 1) rc = ecm_b200_stage1(ctx, b1);
    for (uint32_t r = 0; !rc && nranges > 1 && r < nranges; r++) {           // vececm's range loop, ecm.c:1207-1311
        uint64_t last = 0;
        rc = ecm_b200_stage1_range(ctx, b1, r, &last);
        if (rc || r + 1 == nranges) break;
        rc = ecm_b200_read_stage1(ctx, X.data(), Z.data(), fl.data(), G.data());
        if (rc) break;
        out->in_turn(section++, s->gpu, [&] { write_lines(out, s, "checkpoint.txt", last, true, X, Z, fl, G); });
    }
    if (!rc) rc = ecm_b200_read_stage1(ctx, X.data(), Z.data(), fl.data(), G.data());
    s->t_s1 = now() - t0; t0 = now();
    if (rc) { give_up(ecm_b200_last_error()); ecm_b200_destroy(ctx); return; }
    out->in_turn(section++, s->gpu, [&] { write_lines(out, s, "save_b1.txt", b1, false, X, Z, fl, G); });
    if (do2) {
        rc = ecm_b200_stage2(ctx, b1, b2);
        if (!rc) rc = ecm_b200_read_stage2(ctx, NULL, fl.data(), G.data(), NULL);
        s->t_s2 = now() - t0;
        if (rc) { give_up(ecm_b200_last_error()); ecm_b200_destroy(ctx); return; }
        out->in_turn(section++, s->gpu, [&] {
            mpz_t f; mpz_init(f);
            for (uint32_t i = 0; i < s->count; i++)
                if (fl[i]) { limbs_to_mpz(f, G, s->limbs, s->count, i); report(out->res, f, 2, b2, s->first + i, s->sigma[i]); out->found = 1; }
            mpz_clear(f);
            if (out->res) fflush(out->res);
        });
    }
    ecm_b200_destroy(ctx);
}

int main(int argc, char **argv)
{
    if (argc == 3 && strcmp(argv[1], "--eval") == 0) {          // expression evaluator only
        mpz_t v; mpz_init(v);
        std::string e = calc_eval(argv[2], v);
        if (!e.empty()) { printf("error: %s\n", e.c_str()); return 1; }
        gmp_printf("%Zd\n", v);
        return 0;
    }
    if (argc == 3 && strcmp(argv[1], "--classify") == 0) {      // input classification only (main.c:405-521)
        mpz_t v, b; mpz_init(v); mpz_init(b);
        std::string e = calc_eval(argv[2], v);
        if (!e.empty()) { printf("error: %s\n", e.c_str()); return 1; }
        const SpecialForm f = classify_input(v, b);
        gmp_printf("kind %ld k %d n %Zd base %Zd\n", f.kind, f.k, v, b);
        return 0;
    }
    if (argc < 4) {
        printf("usage: avx-ecm-b200 $input $numcurves $B1 [$gpus] [$B2] [$sigma]\n");
        return 1;
    }
    const double t_start = now();
    printf("starting process %d\n", (int)getpid());
    mpz_t N;
    mpz_init(N);
    std::string err = calc_eval(argv[1], N);
    if (!err.empty()) { printf("could not evaluate input expression: %s\n", err.c_str()); return 1; }
    if (mpz_cmp_ui(N, 3) < 0 || !mpz_odd_p(N)) { printf("input must be an odd integer > 2\n"); return 1; }

    uint32_t numcurves = (uint32_t)strtoul(argv[2], NULL, 10);
    const uint64_t b1 = strtoull(argv[3], NULL, 10);
    uint64_t b2 = 100ULL * b1;                                  // main.c:462
    int gpus = (argc >= 5) ? atoi(argv[4]) : 1;
    if (gpus < 1) gpus = 1;
    bool do2 = true;
    if (argc >= 6) { b2 = strtoull(argv[5], NULL, 10); if (b2 <= b1) { do2 = false; b2 = b1; } }   // main.c:543-552
    uint64_t sigma0 = (argc >= 7) ? strtoull(argv[6], NULL, 10) : 0;
    if (numcurves < (uint32_t)gpus) numcurves = gpus;

    // Mersenne-like inputs: classification, algebraic-factor removal, special base vs REDC (main.c:405-521)
    mpz_t base; mpz_init(base);
    const SpecialForm form = classify_input(N, base);
    if (mpz_cmp_ui(N, 3) < 0) { printf("nothing left to factor after removing algebraic factors\n"); return 1; }

    gmp_printf("commencing parallel ecm on %Zd\n", N);
    const int bits = (int)mpz_sizeinbase(N, 2);
    const int abits = form.kind ? (int)mpz_sizeinbase(base, 2) : bits;      // width of the arithmetic
    printf("ECM has been configured with 32-bit limbs on B200 (%d limbs), GMP_LIMB_BITS = %d\n", (abits + 31) / 32, GMP_LIMB_BITS);
    if (sigma0) printf("starting with sigma = %" PRIu64 "\n", sigma0);
    printf("Input has %d bits, using %d GPU(s) (%u curves/GPU)\n", bits, gpus, (numcurves + gpus - 1) / gpus);

    std::vector<uint32_t> n32((bits + 31) / 32, 0);
    size_t cnt = 0;
    mpz_export(n32.data(), &cnt, -1, 4, 0, 0, N);
    std::vector<uint32_t> base32;
    if (form.kind) { base32.assign((abits + 31) / 32, 0); mpz_export(base32.data(), &cnt, -1, 4, 0, 0, base); }

    // contiguous sigma slices, one per GPU; random sigmas when none was given (ecm.c:1564-1570)
    std::vector<Shard> shards(gpus);
    uint64_t lcg = (uint64_t)(t_start * 1e6) ^ ((uint64_t)getpid() << 32);
    uint32_t first = 0;
    for (int g = 0; g < gpus; g++) {
        Shard &s = shards[g];
        s.gpu = g; s.first = first;
        s.count = numcurves / gpus + ((uint32_t)g < numcurves % gpus ? 1 : 0);
        first += s.count;
        s.sigma.resize(s.count);
        for (uint32_t i = 0; i < s.count; i++) {
            if (sigma0) s.sigma[i] = sigma0 + s.first + i;
            else { uint64_t v; do { v = lcg_next(&lcg); } while (v < 6); s.sigma[i] = v; }
        }
    }
    printf("\nCommencing curves 0-%u of %u\n", numcurves - 1, numcurves);
    uint32_t nranges = 1;
    if (ecm_b200_stage1_ranges(b1, &nranges) || nranges < 1) nranges = 1;
    Output out;
    mpz_init(out.N); mpz_set(out.N, N);
    out.res = fopen("ecm_results.txt", "a");
    out.turn.assign((size_t)nranges + (do2 ? 1 : 0), 0);
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; g++) th.emplace_back(run_shard, &shards[g], &out, &n32, &base32, b1, b2, do2, nranges);
    for (auto &t : th) t.join();
    if (out.res) fclose(out.res);
    double ts1 = 0, ts2 = 0, tb = 0;
    int failed = 0;
    for (auto &s : shards) {
        if (!s.error.empty()) { printf("GPU %d: %s\n", s.gpu, s.error.c_str()); failed = 1; }
        ts1 = std::max(ts1, s.t_s1); ts2 = std::max(ts2, s.t_s2); tb = std::max(tb, s.t_build);
    }
    printf("Building curves took %1.4f seconds.\n", tb);
    printf("Stage 1 took %1.4f seconds\n", ts1);
    if (do2) printf("Stage 2 took %1.4f seconds\n", ts2);
    printf("Process took %1.4f seconds.\n", now() - t_start);
    return failed;
}
