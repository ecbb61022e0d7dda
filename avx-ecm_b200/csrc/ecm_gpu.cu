// ecm_gpu.cu -- C ABI of libecm_b200.so (include/ecm_b200.h): context, launch scheduling and
// the host<->device plumbing.  No CPU fallback: without a usable CUDA device every entry point
// fails with ECM_B200_ENODEV.
#include "../../include/ecm_b200.h"
#include "engine.hpp"
#include "plan.hpp"
#include "plan2.hpp"
#include "rv_prog.hpp"
#include <cuda_runtime.h>
#include <atomic>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <vector>

using namespace ecmb200;
enum { NSLOT_S1 = 13, SP = 12 };   // slot file of the stage-1 machine (vm.cuh)

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ECM_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)

// ---- tiny host big-number helpers (little-endian 32-bit limbs, fixed length) -------------------
int cmp(const Big &a, const Big &b)
{
    for (size_t i = a.size(); i-- > 0;) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}
uint32_t add_in(Big &a, const Big &b)
{
    uint64_t c = 0;
    for (size_t i = 0; i < a.size(); i++) { c += (uint64_t)a[i] + b[i]; a[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
void sub_in(Big &a, const Big &b)
{
    int64_t c = 0;
    for (size_t i = 0; i < a.size(); i++) { c += (int64_t)a[i] - b[i]; a[i] = (uint32_t)c; c >>= 32; }
}
void dbl_mod(Big &x, const Big &n)       // x = 2x mod n
{
    uint32_t top = x.back() >> 31;
    for (size_t i = x.size(); i-- > 0;) x[i] = (x[i] << 1) | (i ? x[i - 1] >> 31 : 0);
    if (top || cmp(x, n) >= 0) sub_in(x, n);
}
void half_mod(Big &x, const Big &n)      // x = x/2 mod n (n odd)
{
    uint32_t carry = 0;
    if (x[0] & 1) carry = add_in(x, n);
    for (size_t i = 0; i < x.size(); i++) x[i] = (x[i] >> 1) | ((i + 1 < x.size() ? x[i + 1] : carry) << 31);
}
uint32_t bitlen(const Big &x)
{
    for (size_t i = x.size(); i-- > 0;) if (x[i]) return (uint32_t)(32 * i + 32 - __builtin_clz(x[i]));
    return 0;
}

// ---- engines: one translation unit per compiled limb count (engine_inst.cu) ---------------------
#ifndef ECM_B200_LIMB_SET
#error "ECM_B200_LIMB_SET must list the compiled limb counts, e.g. X(13) X(32)"
#endif
}  // namespace
namespace ecmb200 {
#define X(n) Engine *make_engine_##n();
ECM_B200_LIMB_SET
#undef X
#ifdef ECM_B200_SPECIAL_LIMB_SET
#define X(n) Engine *make_engine_sp_##n();
ECM_B200_SPECIAL_LIMB_SET
#undef X
#endif
void count_launch() { g_launches++; }
}
namespace {
Engine *make_engine(int nlimbs)
{
#define X(n) if (nlimbs <= n) return make_engine_##n();
    ECM_B200_LIMB_SET
#undef X
    return nullptr;
}
// shift-and-fold kernels for a base 2^kbits -+ c, when they were compiled for a suitable limb count
Engine *make_special_engine(int nlimbs, uint32_t kbits)
{
    (void)nlimbs; (void)kbits;
#ifdef ECM_B200_SPECIAL_LIMB_SET
    // the smallest compiled limb count that holds the base AND has bit kbits strictly inside its limbs
#define X(n) if (nlimbs <= n) { Engine *e = make_engine_sp_##n(); if (e->serves_special(kbits)) return e; delete e; }
    ECM_B200_SPECIAL_LIMB_SET
#undef X
#endif
    return nullptr;
}
// n = 2^k - c with 1 <= c < 2^31, or n = 2^k + 1 ?  (the forms main.c:408-441 detects)
bool special_shape(const uint32_t *n, int nlimbs, int *kind, uint32_t *kbits, uint32_t *cval)
{
    Big v(n, n + nlimbs);
    const uint32_t bl = bitlen(v);
    if (bl < 64) return false;
    // 2^k + 1: bit k, bit 0, nothing between
    bool plus = (v[0] == 1) && (v[(bl - 1) >> 5] == (1u << ((bl - 1) & 31)));
    for (int i = 1; plus && i < (int)((bl - 1) >> 5); i++) if (v[i]) plus = false;
    if (plus && bl >= 65) { *kind = -1; *kbits = bl - 1; *cval = 1; return true; }
    // 2^k - c, k = bitlen: every bit of limbs 1.. below k set, limb 0 = 2^32 - c
    for (uint32_t b = 32; b < bl; b++) if (!((v[b >> 5] >> (b & 31)) & 1u)) return false;
    const uint32_t c = 0u - v[0];
    if (c == 0 || c >= (1u << 31)) return false;
    *kind = 1; *kbits = bl; *cval = c;
    return true;
}

}  // namespace

struct ecm_b200_ctx {
    int device = 0;
    Engine *eng = nullptr;
    int nl = 0;                 // engine limbs
    Big n;                      // modulus padded to nl limbs
    Big chk;                    // modulus of the factor checks (= n unless special-form)
    bool special = false;
    bool fold = false;          // shift-and-fold kernels (special-form base) instead of Montgomery
    uint32_t max_curves = 0, count = 0, groups = 0;
    uint32_t T = 0, groups_max = 0;   // curves per group (one stage-1 block), groups allocated
    uint32_t threads_s1 = 0;          // threads of a stage-1 block: T * lanes per curve
    Geom G1{0, 0, NSLOT_S1};
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    uint32_t *d_state = nullptr;
    size_t state_words = 0;
    uint8_t *d_ops = nullptr; size_t d_ops_cap = 0;
    uint32_t *d_io = nullptr; size_t d_io_words = 0;     // staging for host<->device transfers
    void *d_params = nullptr;
    uint32_t *d_chk = nullptr;  // modulus of the factor checks (the input N), nl limbs
    void *d_flush = nullptr;
    uint8_t *d_flags = nullptr;
    Stage1Plan plan;            // cached for plan.b1
    bool plan_on_device = false;
    // stage-1 schedule
    uint32_t chunk_len = 0; uint64_t total_items = 0, next_item = 0; uint32_t launches_total = 0, launches_issued = 0;
    int p_slot = 0;             // physical point slot holding P
    uint64_t seg_begin = 0, seg_len = 0; int seg_final_slot = 0; bool seg_last = true, seg_done = false;   // op-stream segment being run
    uint32_t next_range = 0;    // range-by-range stage 1: the next prime range to run
    bool have_curves = false, stage1_done = false;
    float last_ms = 0; uint32_t last_launches = 0;
    // stage 2
    Stage2Program prog2; uint64_t prog2_b1 = 0, prog2_b2 = 0;
    uint32_t *d_acc = nullptr; uint8_t *d_fail = nullptr;       // per batch curve: accumulator (Montgomery form), inversion-failure flag
    bool stage2_done = false;
    float s2_ms = 0; uint32_t s2_launches = 0; uint32_t s2_waves = 0;
    // stage-2 wave buffers, kept across calls (s2_reserve)
    uint32_t *d_tab = nullptr, *d_state2 = nullptr; uint8_t *d_wfail = nullptr; uint64_t *d_code = nullptr;
    size_t s2_tab_bytes = 0, s2_state_bytes = 0, s2_code_bytes = 0; uint32_t s2_fail_cap = 0;
    uint32_t s2_cap = 0, s2_first = 0, s2_n = 0, s2_groups = 0;      // current wave
    bool s2_open = false;       // ecm_b200_stage2_init done: tables resident, ecm_b200_stage2_range may follow
    // ECM_B200_S2_TRACE=1: device time of stage 2 split by kernel kind (event pairs around every segment), printed to stderr
    bool s2_trace = false;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> s2_marks;     // kind (0 = slot machine, 1 = pair kernel), begin, end
};

extern "C" {

const char *ecm_b200_last_error(void) { return g_err.c_str(); }
uint64_t ecm_b200_launch_count(void) { return g_launches.load(); }
int ecm_b200_limbs(const ecm_b200_ctx *ctx) { return ctx ? ctx->nl : 0; }
int ecm_b200_uses_fold(const ecm_b200_ctx *ctx) { return ctx && ctx->fold ? 1 : 0; }

static int create_ctx(ecm_b200_ctx **out, int device, const uint32_t *n, int nlimbs, const uint32_t *chk, int chklimbs,
                      uint32_t max_curves);

int ecm_b200_create(ecm_b200_ctx **out, int device, const uint32_t *n, int nlimbs, uint32_t max_curves)
{
    return create_ctx(out, device, n, nlimbs, nullptr, 0, max_curves);
}

int ecm_b200_create_special(ecm_b200_ctx **out, int device, const uint32_t *base, int baselimbs, const uint32_t *n, int nlimbs,
                            uint32_t max_curves)
{
    if (!n || nlimbs < 1) return fail(ECM_B200_EINVAL, "bad argument");
    return create_ctx(out, device, base, baselimbs, n, nlimbs, max_curves);
}

// chk != nullptr: special-form input (main.c:405-457, 597-616).  All arithmetic is done modulo the base number
// `n` = 2^k-1 / 2^k+1 / 2^k-c; the factor checks use the input `chk` (ecm.c:1108-1119); residues are reported
// mod the base number; a failed inversion leaves the plain gcd / plain operand (no Montgomery form on that path,
// ecm.c:1903-1946).
static int create_ctx(ecm_b200_ctx **out, int device, const uint32_t *n, int nlimbs, const uint32_t *chk, int chklimbs,
                      uint32_t max_curves)
{
    if (!out || !n || nlimbs < 1 || max_curves < 1) return fail(ECM_B200_EINVAL, "bad argument");
    *out = nullptr;
    while (nlimbs > 1 && n[nlimbs - 1] == 0) nlimbs--;
    if (!(n[0] & 1) || (nlimbs == 1 && n[0] < 3)) return fail(ECM_B200_EINVAL, "modulus must be odd and > 1");
    if (chk) {
        while (chklimbs > 1 && chk[chklimbs - 1] == 0) chklimbs--;
        if (!(chk[0] & 1) || (chklimbs == 1 && chk[0] < 3)) return fail(ECM_B200_EINVAL, "input number must be odd and > 1");
        if (chklimbs > nlimbs) return fail(ECM_B200_EINVAL, "input number is larger than its base number");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
        return fail(ECM_B200_ENODEV, "no usable CUDA device (this engine has no CPU fallback)");
    // special-form base: shift-and-fold kernels when the base has one of the reference's shapes and the kernel
    // set covers its length; otherwise (and with ECM_B200_NO_FOLD set) the Montgomery kernels modulo the base
    Engine *eng = nullptr;
    int sp_kind = 0; uint32_t sp_k = 0, sp_c = 0;
    if (chk && !getenv("ECM_B200_NO_FOLD") && special_shape(n, nlimbs, &sp_kind, &sp_k, &sp_c)) eng = make_special_engine(nlimbs, sp_k);
    const bool fold = eng != nullptr;
    if (!eng) eng = make_engine(nlimbs);
    if (!eng) return fail(ECM_B200_EINVAL, "modulus too large: kernels are built for up to 2048 bits (64 limbs)");
    ecm_b200_ctx *c = new ecm_b200_ctx();
    c->device = device; c->eng = eng; c->nl = eng->nl; c->max_curves = max_curves;
    const int nl = c->nl;
    c->n.assign(nl, 0);
    for (int i = 0; i < nlimbs; i++) c->n[i] = n[i];
    // Montgomery constants for R = 2^(32*nl)  (main.c:624-640 does this with GMP for R = 2^MAXBITS)
    Big one(nl, 0); one[0] = 1;
    for (int i = 0; i < 32 * nl; i++) dbl_mod(one, c->n);          // R mod N
    Big r2 = one; for (int i = 0; i < 32 * nl; i++) dbl_mod(r2, c->n);
    Big r3 = r2; for (int i = 0; i < 32 * nl; i++) dbl_mod(r3, c->n);
    // reference R: MAXBITS = smallest multiple of 208 strictly above bitlen(N) (main.c:465-483)
    uint32_t maxbits_ref = 208; while (maxbits_ref <= bitlen(c->n)) maxbits_ref += 208;
    Big rri = one; for (uint32_t i = 0; i < maxbits_ref; i++) half_mod(rri, c->n);   // R * 2^-MAXBITS mod N
    Big rref(nl, 0); rref[0] = 1;
    for (uint32_t i = 0; i < maxbits_ref; i++) dbl_mod(rref, c->n);                  // 2^MAXBITS mod N, a plain integer
    c->chk = c->n;
    if (fold) {                                           // plain residues: "R = 1"
        one.assign(nl, 0); one[0] = 1;
        r2 = one; r3 = one;
    }
    if (chk) {                                            // special form: plain residues, R_ref = 1
        rri = one;
        rref.assign(nl, 0); rref[0] = 1;
        c->chk.assign(nl, 0);
        for (int i = 0; i < chklimbs; i++) c->chk[i] = chk[i];
        c->special = true;
    }
    uint32_t inv = 1; for (int i = 0; i < 5; i++) inv *= 2 - c->n[0] * inv;           // N^-1 mod 2^32
    const uint32_t m0inv = 0u - inv;
    eng->set_params(c->n, one, r2, r3, rri, rref, m0inv);
    if (fold) eng->set_special(sp_kind, sp_k, sp_c);
    c->fold = fold;

#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::string m = std::string(#call) + ": " + cudaGetErrorString(e_); ecm_b200_destroy(c); return fail(ECM_B200_ECUDA, m); } } while (0)
    CUC(cudaSetDevice(device));
    CUC(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
    // Which stage-1 kernel: the register-resident macro-op machine (rv.cuh) where it is compiled for this limb count --
    // one thread per curve up to 16 limbs, four lanes per curve at 48/64 limbs -- else the slot-file machine (vm.cuh).
    // The fold kernel set uses it too (2^415-1, 65 536 curves: 88.4 k curves/s against 71.2 k on the slot-file machine).
    // ECM_B200_S1_KERNEL=vm|rv overrides (A/B measurements, tests that run both).
    {
        bool rv = eng->rv_max_threads != 0;
        if (const char *e = getenv("ECM_B200_S1_KERNEL")) {
            if (!strcmp(e, "vm")) rv = false;
            else if (!strcmp(e, "rv")) rv = eng->rv_max_threads != 0;
        }
        eng->use_rv = rv;
        bool cs2 = eng->has_coop_s2;                      // stage 2 on the four-lanes-per-curve layout (coop_s2.cuh)
        if (const char *e = getenv("ECM_B200_S2_KERNEL")) { if (!strcmp(e, "solo")) cs2 = false; }
        eng->use_coop_s2 = cs2;
    }
    CUC(eng->prepare());
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUC(cudaEventCreate(&c->ev0)); CUC(cudaEventCreate(&c->ev1));
    // Curves per group (= one stage-1 block).  Whole multiples of 4 warps keep the four SM
    // sub-partitions evenly loaded (14 warps = 4,4,3,3 measured 6.9 Tprod/s vs 7.2 with 12).  More
    // resident warps hide more latency (throughput ~ w/(w+3.6), measured 12..24 warps), but a batch
    // with fewer groups than SMs leaves SMs idle, while more groups than SMs are time-sliced at full
    // occupancy by the launch schedule.  Pick the block size that maximises the product.
    {
        const uint32_t Lc = eng->use_rv ? (uint32_t)eng->rv_lanes : 1u;        // lanes per curve
        uint32_t S = eng->use_rv ? (uint32_t)eng->rv_max_threads : (uint32_t)eng->stride_s1;
        const uint32_t sms = (uint32_t)c->num_sms;
        // one thread per curve from 20 limbs up: 8 warps per SM beat 12 whatever the machine (65 536 curves, Tprod/s at 256
        // against 384 threads: 7.98 / 7.81 at 20 limbs, 7.65 / 7.47 at 24, 8.09 / 6.80 at 28 -- profiles/r2v_s1_*.log), which
        // the warps-hide-latency model below does not know
        if (!fold && Lc == 1 && nl >= 20 && S > 256) S = 256;
        // the fold kernels (half the IMADs per product, same loads and carry chains) are bound by dependent-issue
        // latency: measured 67.0k -> 72.8k curves/s from 12 to 16 warps even with 20 of 148 SMs left idle
        const double lat = fold ? 40.0 : 3.6;
        uint32_t bestT = 0; double best = 0;
        auto consider = [&](uint32_t T) {                 // T = threads per block
            const uint32_t Tc = T / Lc, gr = (max_curves + Tc - 1) / Tc;
            const double w = T / 32.0;
            const double score = (double)std::min(gr, sms) / sms * (w / (w + lat)) * ((double)max_curves / ((double)gr * Tc));
            if (score > best) { best = score; bestT = T; }
        };
        // small batches: 1-2 warps per block spread over more SMs finish sooner than a few full blocks
        for (uint32_t T : {32u, 64u}) if (T <= S) consider(T);
        for (uint32_t T = 128; T <= S; T += 128) consider(T);
        if (bestT == 0) bestT = S;                        // kernels whose smem budget allows < 128 threads
        if (const char *e = getenv("ECM_B200_THREADS")) { uint32_t t = (uint32_t)atoi(e); if (t >= 32 && t <= S && t % 32 == 0) bestT = t; }
        c->threads_s1 = bestT;
        c->T = bestT / Lc;
        c->groups_max = (max_curves + c->T - 1) / c->T;
        const uint32_t stride = (uint32_t)eng->stride_for_threads(bestT);
        c->G1 = Geom{c->T, stride, NSLOT_S1, Lc};
    }
    c->state_words = (size_t)c->groups_max * NSLOT_S1 * (nl / c->G1.L) * c->G1.stride;
    CUC(cudaMalloc(&c->d_state, c->state_words * 4));
    CUC(cudaMemsetAsync(c->d_state, 0, c->state_words * 4, c->stream));
    c->d_io_words = (size_t)4 * nl * max_curves + 64;
    CUC(cudaMalloc(&c->d_io, c->d_io_words * 4));
    CUC(cudaMalloc(&c->d_flags, (size_t)max_curves * 2 + 64));
    CUC(cudaMalloc(&c->d_chk, (size_t)nl * 4));
    CUC(cudaMemcpyAsync(c->d_chk, c->chk.data(), (size_t)nl * 4, cudaMemcpyHostToDevice, c->stream));
    CUC(cudaMalloc(&c->d_params, eng->params_bytes));
    CUC(cudaMemcpyAsync(c->d_params, eng->params_host(), eng->params_bytes, cudaMemcpyHostToDevice, c->stream));
    eng->set_params_device(c->d_params);
    c->s2_trace = getenv("ECM_B200_S2_TRACE") != nullptr;
    CUC(cudaStreamSynchronize(c->stream));
#undef CUC
    *out = c;
    return ECM_B200_OK;
}

void ecm_b200_destroy(ecm_b200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_state); cudaFree(c->d_ops); cudaFree(c->d_io); cudaFree(c->d_flags); cudaFree(c->d_params); cudaFree(c->d_chk); cudaFree(c->d_flush); cudaFree(c->d_acc); cudaFree(c->d_fail);
    cudaFree(c->d_tab); cudaFree(c->d_state2); cudaFree(c->d_wfail); cudaFree(c->d_code);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c->eng;
    delete c;
}

static int set_count(ecm_b200_ctx *c, uint32_t count)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    if (count < 1 || count > c->max_curves) return fail(ECM_B200_EINVAL, "count exceeds the context's max_curves");
    c->count = count;
    c->groups = (count + c->T - 1) / c->T;
    c->have_curves = true; c->stage1_done = false; c->p_slot = 0; c->next_range = 0; c->seg_done = false;
    c->total_items = c->next_item = 0; c->launches_total = c->launches_issued = 0;
    c->s2_open = false; c->stage2_done = false;
    return ECM_B200_OK;
}

int ecm_b200_load_curves(ecm_b200_ctx *c, uint32_t count, const uint32_t *x, const uint32_t *s)
{
    if (!c || !x || !s) return fail(ECM_B200_EINVAL, "null argument");
    int rc = set_count(c, count); if (rc) return rc;
    CU(cudaSetDevice(c->device));
    const size_t words = (size_t)c->nl * count;
    CU(cudaMemcpyAsync(c->d_io, x, words * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_io + words, s, words * 4, cudaMemcpyHostToDevice, c->stream));
    c->eng->load_curves(c->stream, c->d_state, c->G1, c->groups * c->T, count, c->d_io, c->d_io + words);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return ECM_B200_OK;
}

int ecm_b200_build_curves(ecm_b200_ctx *c, uint32_t count, const uint64_t *sigma)
{
    if (!c || !sigma) return fail(ECM_B200_EINVAL, "null argument");
    int rc = set_count(c, count); if (rc) return rc;
    CU(cudaSetDevice(c->device));
    // u = sigma^2 - 5, v = 4 sigma (ecm.c:1587-1598), reduced mod N when N is below 2^128
    std::vector<uint32_t> uv((size_t)8 * count);
    unsigned __int128 nsmall = 0; bool small = bitlen(c->n) <= 127;
    if (small) for (int i = 3; i >= 0; i--) nsmall = (nsmall << 32) | (i < c->nl ? c->n[i] : 0);
    for (uint32_t i = 0; i < count; i++) {
        if (sigma[i] < 6) return fail(ECM_B200_EINVAL, "sigma must be >= 6");
        unsigned __int128 u = (unsigned __int128)sigma[i] * sigma[i] - 5, v = (unsigned __int128)sigma[i] * 4;
        if (small) { u %= nsmall; v %= nsmall; }
        for (int k = 0; k < 4; k++) { uv[(size_t)k * count + i] = (uint32_t)(u >> (32 * k)); uv[(size_t)(4 + k) * count + i] = (uint32_t)(v >> (32 * k)); }
    }
    if ((size_t)8 * count > c->d_io_words) return fail(ECM_B200_ENOMEM, "staging buffer too small");
    CU(cudaMemcpyAsync(c->d_io, uv.data(), uv.size() * 4, cudaMemcpyHostToDevice, c->stream));
    c->eng->build_curves(c->stream, c->d_state, c->G1, c->groups * c->T, count, c->d_io, c->d_flags);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return ECM_B200_OK;
}

static bool s1_in_progress(const ecm_b200_ctx *c)
{
    return (c->total_items != 0 && !c->seg_done) || (c->next_range != 0 && !c->stage1_done);
}

// plan for b1 on the device; the schedule of one segment [begin, end) of its op stream
static int stage1_prepare(ecm_b200_ctx *c, uint64_t b1)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    if (!c->have_curves) return fail(ECM_B200_ESTATE, "no curves loaded");
    // beyond 1e8 the stream follows the reference's range-by-range driver, quirks included (plan.cpp); the cap keeps
    // p*p inside 64 bits (ecm.c:1832) and the op stream (about 2 bytes per unit of B1) inside host memory
    if (b1 < 2 || b1 > 2000000000ull) return fail(ECM_B200_EINVAL, "B1 out of range (2 .. 2e9)");
    CU(cudaSetDevice(c->device));
    if (c->plan.b1 != b1) { plan_stage1(b1, c->plan); c->plan_on_device = false; }
    if (!c->plan_on_device) {
        if (c->plan.ops.size() > c->d_ops_cap) {
            cudaFree(c->d_ops); c->d_ops = nullptr; c->d_ops_cap = 0;
            CU(cudaMalloc(&c->d_ops, c->plan.ops.size() + 16));
            c->d_ops_cap = c->plan.ops.size();
        }
        CU(cudaMemcpyAsync(c->d_ops, c->plan.ops.data(), c->plan.ops.size(), cudaMemcpyHostToDevice, c->stream));
        c->plan_on_device = true;
    }
    if (c->stage1_done) return fail(ECM_B200_ESTATE, "stage 1 already run on this batch");
    if (c->total_items != 0 && !c->seg_done)
        return fail(ECM_B200_ESTATE, "a stage-1 segment is still in progress on this batch (finish it with ecm_b200_stage1_step, or build the curves again)");
    return ECM_B200_OK;
}

static int stage1_segment(ecm_b200_ctx *c, uint64_t begin, uint64_t end, int final_slot, bool last)
{
    // Schedule: an item is (group of THREADS curves, chunk of the op stream).  Items are ordered
    // chunk-major and a launch takes a run of consecutive items, at most one per SM; because a run
    // is never longer than the number of groups, item (g, c-1) is always in an earlier launch.
    const uint64_t nops = end - begin;
    uint32_t chunk = 32768;
    if (c->groups > (uint32_t)c->num_sms) {
        // several waves: finer chunks keep the last partial launch small
        chunk = 8192;
    }
    c->seg_begin = begin; c->seg_len = nops; c->seg_final_slot = final_slot; c->seg_last = last;
    c->chunk_len = chunk;
    const uint64_t nchunks = (nops + chunk - 1) / chunk;
    c->total_items = nchunks * c->groups;
    c->next_item = 0;
    const uint32_t per = std::min<uint32_t>(c->groups, (uint32_t)c->num_sms);
    c->launches_total = (uint32_t)((c->total_items + per - 1) / per);
    c->launches_issued = 0;
    c->last_launches = 0;
    c->seg_done = false;
    CU(cudaEventRecord(c->ev0, c->stream));
    if (c->total_items == 0) {                    // nothing to do (B1 = 2): the point is unchanged
        c->seg_done = true;
        c->stage1_done = last;
        c->p_slot = final_slot;
        CU(cudaEventRecord(c->ev1, c->stream));
    }
    return ECM_B200_OK;
}

int ecm_b200_stage1_begin(ecm_b200_ctx *c, uint64_t b1)
{
    int rc = stage1_prepare(c, b1); if (rc) return rc;
    if (c->next_range != 0) return fail(ECM_B200_ESTATE, "stage 1 was started range by range on this batch");
    return stage1_segment(c, 0, c->plan.ops.size(), c->plan.final_slot, true);
}

int ecm_b200_stage1_ranges(uint64_t b1, uint32_t *count)
{
    if (!count || b1 < 2) return fail(ECM_B200_EINVAL, "bad argument");
    const uint64_t w = stage1_prime_range();
    *count = (uint32_t)((b1 + w - 1) / w);
    return ECM_B200_OK;
}

// One prime range of stage 1 (ecm.c:1207-1234), run to completion; ranges must be taken in order.  Afterwards
// read_stage1 returns the state the reference writes to checkpoint.txt (ecm.c:1237-1311) -- or, after the last
// range, the stage-1 result.
int ecm_b200_stage1_range(ecm_b200_ctx *c, uint64_t b1, uint32_t range, uint64_t *last_prime)
{
    int rc = stage1_prepare(c, b1); if (rc) return rc;
    const auto &re = c->plan.range_end;
    if (range >= re.size()) return fail(ECM_B200_EINVAL, "no such prime range for this B1");
    if (range != c->next_range) return fail(ECM_B200_ESTATE, "prime ranges must be run in order, starting from freshly built curves");
    const uint64_t begin = range ? re[range - 1].ops : 0;
    const bool last = range + 1 == re.size();
    rc = stage1_segment(c, begin, last ? c->plan.ops.size() : re[range].ops, re[range].slot, last); if (rc) return rc;
    int done = 0;
    rc = ecm_b200_stage1_step(c, 0xffffffffu, &done); if (rc) return rc;
    c->next_range = range + 1;
    if (last_prime) *last_prime = re[range].last_prime;
    return ecm_b200_sync(c);
}

int ecm_b200_stage1_step(ecm_b200_ctx *c, uint32_t max_launches, int *done)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    if (c->total_items == 0 && c->seg_done) { if (done) *done = 1; return ECM_B200_OK; }   // empty op stream (B1 = 2)
    if (c->total_items == 0) return fail(ECM_B200_ESTATE, "stage1_begin not called");
    CU(cudaSetDevice(c->device));
    const uint32_t per = std::min<uint32_t>(c->groups, (uint32_t)c->num_sms);
    uint32_t n = 0;
    while (c->next_item < c->total_items && n < max_launches) {
        const uint32_t blocks = (uint32_t)std::min<uint64_t>(per, c->total_items - c->next_item);
        c->eng->stage1(c->stream, blocks, c->threads_s1, c->d_state, c->d_ops + c->seg_begin, c->seg_len, c->chunk_len, c->groups, c->next_item);
        c->next_item += blocks; c->launches_issued++; c->last_launches++; n++;
    }
    CU(cudaGetLastError());
    const bool fin = c->next_item >= c->total_items;
    if (fin && !c->seg_done) {
        c->seg_done = true;
        c->stage1_done = c->seg_last;
        c->p_slot = c->seg_final_slot;
        CU(cudaEventRecord(c->ev1, c->stream));
    }
    if (done) *done = fin ? 1 : 0;
    return ECM_B200_OK;
}

int ecm_b200_stage1_launches(const ecm_b200_ctx *c, uint32_t *total, uint32_t *issued)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    if (total) *total = c->launches_total;
    if (issued) *issued = c->launches_issued;
    return ECM_B200_OK;
}

int ecm_b200_stage1_progress(const ecm_b200_ctx *c, double *fraction)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    // every item covers chunk_len ops except those of the last chunk
    if (fraction) {
        const uint64_t nops = c->seg_len;
        if (c->total_items == 0 || nops == 0) { *fraction = 0; return ECM_B200_OK; }
        const uint64_t full_chunks = c->next_item / c->groups, rest = c->next_item % c->groups;
        const uint64_t nchunks = c->total_items / c->groups;
        auto chunk_ops = [&](uint64_t ch) { return ch + 1 < nchunks ? (uint64_t)c->chunk_len : nops - (nchunks - 1) * c->chunk_len; };
        double done = 0;
        for (uint64_t ch = 0; ch < full_chunks; ch++) done += (double)chunk_ops(ch) * c->groups;
        if (rest) done += (double)chunk_ops(full_chunks) * rest;
        *fraction = done / ((double)nops * c->groups);
    }
    return ECM_B200_OK;
}

int ecm_b200_timer(ecm_b200_ctx *c, int what, float *ms)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    if (!c->ev_t0) { CU(cudaEventCreate(&c->ev_t0)); CU(cudaEventCreate(&c->ev_t1)); }
    if (what == 0) CU(cudaEventRecord(c->ev_t0, c->stream));
    else if (what == 1) CU(cudaEventRecord(c->ev_t1, c->stream));
    else {
        CU(cudaEventSynchronize(c->ev_t1));
        float t = 0; CU(cudaEventElapsedTime(&t, c->ev_t0, c->ev_t1));
        if (ms) *ms = t;
    }
    return ECM_B200_OK;
}

int ecm_b200_flush_l2(ecm_b200_ctx *c)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    const size_t bytes = 256u << 20;
    if (!c->d_flush) CU(cudaMalloc(&c->d_flush, bytes));
    CU(cudaMemsetAsync(c->d_flush, (int)(c->launches_issued & 0xff), bytes, c->stream));
    return ECM_B200_OK;
}

int ecm_b200_sync(ecm_b200_ctx *c)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->stage1_done && c->last_launches) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_ms = ms;
    }
    return ECM_B200_OK;
}

int ecm_b200_stage1(ecm_b200_ctx *c, uint64_t b1)
{
    int rc = ecm_b200_stage1_begin(c, b1); if (rc) return rc;
    int done = 0;
    rc = ecm_b200_stage1_step(c, 0xffffffffu, &done); if (rc) return rc;
    return ecm_b200_sync(c);
}

int ecm_b200_last_timing(const ecm_b200_ctx *c, float *stage_ms, uint32_t *launches)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    if (stage_ms) *stage_ms = c->last_ms;
    if (launches) *launches = c->last_launches;
    return ECM_B200_OK;
}

int ecm_b200_read_stage1(ecm_b200_ctx *c, uint32_t *x, uint32_t *z, uint8_t *factor_flag, uint32_t *gcd_out)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    if (!c->have_curves) return fail(ECM_B200_ESTATE, "no curves loaded");
    CU(cudaSetDevice(c->device));
    const size_t words = (size_t)c->nl * c->count;
    uint32_t *dx = c->d_io, *dz = c->d_io + words, *dg = c->d_io + 2 * words;
    const bool want_flag = factor_flag || gcd_out;
    c->eng->read_point(c->stream, c->d_state, c->G1, c->count, 2 * c->p_slot, 2 * c->p_slot + 1, x ? dx : nullptr, dz,
                       want_flag ? c->d_flags : nullptr, gcd_out ? dg : nullptr, c->d_chk);
    CU(cudaGetLastError());
    if (x) CU(cudaMemcpyAsync(x, dx, words * 4, cudaMemcpyDeviceToHost, c->stream));
    if (z) CU(cudaMemcpyAsync(z, dz, words * 4, cudaMemcpyDeviceToHost, c->stream));
    if (factor_flag) CU(cudaMemcpyAsync(factor_flag, c->d_flags, c->count, cudaMemcpyDeviceToHost, c->stream));
    if (gcd_out) CU(cudaMemcpyAsync(gcd_out, dg, words * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return ECM_B200_OK;
}

struct S2Mark {        // one traced segment (no-op unless the context traces)
    ecm_b200_ctx *c; cudaEvent_t e0 = nullptr, e1 = nullptr; int kind;
    S2Mark(ecm_b200_ctx *c_, int kind_) : c(c_), kind(kind_) {
        if (!c->s2_trace) return;
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream);
    }
    ~S2Mark() { if (e0) { cudaEventRecord(e1, c->stream); c->s2_marks.push_back({kind, {e0, e1}}); } }
};

// after a stream synchronize: sum and print the traced segments
static void s2_trace_report(ecm_b200_ctx *c, const char *what)
{
    if (!c->s2_trace) return;
    double ms[2] = {0, 0}; size_t n[2] = {0, 0};
    for (auto &m : c->s2_marks) {
        float t = 0; cudaEventElapsedTime(&t, m.second.first, m.second.second);
        ms[m.first] += t; n[m.first]++;
        cudaEventDestroy(m.second.first); cudaEventDestroy(m.second.second);
    }
    c->s2_marks.clear();
    fprintf(stderr, "[ecm_b200 s2 trace] %s: slot machine %.1f ms in %zu segments, pair kernel %.1f ms in %zu segments\n", what, ms[0], n[0], ms[1], n[1]);
}

// Generic part of a stage-2 program: the slot-file machine over a wave of `groups` curve groups with the
// same round-robin (group, chunk) schedule as stage 1.
static int run_vm2(ecm_b200_ctx *c, const uint64_t *d_code, uint64_t ncode, uint32_t *state2, uint32_t cap2, uint32_t *tab,
                   uint32_t groups, uint8_t *inv_fail)
{
    if (ncode == 0) return ECM_B200_OK;
    S2Mark mark(c, 0);
    const uint32_t per = std::min<uint32_t>(groups, (uint32_t)c->num_sms);
    // Items are (group, chunk), chunk-major, at most `per` per launch.  With more groups than SMs one chunk per segment
    // leaves the last launch of every segment part empty (256 groups on 148 SMs: 148 + 108 blocks, 86 % of the machine);
    // cutting the segment into k chunks lets the launches interleave groups of neighbouring chunks, so that only the very
    // last launch is short.  k = the split (at least 32 instructions per chunk) that wastes the fewest block slots.
    uint32_t k = 1;
    if (groups > per) {
        double best = 0;
        for (uint32_t t = 1; t <= 16 && ncode / t >= 32; t++) {
            const uint64_t it = (uint64_t)groups * t, launches = (it + per - 1) / per;
            const double eff = (double)it / (double)(launches * per) - 0.002 * t;       // small price per extra chunk (slot reloads)
            if (eff > best) { best = eff; k = t; }
        }
    }
    const uint32_t chunk = (uint32_t)std::min<uint64_t>(65536, (ncode + k - 1) / k);
    const uint64_t nchunks = (ncode + chunk - 1) / chunk;
    const uint64_t items = nchunks * groups;
    for (uint64_t it = 0; it < items;) {
        const uint32_t blocks = (uint32_t)std::min<uint64_t>(per, items - it);
        c->eng->vm2(c->stream, blocks, state2, cap2, tab, d_code, ncode, chunk, groups, it, inv_fail);
        c->s2_launches++;
        it += blocks;
    }
    CU(cudaGetLastError());
    return ECM_B200_OK;
}

// A compiled program is dispatched in segments: maximal runs of pair steps go to the register-resident
// pair kernel, everything else (window builds, inversions, ladders) to the slot-file machine.
static int run_program(ecm_b200_ctx *c, const std::vector<uint64_t> &host_code, const uint64_t *d_code, uint32_t *state2,
                       uint32_t cap2, uint32_t *tab, uint32_t groups, uint32_t ncurves, uint8_t *inv_fail)
{
    const uint64_t n = host_code.size();
    const uint32_t kMinRun = 16;
    uint64_t seg = 0, i = 0;
    while (i < n) {
        if ((host_code[i] & 0xff) == V_PAIR) {
            uint64_t j = i;
            while (j < n && (host_code[j] & 0xff) == V_PAIR) j++;
            if (j - i >= kMinRun && c->eng->use_pair_kernel) {
                int rc = run_vm2(c, d_code + seg, i - seg, state2, cap2, tab, groups, inv_fail);
                if (rc) return rc;
                {   // chunk-major (group, chunk) items, at most one resident wave per launch
                    S2Mark mark(c, 1);
                    const uint32_t TP = (uint32_t)c->eng->threads_pair, npairs = (uint32_t)(j - i);
                    const uint32_t pgroups = (ncurves + TP - 1) / TP;
                    const uint32_t per = std::min<uint32_t>(pgroups, (uint32_t)(c->num_sms * c->eng->pair_blocks_per_sm));
                    // chunk length: the split of the run into k chunks whose k * pgroups items leave the fewest block slots of
                    // the last launch empty, each chunk paying about two products (accumulator reload, re-join of the halves)
                    uint32_t pchunk = 512;
                    if (const char *e = getenv("ECM_B200_PAIR_CHUNK")) { const int v = atoi(e); if (v >= 16) pchunk = (uint32_t)v; }
                    else if (pgroups > per || npairs > 512) {
                        double best = 0;
                        const uint32_t kmin = (npairs + 1023) / 1024, kmax = std::max<uint32_t>(kmin, npairs / 96);
                        for (uint32_t k = kmin; k <= kmax && k < kmin + 64; k++) {
                            const uint32_t len = (npairs + k - 1) / k;
                            const uint64_t it = (uint64_t)((npairs + len - 1) / len) * pgroups, launches = (it + per - 1) / per;
                            const double eff = (double)it / (double)(launches * per) * (double)len / (double)(len + 2);
                            if (eff > best) { best = eff; pchunk = len; }
                        }
                    }
                    const uint64_t items = (uint64_t)((npairs + pchunk - 1) / pchunk) * pgroups;
                    for (uint64_t it = 0; it < items;) {
                        const uint32_t blocks = (uint32_t)std::min<uint64_t>(per, items - it);
                        c->eng->pair_run(c->stream, blocks, state2, cap2, tab, d_code + i, npairs, ncurves, pchunk, pgroups, it);
                        c->s2_launches++;
                        it += blocks;
                    }
                }
                seg = j;
            }
            i = j;
        } else i++;
    }
    return run_vm2(c, d_code + seg, n - seg, state2, cap2, tab, groups, inv_fail);
}

// ---- stage-2 device buffers: tables, slot file, failure flags and program buffer of one wave.  Allocated on first
// use and kept by the context (a 65 536-curve wave at 415 bits is ~80 GB: allocating it per call cost more than the
// set-up program); re-allocated only when a later call needs more.
static int s2_reserve(ecm_b200_ctx *c, uint32_t entries, uint32_t cap2, size_t code_bytes)
{
    const int nl = c->nl;
    const size_t tab_b = (size_t)entries * nl * 4 * cap2, st_b = (size_t)c->eng->nslot_s2 * nl * 4 * cap2;
    if (tab_b > c->s2_tab_bytes) {
        cudaFree(c->d_tab); c->d_tab = nullptr; c->s2_tab_bytes = 0;
        CU(cudaMalloc(&c->d_tab, tab_b)); c->s2_tab_bytes = tab_b;
    }
    if (st_b > c->s2_state_bytes) {
        cudaFree(c->d_state2); c->d_state2 = nullptr; c->s2_state_bytes = 0;
        CU(cudaMalloc(&c->d_state2, st_b)); c->s2_state_bytes = st_b;
    }
    if (cap2 > c->s2_fail_cap) {
        cudaFree(c->d_wfail); c->d_wfail = nullptr; c->s2_fail_cap = 0;
        CU(cudaMalloc(&c->d_wfail, cap2)); c->s2_fail_cap = cap2;
    }
    if (code_bytes > c->s2_code_bytes) {
        cudaFree(c->d_code); c->d_code = nullptr; c->s2_code_bytes = 0;
        CU(cudaMalloc(&c->d_code, code_bytes + 64)); c->s2_code_bytes = code_bytes;
    }
    if (!c->d_acc) { CU(cudaMalloc(&c->d_acc, (size_t)nl * 4 * c->max_curves)); CU(cudaMalloc(&c->d_fail, c->max_curves)); }
    return ECM_B200_OK;
}

// curves per wave that the device memory allows (tables + slot file per curve), a multiple of the group size
static int s2_wave_capacity(ecm_b200_ctx *c, uint32_t entries, size_t code_bytes, uint32_t *cap_out)
{
    const uint32_t T = c->eng->threads_s2;
    const size_t per_curve = ((size_t)entries + c->eng->nslot_s2) * c->nl * 4 + 1;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    free_b += c->s2_tab_bytes + c->s2_state_bytes + c->s2_fail_cap + c->s2_code_bytes;    // ours to re-use
    const size_t budget = (size_t)((double)free_b * 0.90) - std::min<size_t>(code_bytes + (64u << 20), free_b / 4);
    uint32_t cap2 = (uint32_t)std::min<size_t>((c->count + T - 1) / T * T, budget / per_curve / T * T);
    if (cap2 < T) return fail(ECM_B200_ENOMEM, "not enough device memory for one stage-2 group");
    if (const char *e = getenv("ECM_B200_S2_WAVE")) {        // test hook: force small waves
        const uint32_t w = ((uint32_t)atoi(e) + T - 1) / T * T;
        if (w >= T && w < cap2) cap2 = w;
    }
    *cap_out = cap2;
    return ECM_B200_OK;
}

// stage 2 starts from the point stage 1 left (or from freshly loaded curves: resuming save_b1.txt lines), never from
// the middle of a stage-1 run
static int s2_check_start(ecm_b200_ctx *c, uint64_t b1)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    if (!c->have_curves) return fail(ECM_B200_ESTATE, "no curves loaded");
    if (s1_in_progress(c)) return fail(ECM_B200_ESTATE, "stage 1 is still in progress on this batch");
    const Stage2Params prm = stage2_params(b1);
    if (b1 < 2 || (b1 + prm.D) / (2 * (uint64_t)prm.D) == 0)
        return fail(ECM_B200_EINVAL, "B1 too small for stage 2: the first giant step [A-w]Q would be negative (B1 >= 30 needed)");
    return ECM_B200_OK;
}

// one wave: curves [first, first+n) of the batch -> tables; runs the ecm_stage2_init program
static int s2_wave_init(ecm_b200_ctx *c, uint32_t first, uint32_t n)
{
    const Stage2Program &pg = c->prog2;
    const uint32_t T = c->eng->threads_s2;
    c->s2_first = first; c->s2_n = n; c->s2_groups = (n + T - 1) / T;
    c->eng->s2_setup(c->stream, c->d_state, c->G1, 2 * c->p_slot, 2 * c->p_slot + 1, SP, first, c->count, c->d_state2, c->s2_cap, c->d_tab,
                     pg.lay.qx, pg.lay.qz, c->d_wfail);
    CU(cudaMemcpyAsync(c->d_code, pg.init.data(), pg.init.size() * 8, cudaMemcpyHostToDevice, c->stream));
    return run_program(c, pg.init, c->d_code, c->d_state2, c->s2_cap, c->d_tab, c->s2_groups, n, c->d_wfail);
}

static int s2_wave_run(ecm_b200_ctx *c, const std::vector<uint64_t> &code)
{
    if (code.size() * 8 > c->s2_code_bytes) return fail(ECM_B200_ENOMEM, "stage-2 program larger than its buffer bound");
    // the previous program may still be executing from d_code: the copy is ordered behind it on the same stream
    CU(cudaMemcpyAsync(c->d_code, code.data(), code.size() * 8, cudaMemcpyHostToDevice, c->stream));
    return run_program(c, code, c->d_code, c->d_state2, c->s2_cap, c->d_tab, c->s2_groups, c->s2_n, c->d_wfail);
}

static void s2_wave_collect(ecm_b200_ctx *c)
{
    c->eng->s2_collect(c->stream, c->d_state2, c->s2_cap, c->d_wfail, c->s2_first, c->s2_n, c->count, c->d_acc, c->d_fail);
}

// device buffer bound for the program of one prime range: a range never has more instructions than
// (#primes <= span/6 for span >= 1e5) + 1300 per window shift + ladder/window set-up
static size_t s2_code_bound(const Stage2Params &prm, uint64_t span)
{
    return (size_t)(span / 6 + (span / ((uint64_t)prm.D * prm.U * 2) + 2) * 1300 + 200000) * 8;
}

int ecm_b200_stage2(ecm_b200_ctx *c, uint64_t b1, uint64_t b2)
{
    int rc = s2_check_start(c, b1); if (rc) return rc;
    if (b2 <= b1) return fail(ECM_B200_EINVAL, "B2 must exceed B1 (B2 <= B1 disables stage 2, main.c:548-552)");
    CU(cudaSetDevice(c->device));
    // ---- compile (host): ecm_stage2_init now; the per-range ecm_stage2_pair programs (sieve + PAIR, about
    // 2.5 s per 1e8 primes) are compiled by a background thread while the GPU already executes the init
    // program and the earlier ranges (ecm.c:1424-1476 does this serially on the main thread).
    const uint64_t PRIME_RANGE = 100000000ull;
    std::vector<std::pair<uint64_t, uint64_t>> rng;
    for (uint64_t p = b1; p < b2; p += PRIME_RANGE) rng.push_back({p, std::min(p + PRIME_RANGE, b2)});
    std::mutex mu; std::condition_variable cv; int nready = (int)rng.size();
    std::thread planner;
    if (c->prog2_b1 != b1 || c->prog2_b2 != b2) {
        plan_stage2_init(b1, c->prog2);
        c->prog2.ranges.assign(rng.size(), std::vector<uint64_t>());
        nready = 0;
        Stage2Program *pp = &c->prog2;
        planner = std::thread([pp, rng, &nready, &mu, &cv]() {
            for (size_t r = 0; r < rng.size(); r++) {
                plan_stage2_range(rng[r].first, rng[r].second, *pp, (int)r);
                { std::lock_guard<std::mutex> lk(mu); nready++; }
                cv.notify_all();
            }
        });
        c->prog2_b1 = b1; c->prog2_b2 = b2;
    }
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{planner};
    const Stage2Program &pg = c->prog2;
    const uint32_t T = c->eng->threads_s2;
    size_t code_bytes = pg.init.size() * 8;
    for (const auto &r : rng) code_bytes = std::max(code_bytes, s2_code_bound(pg.prm, r.second - r.first));
    uint32_t cap2 = 0;
    rc = s2_wave_capacity(c, pg.lay.entries, code_bytes, &cap2); if (rc) return rc;
    {   // several waves: make them equal instead of one full wave plus a small remainder
        const uint32_t waves = (c->count + cap2 - 1) / cap2;
        const uint32_t even = ((c->count + waves - 1) / waves + T - 1) / T * T;
        if (even < cap2) cap2 = even;
    }
    rc = s2_reserve(c, pg.lay.entries, cap2, code_bytes); if (rc) return rc;
    c->s2_cap = cap2; c->s2_open = false;
    c->s2_launches = 0; c->s2_waves = 0;
    CU(cudaEventRecord(c->ev0, c->stream));
    for (uint32_t first = 0; first < c->count; first += cap2) {
        rc = s2_wave_init(c, first, std::min<uint32_t>(cap2, c->count - first)); if (rc) return rc;
        for (size_t ri = 0; ri < rng.size(); ri++) {
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return nready > (int)ri; }); }
            rc = s2_wave_run(c, pg.ranges[ri]); if (rc) return rc;
        }
        s2_wave_collect(c);
        c->s2_waves++;
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&c->s2_ms, c->ev0, c->ev1);
    c->last_ms = c->s2_ms; c->last_launches = c->s2_launches;
    s2_trace_report(c, "stage2");
    c->stage2_done = true;
    return ECM_B200_OK;
}

// ---- stage 2 at the reference's own granularity (ecm.c:67-72) -------------------------------------------------------
// ecm_stage2_init: baby steps Pb[], their batch inversion, Pd = [w]Q for the whole batch, which must fit the device in
// one wave (the tables stay resident for the range calls that follow).  *found_inv = 1 when an inversion met a
// non-invertible element on some curve (the reference's return value foundDuringInv).
int ecm_b200_stage2_init(ecm_b200_ctx *c, uint64_t b1, int *found_inv)
{
    int rc = s2_check_start(c, b1); if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if (c->prog2_b1 != b1 || c->prog2_b2 != 0) { plan_stage2_init(b1, c->prog2); c->prog2_b1 = b1; c->prog2_b2 = 0; }
    const size_t code_bytes = std::max(c->prog2.init.size() * 8, s2_code_bound(c->prog2.prm, 100000000ull));
    uint32_t cap2 = 0;
    rc = s2_wave_capacity(c, c->prog2.lay.entries, code_bytes, &cap2); if (rc) return rc;
    if (cap2 < c->count)
        return fail(ECM_B200_ENOMEM, "batch does not fit one stage-2 wave on this device: use ecm_b200_stage2 (it runs waves) or a smaller batch");
    rc = s2_reserve(c, c->prog2.lay.entries, cap2, code_bytes); if (rc) return rc;
    c->s2_cap = cap2; c->s2_launches = 0; c->s2_waves = 1; c->s2_ms = 0;
    CU(cudaEventRecord(c->ev0, c->stream));
    rc = s2_wave_init(c, 0, c->count); if (rc) return rc;
    s2_wave_collect(c);
    CU(cudaEventRecord(c->ev1, c->stream));
    std::vector<uint8_t> fl(c->count);
    CU(cudaMemcpyAsync(fl.data(), c->d_fail, c->count, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); c->s2_ms += ms; c->last_ms = ms; c->last_launches = c->s2_launches;
    s2_trace_report(c, "stage2_init");
    if (found_inv) { *found_inv = 0; for (uint8_t f : fl) if (f) { *found_inv = 1; break; } }
    c->s2_open = true; c->stage2_done = true;
    return ECM_B200_OK;
}

// ecm_stage2_pair(steps, pm_v, pm_u, ...) with work->amin = amin (ecm.c:2342-2540): giant-step window from [2*amin*w]Q,
// then the caller's pairmap -- (0,0) entries slide the window by U.  The pairmap may come from the reference's pair()
// or from ecm_b200_pair; it is validated (EINVAL) before anything runs.
int ecm_b200_stage2_range(ecm_b200_ctx *c, uint32_t amin, const uint32_t *pm_v, const uint32_t *pm_u, uint32_t steps)
{
    if (!c || (steps && (!pm_v || !pm_u))) return fail(ECM_B200_EINVAL, "null argument");
    if (!c->s2_open) return fail(ECM_B200_ESTATE, "ecm_b200_stage2_init has not been run on this batch");
    if (!stage2_pairmap_valid(c->prog2.prm, amin, pm_v, pm_u, steps))
        return fail(ECM_B200_EINVAL, "pairmap leaves the stage-2 tables (window index outside [amin, amin+2L), unmapped baby step, or amin = 0)");
    CU(cudaSetDevice(c->device));
    c->prog2.ranges.assign(1, std::vector<uint64_t>());
    plan_stage2_pairmap(amin, pm_v, pm_u, steps, c->prog2, 0);
    c->prog2_b2 = 0;                                      // the cached programs no longer describe a (B1,B2) run
    const uint32_t before = c->s2_launches;
    CU(cudaEventRecord(c->ev0, c->stream));
    int rc = s2_wave_run(c, c->prog2.ranges[0]); if (rc) return rc;
    s2_wave_collect(c);
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); c->s2_ms += ms; c->last_ms = ms; c->last_launches = c->s2_launches - before;
    s2_trace_report(c, "stage2_range");
    return ECM_B200_OK;
}

int ecm_b200_read_stage2(ecm_b200_ctx *c, uint32_t *acc, uint8_t *factor_flag, uint32_t *gcd_out, uint8_t *inv_fail)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    if (!c->stage2_done) return fail(ECM_B200_ESTATE, "stage 2 has not been run");
    CU(cudaSetDevice(c->device));
    const size_t words = (size_t)c->nl * c->count;
    uint32_t *da = c->d_io, *dg = c->d_io + words;
    const bool want_flag = factor_flag || gcd_out;
    // d_acc is laid out like a one-slot state with cap = count
    c->eng->read_point(c->stream, c->d_acc, Geom{c->count, c->count, 1}, c->count, 0, 0, nullptr, da, want_flag ? c->d_flags : nullptr,
                       gcd_out ? dg : nullptr, c->d_chk);
    CU(cudaGetLastError());
    if (acc) CU(cudaMemcpyAsync(acc, da, words * 4, cudaMemcpyDeviceToHost, c->stream));
    if (factor_flag) CU(cudaMemcpyAsync(factor_flag, c->d_flags, c->count, cudaMemcpyDeviceToHost, c->stream));
    if (gcd_out) CU(cudaMemcpyAsync(gcd_out, dg, words * 4, cudaMemcpyDeviceToHost, c->stream));
    if (inv_fail) CU(cudaMemcpyAsync(inv_fail, c->d_fail, c->count, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return ECM_B200_OK;
}

// counters of the compiled stage-2 program (what the reference prints, ecm.c:1481-1483)
int ecm_b200_stage2_counters(const ecm_b200_ctx *c, uint64_t *ptadds, uint64_t *numinv, uint64_t *paired, uint64_t *steps)
{
    if (!c) return fail(ECM_B200_EINVAL, "null context");
    if (ptadds) *ptadds = c->prog2.ptadds;
    if (numinv) *numinv = c->prog2.numinv;
    if (paired) *paired = c->prog2.paired;
    if (steps) *steps = c->prog2.pairmap_steps;
    return ECM_B200_OK;
}

int ecm_b200_fieldop(ecm_b200_ctx *c, int op, uint32_t count, const uint32_t *a, const uint32_t *b, uint32_t *r, int repeat)
{
    if (!c || !a || !b || !r || op < 0 || op > 3 || repeat < 1) return fail(ECM_B200_EINVAL, "bad argument");
    if (count > c->max_curves) return fail(ECM_B200_EINVAL, "count exceeds the context's capacity");
    CU(cudaSetDevice(c->device));
    const size_t words = (size_t)c->nl * count;
    CU(cudaMemcpyAsync(c->d_io, a, words * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_io + words, b, words * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev0, c->stream));
    c->eng->fieldop(c->stream, op, count, c->d_io, c->d_io + words, c->d_io + 2 * words, repeat);
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(r, c->d_io + 2 * words, words * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); c->last_ms = ms; c->last_launches = 1;
    return ECM_B200_OK;
}

// ---- planners --------------------------------------------------------------------------------
uint64_t ecm_b200_plan_stage1(uint64_t b1, uint8_t *ops, uint64_t cap, uint64_t *counts)
{
    Stage1Plan p;
    plan_stage1(b1, p);
    if (ops && cap >= p.n_ops) memcpy(ops, p.ops.data(), p.n_ops);
    if (counts) { counts[0] = p.ptadds; counts[1] = p.ptdups; }
    return p.n_ops;
}

uint32_t ecm_b200_pair(uint64_t lo, uint64_t hi, uint32_t D, uint32_t U, uint32_t *pairmap_v, uint32_t *pairmap_u,
                       uint32_t cap, uint32_t *amin_final, uint32_t *npairs)
{
    Stage2Params p; p.D = D; p.U = U; p.L = 2 * U; p.R = 0;
    std::vector<uint32_t> v, u;
    uint32_t steps = pair_plan(lo, hi, p, v, u, amin_final, npairs);
    if (pairmap_v && pairmap_u && cap >= steps) { memcpy(pairmap_v, v.data(), steps * 4ull); memcpy(pairmap_u, u.data(), steps * 4ull); }
    return steps;
}

uint64_t ecm_b200_plan_stage2(uint64_t b1, uint64_t b2, uint64_t *counts)
{
    Stage2Program pg;
    plan_stage2_init(b1, pg);
    for (uint64_t p = b1; p < b2; p += 100000000ull) plan_stage2_range(p, std::min<uint64_t>(p + 100000000ull, b2), pg);
    uint64_t n = pg.init.size();
    for (const auto &r : pg.ranges) n += r.size();
    if (counts) { counts[0] = pg.ptadds; counts[1] = pg.numinv; counts[2] = pg.paired; counts[3] = pg.pairmap_steps; counts[4] = pg.last_amin; counts[5] = pg.lay.entries; }
    return n;
}

uint64_t ecm_b200_stage2_program(uint64_t b1, uint64_t b2, int which, uint64_t *out, uint64_t cap, uint32_t *layout)
{
    Stage2Program pg;
    plan_stage2_init(b1, pg);
    const std::vector<uint64_t> *src = &pg.init;
    if (which >= 0) {
        uint64_t p = b1 + (uint64_t)which * 100000000ull;
        if (p >= b2) return 0;
        plan_stage2_range(p, std::min<uint64_t>(p + 100000000ull, b2), pg);
        src = &pg.ranges.back();
    }
    if (out && cap >= src->size()) memcpy(out, src->data(), src->size() * 8);
    if (layout) {
        const Stage2Layout &L = pg.lay;
        const uint32_t v[12] = {L.npb, L.pbx, L.pbz, L.pba, L.pax, L.paz, L.pai, L.paa, L.qx, L.qz, L.pdx, L.pdz};
        memcpy(layout, v, sizeof v);
        layout[12] = L.entries;
    }
    return src->size();
}

uint64_t ecm_b200_stage2_pairmap_program(uint64_t b1, uint32_t amin, const uint32_t *pm_v, const uint32_t *pm_u, uint32_t steps,
                                         uint64_t *out, uint64_t cap)
{
    Stage2Program pg;
    pg.prm = stage2_params(b1);
    pg.lay = stage2_layout(pg.prm);
    if ((steps && (!pm_v || !pm_u)) || !stage2_pairmap_valid(pg.prm, amin, pm_v, pm_u, steps)) return 0;
    plan_stage2_pairmap(amin, pm_v, pm_u, steps, pg);
    const std::vector<uint64_t> &src = pg.ranges.back();
    if (out && cap >= src.size()) memcpy(out, src.data(), src.size() * 8);
    return src.size();
}

int ecm_b200_rv_program(int macro_op, uint32_t *phases, int cap)
{
    static const uint32_t prog[8][RV_MAXPROG] = { RV_PROGRAMS };
    if (macro_op < 0 || macro_op > 7 || !phases) return 0;
    int n = 0;
    while ((prog[macro_op][n] & 15u) != PH_END) n++;
    for (int k = 0; k < n && k < cap; k++) phases[k] = prog[macro_op][k];
    return n;
}

void ecm_b200_stage2_params(uint64_t b1, uint32_t *D, uint32_t *U, uint32_t *L, uint32_t *R)
{
    Stage2Params p = stage2_params(b1);
    if (D) *D = p.D; if (U) *U = p.U; if (L) *L = p.L; if (R) *R = p.R;
}


}  // extern "C"
