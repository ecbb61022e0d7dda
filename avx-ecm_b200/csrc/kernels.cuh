// kernels.cuh -- __global__ entry points, templated on the limb count NL.
// Device state of a batch: curves are cut into groups of T (= threads per block); group g owns
//   state[((g*NSLOT + slot)*NL + limb)*STRIDE + lane]        (Geom below)
// limb-major with the curve (lane) fastest: a warp's 32 curves read 128 contiguous bytes per limb --
// the coalesced analogue of the reference's bignum.data[lane + word*VECLEN], vec_common.c:32-54.
#pragma once
#include "geom.hpp"
#include "vm.cuh"
#include "modinv.cuh"
#include "modinv_fast.hpp"

// ECM_FAST_INV = 1 (default): word-batched inverse / gcd (modinv_fast.hpp); 0: the bit-serial routine of modinv.cuh
#ifndef ECM_FAST_INV
#define ECM_FAST_INV 1
#endif

namespace ecmb200 {
inline namespace ECM_VNS {

constexpr int kSmemBudget = 227 * 1024;
constexpr int NSMEM_S1 = 5;      // stage-1 slots kept in shared memory (s1,d1,s2,d2,sp); the 8 point slots stay in global


// Stage-2 tables: entry e of curve c lives at tab[((e*nwg + c/32)*NL + limb)*32 + c%32]: the NL limbs of a
// warp's 32 curves form one contiguous NL*128-byte block, so an operand costs one 64-bit address
// computation and NL loads at immediate offsets, each a single 128-byte line.
__device__ __forceinline__ size_t tab_base(uint32_t entry, uint32_t nwg, uint32_t curve, int NL)
{
    return (((size_t)entry * nwg + (curve >> 5)) * NL) * 32 + (curve & 31);
}
template <int NL>
__device__ __forceinline__ void tload(uint32_t (&r)[NL], const uint32_t *tab, size_t base)
{
    const uint32_t *p = tab + base;
#pragma unroll
    for (int k = 0; k < NL; k++) r[k] = p[k * 32];
}
// hint the NL lines of a table operand into L2/L1 ahead of its use (one 128-byte line per limb and warp)
template <int NL>
__device__ __forceinline__ void tprefetch(const uint32_t *tab, size_t base)
{
    const uint32_t *p = tab + base;
#pragma unroll
    for (int k = 0; k < NL; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + k * 32));
}
// The slot machine executes `ldg; mul; stg` sequences of the batch inversions with the table in HBM and only 8-12 warps
// per SM: ncu showed 1.7 "long scoreboard" stall cycles per issued instruction (55 % fmaheavy).  Looking a few
// instructions ahead and prefetching the operands of the next table read hides that latency behind the current product.
template <int NL>
__device__ __forceinline__ void vm2_lookahead(const uint64_t *code, uint64_t i, uint64_t end, const uint32_t *tab, uint32_t nwg, uint32_t lane)
{
    constexpr int AHEAD = 3;
    if (i + AHEAD < end) {
        const uint64_t ins = __ldg(code + i + AHEAD);
        const uint32_t op = (uint32_t)ins & 0xffu, imm = (uint32_t)(ins >> 32);
        if (op == 6u /* V2_LDG */) tprefetch<NL>(tab, tab_base(imm, nwg, lane, NL));
        else if (op == 10u /* V2_PAIR */) {
            tprefetch<NL>(tab, tab_base(imm & 0xffffu, nwg, lane, NL));
            tprefetch<NL>(tab, tab_base(imm >> 16, nwg, lane, NL));
        }
    }
}
template <int NL>
__device__ __forceinline__ void tstore(uint32_t *tab, size_t base, const uint32_t (&r)[NL])
{
    uint32_t *p = tab + base;
#pragma unroll
    for (int k = 0; k < NL; k++) p[k * 32] = r[k];
}

// ---- stage 1: interpret macro-ops [chunk*chunk_len, ...) for one group of blockDim.x curves ------
template <int NL>
struct S1Cfg {
    // wide moduli stream their operands from shared memory and need two staging slots for point operands
    static constexpr int nsmem = (NL > 32) ? NSMEM_S1 + 2 : NSMEM_S1;
    static constexpr int per_thread = nsmem * NL * 4;
    static constexpr int fit = (kSmemBudget / per_thread) / 32 * 32;
    // the kernels are bound by dependent-issue latency, so the cap is what the register count of the unrolled
    // dual product allows without spills (78 / 96 / 114 registers at 10 / 13 / 16 limbs); the host picks the
    // block size per batch (ecm_gpu.cu)
    static constexpr int cap = (NL <= 10) ? 768 : (NL <= 13) ? 640 : (NL <= 16) ? 512 : 768;
    static constexpr int STRIDE = fit > cap ? cap : fit;          // max threads per block = lane stride
    static constexpr int smem = per_thread * STRIDE;
};

// MAXT = the largest block this instance is launched with = lane stride of its state and shared-memory layout.
// Two instances per limb count: blocks up to 384 threads keep the layout, shared-memory footprint and register
// allocation the 65 536-curve configurations were tuned with; larger blocks (bigger batches, see the block-size
// model in ecm_gpu.cu) use the wider one (measured at 13 limbs: a 640-lane layout costs the 384-thread
// configuration 3 %, 7.64 -> 7.41 Tprod/s, while 640 threads on 94 720 curves reach 8.07).
template <int NL> struct S1Small { static constexpr int MAXT = S1Cfg<NL>::STRIDE < 384 ? S1Cfg<NL>::STRIDE : 384; };

template <int NL, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
k_stage1(const ModParams<NL> P, uint32_t *__restrict__ state, const uint8_t *__restrict__ ops,
         uint64_t nops, uint32_t chunk_len, uint32_t groups, uint64_t item0)
{
    constexpr int STRIDE = MAXT;
    extern __shared__ uint32_t smem[];
    const uint64_t item = item0 + blockIdx.x;
    const uint32_t g = (uint32_t)(item % groups);
    const uint64_t chunk = item / groups;
    HybridSlots<NL, STRIDE> S{state + (size_t)g * (NSLOT_S1 * NL * STRIDE) + threadIdx.x, smem + threadIdx.x};

    uint32_t r[NL];
#pragma unroll 1
    for (uint32_t s = 8; s < NSLOT_S1; s++) {          // scratch slots: global image -> shared
#pragma unroll
        for (int k = 0; k < NL; k++) r[k] = S.gl[(s * NL + k) * STRIDE];
        S.store(s, r);
    }

    uint64_t i = chunk * chunk_len;
    const uint64_t end = (i + chunk_len < nops) ? i + chunk_len : nops;
    const uint32_t *ops32 = reinterpret_cast<const uint32_t *>(ops);
    uint32_t cur = 0;
#pragma unroll 1
    for (; i < end; i++) {
        if ((i & 3) == 0) cur = __ldg(ops32 + (i >> 2));
        const uint32_t byte = cur & 0xffu;
        cur >>= 8;
        const uint32_t type = byte & 7u;
        const uint32_t permbits = c_perm[byte >> 3];
#pragma unroll 1
        for (int k = 0;; k++) {
            const uint32_t u = c_prog_s1[type][k];
            if ((u & 15u) == U_END) break;
            exec_uop<NL>(S, u, permbits, P);
        }
    }
#pragma unroll 1
    for (uint32_t s = 8; s < NSLOT_S1; s++) {
        S.load(r, s);
#pragma unroll
        for (int k = 0; k < NL; k++) S.gl[(s * NL + k) * STRIDE] = r[k];
    }
}

// ---- out-of-line helpers for the set-up / read-out kernels (speed is irrelevant there; keeping
// one copy of each routine keeps compile time and code size down).  Pg points to a copy of the
// parameters in global memory.
template <int NL>
__device__ __noinline__ void nm_mul(uint32_t *r, const uint32_t *a, const uint32_t *b, const ModParams<NL> *Pg)
{
    uint32_t x[NL], y[NL], z[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) { x[k] = a[k]; y[k] = b[k]; }
    mont_mul<NL>(z, x, y, *Pg);
#pragma unroll
    for (int k = 0; k < NL; k++) r[k] = z[k];
}
template <int NL>
__device__ __noinline__ void nm_addsub(uint32_t *r, const uint32_t *a, const uint32_t *b, bool sub, const ModParams<NL> *Pg)
{
    uint32_t x[NL], y[NL], z[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) { x[k] = a[k]; y[k] = b[k]; }
    if (sub) mod_sub<NL>(z, x, y, *Pg); else mod_add<NL>(z, x, y, *Pg);
#pragma unroll
    for (int k = 0; k < NL; k++) r[k] = z[k];
}
template <int NL>
__device__ __noinline__ bool nm_inverse(uint32_t *inv, uint32_t *g, const uint32_t *x, const ModParams<NL> *Pg)
{
#if ECM_FAST_INV
    return fastinv::mod_inverse<NL, true>(inv, g, x, Pg->n);
#else
    uint32_t a[NL], i[NL], gg[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) a[k] = x[k];
    const bool ok = mod_inverse<NL, true>(i, gg, a, Pg->n);
#pragma unroll
    for (int k = 0; k < NL; k++) { inv[k] = i[k]; g[k] = gg[k]; }
    return ok;
#endif
}
// g = gcd(x, m) for an odd m of NL limbs in global memory (x need not be below m)
template <int NL>
__device__ __noinline__ void nm_gcd(uint32_t *g, const uint32_t *x, const uint32_t *m)
{
#if ECM_FAST_INV
    fastinv::mod_inverse<NL, false>(nullptr, g, x, m);
#else
    uint32_t a[NL], i[NL], gg[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) a[k] = x[k];
    mod_inverse<NL, false>(i, gg, a, m);
#pragma unroll
    for (int k = 0; k < NL; k++) g[k] = gg[k];
#endif
}

// ---- curve set-up -----------------------------------------------------------------------------
// host-built curves: plain x = X/Z and s = (A+2)/4  ->  Montgomery form, Z = 1; P sits in point slot 0
template <int NL>
__global__ void k_load_curves(const ModParams<NL> *Pg, uint32_t *state, Geom G, uint32_t lanes, uint32_t count,
                              const uint32_t *x, const uint32_t *s)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= lanes) return;
    const uint32_t src = c < count ? c : count - 1;          // padding lanes replicate the last curve
    uint32_t a[NL], t[NL];
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
        const uint32_t *in = which ? s : x;
        for (int k = 0; k < NL; k++) a[k] = in[(size_t)k * count + src];
        nm_mul<NL>(t, a, Pg->r2, Pg);
        for (int k = 0; k < NL; k++) state[G.idx(c, which ? SP : 0, k, NL)] = t[k];
    }
    for (int k = 0; k < NL; k++) state[G.idx(c, 1, k, NL)] = Pg->one[k];
}

// Suyama "param 0" curve from u = sigma^2-5, v = 4*sigma (both already reduced mod N, 4 limbs each,
// [limb][curve]); build_one_curve, ecm.c:1587-1772:
//   X = u^3 / v^3,  Z = 1,  s = (v-u)^3 (3u+v) / (16 u^3 v)
// One shared inversion of (v^3 * 16u^3v) replaces the reference's two mpz_invert calls; the
// quotients are the same residues.  ok[c] = 0 when that product is not invertible mod N; then the
// reference's two mpz_invert calls (ecm.c:1745,1759) each leave their output operand untouched
// when they fail, i.e. s = num * (16 u^3 v invertible ? its inverse : 16 u^3) and
// X = u^3 * (v^3 invertible ? its inverse : num).  16u^3v is invertible iff the shared product is,
// so on failure s uses 16u^3 and only v^3 is tried again on its own.  (Common for special-form
// inputs, where the arithmetic modulus 2^k+-1 keeps its small algebraic factors.)
template <int NL>
__global__ void k_build_curves(const ModParams<NL> *Pg, uint32_t *state, Geom G, uint32_t lanes, uint32_t count,
                               const uint32_t *uv, uint8_t *ok)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= lanes) return;
    const uint32_t src = c < count ? c : count - 1;
    uint32_t u[NL], v[NL], t1[NL], t2[NL], u3[NL], v3[NL], num[NL], den[NL], inv[NL], g[NL];
    for (int k = 0; k < NL; k++) {
        u[k] = (k < 4) ? uv[(size_t)k * count + src] : 0;
        v[k] = (k < 4) ? uv[(size_t)(4 + k) * count + src] : 0;
    }
    nm_mul<NL>(u, u, Pg->r2, Pg);                    // Montgomery form
    nm_mul<NL>(v, v, Pg->r2, Pg);
    nm_mul<NL>(t1, u, u, Pg); nm_mul<NL>(u3, t1, u, Pg);          // u^3
    nm_mul<NL>(t1, v, v, Pg); nm_mul<NL>(v3, t1, v, Pg);          // v^3
    nm_addsub<NL>(t1, v, u, true, Pg);                            // v-u
    nm_mul<NL>(t2, t1, t1, Pg); nm_mul<NL>(num, t2, t1, Pg);      // (v-u)^3
    nm_addsub<NL>(t1, u, u, false, Pg); nm_addsub<NL>(t1, t1, u, false, Pg); nm_addsub<NL>(t1, t1, v, false, Pg);   // 3u+v
    nm_mul<NL>(num, num, t1, Pg);                                 // (v-u)^3 (3u+v)
    nm_mul<NL>(den, u3, v, Pg);                                   // u^3 v
    for (int k = 0; k < 4; k++) nm_addsub<NL>(den, den, den, false, Pg);   // 16 u^3 v
    nm_mul<NL>(t1, den, v3, Pg);                                  // den * v^3   (Montgomery form)
    const bool good = nm_inverse<NL>(inv, g, t1, Pg);             // (den v^3 R)^-1
    if (good) {
        nm_mul<NL>(inv, inv, Pg->r3, Pg);                         // -> (den v^3)^-1 in Montgomery form
        nm_mul<NL>(t1, inv, den, Pg);                             // 1/v^3
        nm_mul<NL>(t2, inv, v3, Pg);                              // 1/den
    } else {
        for (int k = 0; k < NL; k++) t2[k] = u3[k];
        for (int k = 0; k < 4; k++) nm_addsub<NL>(t2, t2, t2, false, Pg);      // 16 u^3 stands in for 1/den
        if (nm_inverse<NL>(inv, g, v3, Pg)) nm_mul<NL>(t1, inv, Pg->r3, Pg);   // 1/v^3 still exists
        else for (int k = 0; k < NL; k++) t1[k] = num[k];                       // num stands in for 1/v^3
    }
    nm_mul<NL>(t1, t1, u3, Pg);                                   // X = u^3/v^3
    nm_mul<NL>(t2, t2, num, Pg);                                  // s
    for (int k = 0; k < NL; k++) {
        state[G.idx(c, 0, k, NL)] = t1[k];
        state[G.idx(c, 1, k, NL)] = Pg->one[k];
        state[G.idx(c, SP, k, NL)] = t2[k];
    }
    if (c < count) ok[c] = good ? 1 : 0;
}

// ---- results ---------------------------------------------------------------------------------
// X*1, Z*1 leave Montgomery form (ecm.c:1327-1331); gcd(Z,N) is the stage-1 factor test
// (ecm.c:1335-1342, check_factor ecm.c:2542-2557: 1 < g < N).
template <int NL>
__global__ void k_read_point(const ModParams<NL> *Pg, const uint32_t *state, Geom G, uint32_t count,
                             uint32_t xslot, uint32_t zslot, uint32_t *x_out, uint32_t *z_out,
                             uint8_t *flag, uint32_t *g_out, const uint32_t *chk)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    uint32_t a[NL], one[NL], r[NL], g[NL];
    for (int k = 0; k < NL; k++) one[k] = (k == 0);
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
        uint32_t *out = which ? z_out : x_out;
        if (!out && !(which && flag)) continue;
        for (int k = 0; k < NL; k++) a[k] = state[G.idx(c, which ? zslot : xslot, k, NL)];
        if (out) {
            nm_mul<NL>(r, a, one, Pg);
            for (int k = 0; k < NL; k++) out[(size_t)k * count + c] = r[k];
        }
    }
    if (flag) {                                   // a = Z (Montgomery form; gcd(Z*R,N) = gcd(Z,N))
        // chk = the input N: the arithmetic modulus itself, or for special-form inputs the cofactor
        // of 2^k+-c that is being factored (ecm.c:1108-1119) -- R stays coprime to it
        nm_gcd<NL>(g, a, chk);
        bool is_one = (g[0] == 1), is_n = true;
        for (int k = 0; k < NL; k++) { if (k && g[k]) is_one = false; if (g[k] != chk[k]) is_n = false; }
        const bool found = !is_one && !is_n;
        flag[c] = found ? 1 : 0;
        if (g_out) for (int k = 0; k < NL; k++) g_out[(size_t)k * count + c] = found ? g[k] : 0;
    }
}

// ---- field-op hook ----------------------------------------------------------------------------
// op 0 mul, 1 sqr, 2 add, 3 sub; `repeat` chained applications (a <- op(a,b)) so that the same
// kernel serves as the modmul throughput micro-benchmark.
template <int NL>
__global__ void k_fieldop(const ModParams<NL> P, const ModParams<NL> *Pg, int op, uint32_t count, const uint32_t *a_in,
                          const uint32_t *b_in, uint32_t *r_out, int repeat)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    uint32_t a[NL], b[NL], r[NL], one[NL];
    for (int k = 0; k < NL; k++) { a[k] = a_in[(size_t)k * count + c]; b[k] = b_in[(size_t)k * count + c]; one[k] = (k == 0); }
    nm_mul<NL>(a, a, Pg->r2, Pg);
    nm_mul<NL>(b, b, Pg->r2, Pg);
#pragma unroll 1
    for (int it = 0; it < repeat; it++) {
        if (op == 0) mont_mul<NL>(r, a, b, P);
        else if (op == 1) mont_sqr<NL>(r, a, P);
        else if (op == 2) mod_add<NL>(r, a, b, P);
        else mod_sub<NL>(r, a, b, P);
#pragma unroll
        for (int k = 0; k < NL; k++) a[k] = r[k];
    }
    nm_mul<NL>(r, a, one, Pg);
    for (int k = 0; k < NL; k++) r_out[(size_t)k * count + c] = r[k];
}

// =================================================================================================
// stage 2: the second field-op machine.  Same execution model as stage 1 (one thread = one curve,
// slot file in shared memory, one op stream for the whole batch) with 64-bit instructions
// compiled by plan2.cpp, global-memory point tables, a modular inverse and the fused pair step.
// Table entry e of curve c: see tab_base().
// =================================================================================================
enum : uint32_t { V2_MUL = 0, V2_SQR = 1, V2_ADD = 2, V2_SUB = 3, V2_ADDSUB = 4, V2_COPY = 5, V2_LDG = 6, V2_STG = 7,
                  V2_INV = 8, V2_ONE = 9, V2_PAIR = 10, V2_NOP = 11, V2_MUL2 = 12 };
enum : uint32_t { V2_ACC = 6, V2_SP = 11, NSLOT_S2 = 14, NGLOBAL_S2 = 7 };   // slots 0..6 (points, acc) global, 7..13 shared

// d = 1/x (both Montgomery form).  Non-invertible x: reproduce what lane 0 of the reference's
// vectors does (ecm.c:1925-1949 with insert_mpz_to_vec main.c:117-138): the accumulator becomes the
// raw gcd and the "inverse" is the raw, un-inverted product -- as Montgomery-domain values of the
// reference (R_ref = 2^MAXBITS) these are g/R_ref and x/R_ref.
// Successful inversions reproduce one more accident of the reference, "stale high words": the vector lane that receives
// the new value v = x^-1 * R_ref mod N still holds the operand x taken out of Montgomery form (ecm.c:1903-1911), and
// insert_mpz_to_vec (main.c:117-138) only writes the 52-bit words of v up to its top non-zero one -- words of x above
// the length of v survive.  The lane then holds v + (x with its low 52k bits cleared), k = 52-bit words of v, and the
// reference computes on with that residue.  With b bits in the top word of N this hits about 2^-b of all inversions:
// never in practice for most inputs, constantly when N is one bit longer than a multiple of 52 (test.csh lines 2 and
// 25: 729- and 417-bit inputs, golden cases csh_line02 / csh_line25).  Pg->rref = R_ref mod N as a plain integer
// (1 for special-form inputs, whose residues are plain), Pg->rrefinv = R_ref^-1 in Montgomery form.
template <int NL>
__device__ __noinline__ void stale_high_words(uint32_t *t, const uint32_t *xmont, const ModParams<NL> *Pg)
{
    uint32_t v[NL], xp[NL], one[NL];
    for (int k = 0; k < NL; k++) one[k] = (k == 0);
    nm_mul<NL>(v, t, Pg->rref, Pg);                      // x^-1 R * R_ref / R = v (plain, canonical)
    nm_mul<NL>(xp, xmont, one, Pg);                      // x (plain)
    int bl = 0;
    for (int k = NL - 1; k >= 0; k--) if (v[k]) { bl = 32 * k + 32 - __clz(v[k]); break; }
    const int cut = 52 * ((bl + 51) / 52);               // bits of v's 52-bit words
    bool any = false;
    for (int k = 0; k < NL; k++) {
        uint32_t w = xp[k];
        if (32 * k + 32 <= cut) w = 0;
        else if (32 * k < cut) w &= ~0u << (cut - 32 * k);
        xp[k] = w;
        any |= (w != 0);
    }
    if (!any) return;                                    // the usual case: nothing survives
    nm_addsub<NL>(v, v, xp, false, Pg);                  // (v + stale) mod N; v + stale < 2N
    nm_mul<NL>(v, v, Pg->r2, Pg);
    nm_mul<NL>(t, v, Pg->rrefinv, Pg);                   // back: (v + stale) / R_ref in Montgomery form
}

template <int NL, int STRIDE>
__device__ __noinline__ void vm2_inverse(uint32_t *dptr, const uint32_t *xptr, uint32_t *accptr, const ModParams<NL> *Pg, uint8_t *fail_flag)
{
    uint32_t a[NL], inv[NL], g[NL], t[NL];
    for (int k = 0; k < NL; k++) a[k] = xptr[k * STRIDE];
    const bool ok = nm_inverse<NL>(inv, g, a, Pg);
    if (ok) {
        nm_mul<NL>(t, inv, Pg->r3, Pg);                   // (xR)^-1 * R^3 * R^-1 = x^-1 R
        stale_high_words<NL>(t, a, Pg);
    } else {
        *fail_flag = 1;
        nm_mul<NL>(t, g, Pg->r2, Pg);                     // g in Montgomery form
        nm_mul<NL>(t, t, Pg->rrefinv, Pg);                // g / R_ref
        for (int k = 0; k < NL; k++) accptr[k * STRIDE] = t[k];
        nm_mul<NL>(t, a, Pg->rrefinv, Pg);                // x / R_ref
    }
    for (int k = 0; k < NL; k++) dptr[k * STRIDE] = t[k];
}

// Block geometry of the stage-2 machine.  The slot file is hybrid like stage 1's (the three work points and the
// accumulator stay in the L2-resident state, 7 scratch slots in shared memory), which buys resident warps: 128 -> 256
// threads per SM at 1024 bits (+21 %), 256 -> 384 at 13 limbs; wider moduli keep all 14 slots in shared.
template <int NL>
struct S2Cfg {
    // 13 limbs, stage 2 at B1=1e6/B2=1e8 with 65 536 curves: with one chunk per segment the 171 groups of 384 left the
    // second launch of every segment 85 % empty and the hybrid file lost (12.6 s against 11.9 s); with run_vm2() cutting
    // segments into chunks that fill the launches it wins (slot-machine segments 2.95 -> 2.69 s, stage 2 11.43 -> 11.17 s)
#ifndef ECM_S2_HYBRID_SMALL
#define ECM_S2_HYBRID_SMALL 1
#endif
    static constexpr bool HYBRID = (NL >= 20 && NL <= 32) || (ECM_S2_HYBRID_SMALL && NL <= 16);
    static constexpr int nsmem = HYBRID ? (NSLOT_S2 - NGLOBAL_S2) : NSLOT_S2;
    static constexpr int per_thread = nsmem * NL * 4;
    static constexpr int fit = (kSmemBudget / per_thread) / 32 * 32;
    static constexpr int even = fit >= 128 ? fit / 128 * 128 : fit;
    static constexpr int cap = (NL <= 16) ? 384 : 256;           // register budget of the dual-product path
    static constexpr int THREADS = even > cap ? cap : (even < 32 ? 32 : even);
    static constexpr int smem = per_thread * THREADS;
};

template <int NL>
__global__ void __launch_bounds__(S2Cfg<NL>::THREADS, 1)
k_vm2(const ModParams<NL> P, const ModParams<NL> *Pg, uint32_t *__restrict__ state2, uint32_t cap, uint32_t *__restrict__ tab,
      const uint64_t *__restrict__ code, uint64_t ncode, uint32_t chunk_len, uint32_t groups, uint64_t item0,
      uint8_t *__restrict__ inv_fail)
{
    constexpr int THREADS = S2Cfg<NL>::THREADS;
    constexpr int NG = S2Cfg<NL>::HYBRID ? NGLOBAL_S2 : 0;          // slots served from the global state
    extern __shared__ uint32_t smem[];
    const uint64_t item = item0 + blockIdx.x;
    const uint32_t g = (uint32_t)(item % groups);
    const uint64_t chunk = item / groups;
    const uint32_t curve = g * THREADS + threadIdx.x;
    const uint32_t nwg = cap >> 5;
    // state2 is group-blocked: [group][slot][limb][THREADS]
    HybridSlots<NL, THREADS, NG, NSLOT_S2 - NG> S{state2 + (size_t)g * (NSLOT_S2 * NL * THREADS) + threadIdx.x, smem + threadIdx.x};
    uint32_t a[NL], b[NL], r[NL];
#pragma unroll 1
    for (uint32_t s = NG; s < NSLOT_S2; s++) {
#pragma unroll
        for (int k = 0; k < NL; k++) r[k] = S.gl[(s * NL + k) * THREADS];
        S.store(s, r);
    }

    uint64_t i = chunk * chunk_len;
    const uint64_t end = (i + chunk_len < ncode) ? i + chunk_len : ncode;
#pragma unroll 1
    for (; i < end; i++) {
        const uint64_t ins = __ldg(code + i);
        vm2_lookahead<NL>(code, i, end, tab, nwg, curve);
        const uint32_t lo = (uint32_t)ins, imm = (uint32_t)(ins >> 32);
        const uint32_t op = lo & 0xffu, d = (lo >> 8) & 0xffu, x = (lo >> 16) & 0xffu, y = lo >> 24;
        if (op == V2_MUL2) {
            const uint32_t d4 = (lo >> 8) & 15u, x4 = (lo >> 12) & 15u, y4 = (lo >> 16) & 15u;
            const uint32_t e4 = (lo >> 20) & 15u, u4 = (lo >> 24) & 15u, v4 = lo >> 28;
            if (NL <= 16) {
                uint32_t a1[NL], b1[NL], r1[NL];
                S.load(a, x4); S.load(b, y4); S.load(a1, u4); S.load(b1, v4);
                mont_mul2<NL>(r, a, b, r1, a1, b1, P);
                S.store(d4, r); S.store(e4, r1);
            } else {
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    const uint32_t xs = h ? u4 : x4, ys = h ? v4 : y4;
                    S.load(a, xs);
                    if (xs == ys && UseSqr<NL>::value) mont_sqr<NL>(r, a, P);
                    else { S.load(b, ys); mont_mul<NL>(r, a, b, P); }
                    S.store(h ? e4 : d4, r);
                }
            }
        } else if (op == V2_SQR && UseSqr<NL>::value) {
            S.load(a, x);
            mont_sqr<NL>(r, a, P);
            S.store(d, r);
        } else if (op <= V2_SQR || op == V2_PAIR) {
            uint32_t dst = d;
            if (op == V2_PAIR) {                         // acc *= Pa_inv[pa] - Pb[pb].X  (ecm.c:1857-1859)
                uint32_t u[NL], v[NL];
                tload<NL>(u, tab, tab_base(imm & 0xffffu, nwg, curve, NL));
                tload<NL>(v, tab, tab_base(imm >> 16, nwg, curve, NL));
                mod_sub<NL>(a, u, v, P);
                S.load(b, V2_ACC);
                dst = V2_ACC;
            } else {
                S.load(a, x);
                S.load(b, y);
            }
            mont_mul<NL>(r, a, b, P);
            S.store(dst, r);
        } else if (op == V2_ADDSUB) {
            S.load(a, x); S.load(b, y);
            mod_add<NL>(r, a, b, P); S.store(d, r);
            mod_sub<NL>(r, a, b, P); S.store(imm, r);
        } else if (op == V2_ADD) {
            S.load(a, x); S.load(b, y); mod_add<NL>(r, a, b, P); S.store(d, r);
        } else if (op == V2_SUB) {
            S.load(a, x); S.load(b, y); mod_sub<NL>(r, a, b, P); S.store(d, r);
        } else if (op == V2_COPY) {
            S.load(a, x); S.store(d, a);
        } else if (op == V2_LDG) {
            tload<NL>(a, tab, tab_base(imm, nwg, curve, NL)); S.store(d, a);
        } else if (op == V2_STG) {
            S.load(a, x); tstore<NL>(tab, tab_base(imm, nwg, curve, NL), a);
        } else if (op == V2_INV) {
            vm2_inverse<NL, THREADS>(S.ptr(d), S.ptr(x), S.ptr(V2_ACC), Pg, inv_fail + curve);
        } else if (op == V2_ONE) {
#pragma unroll
            for (int k = 0; k < NL; k++) a[k] = P.one[k];
            S.store(d, a);
        }
    }
#pragma unroll 1
    for (uint32_t s = NG; s < NSLOT_S2; s++) {
        S.load(r, s);
#pragma unroll
        for (int k = 0; k < NL; k++) S.gl[(s * NL + k) * THREADS] = r[k];
    }
}

// ---- the pair loop as its own kernel ----------------------------------------------------------------
// Between two window shifts stage 2 is a long run of  acc *= Pa_inv[pa] - Pb[pb].X  (ecm.c:2526-2531,
// 77 % of all stage-2 multiplies at B2 = 100*B1).  The run needs no slot file: the accumulator lives in
// registers, the two operands stream from the tables ([entry][limb][curve], one 128-byte line per warp
// and limb; Pa_inv is L2-resident, Pb comes from HBM), and for narrow moduli the next step's operands
// are prefetched while the current product is computed.  No shared memory => 12-20 warps per SM.
// code[i] is the 64-bit V2_PAIR instruction; only its imm half is read.
template <int NL>
struct PairCfg {
    static constexpr bool DUAL = (NL <= 16);   // two accumulators per curve: twice the independent carry chains
    static constexpr int THREADS = 128;        // 4 warps: one per SM sub-partition
};

// One block = one item = (group of 128 curves, chunk of the run), same chunk-major schedule as stage 1
// so that every SM always holds its full complement of blocks even when the batch is not a multiple of
// the resident wave.  The product of a run is commutative, so splitting it over two accumulators (and
// re-joining them at the end of the chunk) yields the identical canonical residue.
template <int NL, int PT = PairCfg<NL>::THREADS>
__global__ void __launch_bounds__(PT, (PT >= 384) ? 1 : 0)
k_pair(const ModParams<NL> P, uint32_t *__restrict__ state2, uint32_t cap, const uint32_t *__restrict__ tab,
       const uint64_t *__restrict__ code, uint32_t npairs, uint32_t ncurves, uint32_t chunk_len, uint32_t groups, uint64_t item0)
{
    constexpr int T2 = S2Cfg<NL>::THREADS;                        // group size of the state2 layout
    const uint64_t item = item0 + blockIdx.x;
    const uint32_t g = (uint32_t)(item % groups);
    const uint32_t chunk = (uint32_t)(item / groups);
    const uint32_t curve = g * PT + threadIdx.x;
    if (curve >= ncurves) return;
    uint32_t i = chunk * chunk_len;
    const uint32_t end = (i + chunk_len < npairs) ? i + chunk_len : npairs;
    const uint32_t nwg = cap >> 5;
    if (PairCfg<NL>::DUAL) {
        // software pipeline: the table operands of steps i+2, i+3 are in flight while steps i, i+1 multiply
        uint32_t acc0[NL], acc1[NL], t0[NL], t1[NL], pu0[NL], pv0[NL], pu1[NL], pv1[NL];
        uint32_t *accp = state2 + ((size_t)(curve / T2) * NSLOT_S2 + V2_ACC) * (NL * T2) + (curve % T2);
#pragma unroll
        for (int k = 0; k < NL; k++) { acc0[k] = accp[k * T2]; acc1[k] = P.one[k]; }
        auto fetch = [&](uint32_t j, uint32_t (&pu)[NL], uint32_t (&pv)[NL]) {
            const uint32_t imm = (uint32_t)(__ldg(code + j) >> 32);
            tload<NL>(pu, tab, tab_base(imm & 0xffffu, nwg, curve, NL));
            tload<NL>(pv, tab, tab_base(imm >> 16, nwg, curve, NL));
        };
        if (i < end) fetch(i, pu0, pv0);
        if (i + 1 < end) fetch(i + 1, pu1, pv1);
#pragma unroll 1
        for (; i < end; i += 2) {
            mod_sub<NL>(t0, pu0, pv0, P);
            if (i + 1 < end) mod_sub<NL>(t1, pu1, pv1, P);
            else {
#pragma unroll
                for (int k = 0; k < NL; k++) t1[k] = P.one[k];
            }
            if (i + 2 < end) fetch(i + 2, pu0, pv0);
            if (i + 3 < end) fetch(i + 3, pu1, pv1);
            mont_mul2<NL>(acc0, acc0, t0, acc1, acc1, t1, P);
        }
        mont_mul<NL>(acc0, acc0, acc1, P);
#pragma unroll
        for (int k = 0; k < NL; k++) accp[k * T2] = acc0[k];
    } else {
        uint32_t acc[NL], u[NL], v[NL];
        uint32_t *accp = state2 + ((size_t)(curve / T2) * NSLOT_S2 + V2_ACC) * (NL * T2) + (curve % T2);
#pragma unroll
        for (int k = 0; k < NL; k++) acc[k] = accp[k * T2];
#pragma unroll 1
        for (; i < end; i++) {
            const uint32_t imm = (uint32_t)(__ldg(code + i) >> 32);
            tload<NL>(u, tab, tab_base(imm & 0xffffu, nwg, curve, NL));
            tload<NL>(v, tab, tab_base(imm >> 16, nwg, curve, NL));
            mod_sub<NL>(u, u, v, P);
            mont_mul<NL>(acc, acc, u, P);
        }
#pragma unroll
        for (int k = 0; k < NL; k++) accp[k * T2] = acc[k];
    }
}

// stage-2 wave set-up: Q = stage-1 result and the curve parameter move from the stage-1 state into
// the wave's table / slot file.  Wave curve c is batch curve first + c (clamped: padding lanes repeat
// the last curve).
template <int NL>
__global__ void k_s2_setup(const uint32_t *state1, Geom G1, uint32_t xslot, uint32_t zslot, uint32_t spslot,
                           uint32_t first, uint32_t count, uint32_t *state2, uint32_t cap2, uint32_t *tab,
                           uint32_t e_qx, uint32_t e_qz, uint8_t *inv_fail)
{
    const Geom G2{(uint32_t)S2Cfg<NL>::THREADS, (uint32_t)S2Cfg<NL>::THREADS, NSLOT_S2};
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cap2) return;
    uint32_t src = first + c; if (src >= count) src = count - 1;
    for (int k = 0; k < NL; k++) {
        tab[tab_base(e_qx, cap2 >> 5, c, NL) + (size_t)k * 32] = state1[G1.idx(src, xslot, k, NL)];
        tab[tab_base(e_qz, cap2 >> 5, c, NL) + (size_t)k * 32] = state1[G1.idx(src, zslot, k, NL)];
        for (uint32_t s = 0; s < NSLOT_S2; s++)
            state2[G2.idx(c, s, k, NL)] = (s == V2_SP) ? state1[G1.idx(src, spslot, k, NL)] : 0;
    }
    inv_fail[c] = 0;
}
// collect the accumulators (and failure flags) of a finished wave into batch-indexed arrays
template <int NL>
__global__ void k_s2_collect(const uint32_t *state2, uint32_t cap2, const uint8_t *inv_fail, uint32_t first, uint32_t n,
                             uint32_t count, uint32_t *acc_out, uint8_t *fail_out)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const Geom G2{(uint32_t)S2Cfg<NL>::THREADS, (uint32_t)S2Cfg<NL>::THREADS, NSLOT_S2};
    for (int k = 0; k < NL; k++) acc_out[(size_t)k * count + first + c] = state2[G2.idx(c, V2_ACC, k, NL)];
    fail_out[first + c] = inv_fail[c];
}

}  // inline namespace ECM_VNS
}  // namespace ecmb200
