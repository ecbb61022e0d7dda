// coop_s2.cuh -- stage 2 on the warp-cooperative layout (48 and 64 limbs): the stage-2 field-op machine and the pair loop
// of kernels.cuh with every value striped over L lanes, M limbs each (coop.cuh).  Same instruction stream (plan2.cpp),
// same table entries, same results; what changes is who holds which limb:
//   tables   entry e, lane l = curve*L + part:  tab[((e*nwl + l/32)*M + j)*32 + l%32],  nwl = cap*L/32
//            (tab_base() of kernels.cuh with "lane" for "curve" and M for NL: a warp still reads 128 contiguous bytes per limb)
//   state2   [group][slot][j][THREADS] with lane = curve_local*L + part              (Geom::idx with L lanes per curve)
// A 64-limb value is 16 registers per lane, so the pair loop keeps the accumulator, the operand difference and the
// prefetched operands of the next step in registers at 3-4 warps per scheduler, where the one-thread-per-curve kernels
// at this size ran every pair step through the all-shared slot machine at one warp per scheduler.
// The modular inverse (one per window shift) is computed by the group's lane 0 on the gathered value.
#pragma once
#include "kernels.cuh"
#include "coop.cuh"

namespace ecmb200 {
inline namespace ECM_VNS {

template <int M, int L>
struct CoopS2Cfg {
    static constexpr int NL = M * L;
    // hybrid slot file as for 20-32 limbs: the three work points and the accumulator (slots 0..6) stay in the L2-resident
    // state, 7 scratch slots in shared memory: 7 * M * 4 bytes per lane
    static constexpr int NG = NGLOBAL_S2;
    static constexpr int per_lane = (NSLOT_S2 - NG) * M * 4;
    static constexpr int THREADS = 384;
    static constexpr int smem = per_lane * THREADS;
    static constexpr int CURVES = THREADS / L;                // curves per group of the state2 layout
    static constexpr int PAIR_THREADS = 128;
    static_assert(smem <= kSmemBudget, "stage-2 scratch slots exceed shared memory");
};

// gather a striped value into lane 0 of the group (all lanes call; only lane 0's result is complete)
template <int M, int L>
__device__ __forceinline__ void coop_gather(uint32_t *full, const uint32_t (&v)[M], const coop::WarpComm<L> &cm)
{
#pragma unroll
    for (int q = 0; q < L; q++)
#pragma unroll
        for (int j = 0; j < M; j++) full[q * M + j] = cm.shfl(v[j], q);
}
template <int M, int L>
__device__ __forceinline__ void coop_scatter(uint32_t (&v)[M], const uint32_t *full, const coop::WarpComm<L> &cm)
{
#pragma unroll
    for (int q = 0; q < L; q++)
#pragma unroll
        for (int j = 0; j < M; j++) {
            const uint32_t w = cm.shfl(full[q * M + j], 0);
            if (cm.part == q) v[j] = w;
        }
}

// d = 1/x with the failure and stale-word semantics of vm2_inverse (kernels.cuh), on striped values.  Returns true when
// the accumulator was replaced (inversion failed on this curve).
template <int M, int L>
__device__ __noinline__ bool coop_inverse(uint32_t (&d)[M], const uint32_t (&x)[M], uint32_t (&acc)[M], const ModParams<M * L> *Pg,
                                          const coop::WarpComm<L> &cm)
{
    constexpr int NL = M * L;
    uint32_t a[NL], t[NL], ac[NL];
    coop_gather<M, L>(a, x, cm);
    uint32_t failed = 0;
    if (cm.part == 0) {
        uint32_t inv[NL], g[NL];
        if (nm_inverse<NL>(inv, g, a, Pg)) {
            nm_mul<NL>(t, inv, Pg->r3, Pg);
            stale_high_words<NL>(t, a, Pg);
        } else {
            failed = 1;
            nm_mul<NL>(ac, g, Pg->r2, Pg);
            nm_mul<NL>(ac, ac, Pg->rrefinv, Pg);
            nm_mul<NL>(t, a, Pg->rrefinv, Pg);
        }
    }
    failed = cm.shfl(failed, 0);
    coop_scatter<M, L>(d, t, cm);
    // every lane of the warp takes part in every shuffle: a warp holds 32/L curves and only some of them may have failed
    uint32_t na[M];
    coop_scatter<M, L>(na, ac, cm);
    if (failed) {
#pragma unroll
        for (int j = 0; j < M; j++) acc[j] = na[j];
    }
    return failed != 0;
}

template <int M, int L>
__global__ void __launch_bounds__(CoopS2Cfg<M, L>::THREADS, 1)
k_vm2_coop(const ModParams<M * L> P, const ModParams<M * L> *Pg, uint32_t *__restrict__ state2, uint32_t cap, uint32_t *__restrict__ tab,
           const uint64_t *__restrict__ code, uint64_t ncode, uint32_t chunk_len, uint32_t groups, uint64_t item0,
           uint8_t *__restrict__ inv_fail)
{
    typedef CoopS2Cfg<M, L> C;
    constexpr int THREADS = C::THREADS, NG = C::NG;
    extern __shared__ uint32_t smem[];
    const uint64_t item = item0 + blockIdx.x;
    const uint32_t g = (uint32_t)(item % groups);
    const uint64_t chunk = item / groups;
    const uint32_t lane = g * THREADS + threadIdx.x;                 // lane index of the wave = curve*L + part
    const uint32_t nwl = (cap * L) >> 5;
    HybridSlots<M, THREADS, NG, NSLOT_S2 - NG> S{state2 + (size_t)g * (NSLOT_S2 * M * THREADS) + threadIdx.x, smem + threadIdx.x};
    const CoopField<M, L> F(P);
    uint32_t a[M], b[M], r[M];
#pragma unroll 1
    for (uint32_t s = NG; s < NSLOT_S2; s++) {
#pragma unroll
        for (int k = 0; k < M; k++) r[k] = S.gl[(s * M + k) * THREADS];
        S.store(s, r);
    }
    uint64_t i = chunk * chunk_len;
    const uint64_t end = (i + chunk_len < ncode) ? i + chunk_len : ncode;
#pragma unroll 1
    for (; i < end; i++) {
        const uint64_t ins = __ldg(code + i);
        vm2_lookahead<M>(code, i, end, tab, nwl, lane);
        const uint32_t lo = (uint32_t)ins, imm = (uint32_t)(ins >> 32);
        const uint32_t op = lo & 0xffu, d = (lo >> 8) & 0xffu, x = (lo >> 16) & 0xffu, y = lo >> 24;
        if (op == V2_MUL2) {
            const uint32_t d4 = (lo >> 8) & 15u, x4 = (lo >> 12) & 15u, y4 = (lo >> 16) & 15u;
            const uint32_t e4 = (lo >> 20) & 15u, u4 = (lo >> 24) & 15u, v4 = lo >> 28;
            uint32_t a1[M], b1[M];
            S.load(a, x4); S.load(b, y4); S.load(a1, u4); S.load(b1, v4);     // all four operands before either result
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                F.mul(a, b);
                S.store(h ? e4 : d4, a);
#pragma unroll
                for (int k = 0; k < M; k++) { a[k] = a1[k]; b[k] = b1[k]; }
            }
        } else if (op <= V2_SQR || op == V2_PAIR) {
            uint32_t dst = d;
            if (op == V2_PAIR) {                         // acc *= Pa_inv[pa] - Pb[pb].X  (ecm.c:1857-1859)
                uint32_t u[M], v[M];
                tload<M>(u, tab, tab_base(imm & 0xffffu, nwl, lane, M));
                tload<M>(v, tab, tab_base(imm >> 16, nwl, lane, M));
                F.sub(a, u, v);
                S.load(b, V2_ACC);
                dst = V2_ACC;
            } else {
                S.load(a, x);
                S.load(b, y);
            }
            F.mul(a, b);
            S.store(dst, a);
        } else if (op == V2_ADDSUB) {
            S.load(a, x); S.load(b, y);
            F.add(r, a, b); S.store(d, r);
            F.sub(r, a, b); S.store(imm, r);
        } else if (op == V2_ADD) {
            S.load(a, x); S.load(b, y); F.add(r, a, b); S.store(d, r);
        } else if (op == V2_SUB) {
            S.load(a, x); S.load(b, y); F.sub(r, a, b); S.store(d, r);
        } else if (op == V2_COPY) {
            S.load(a, x); S.store(d, a);
        } else if (op == V2_LDG) {
            tload<M>(a, tab, tab_base(imm, nwl, lane, M)); S.store(d, a);
        } else if (op == V2_STG) {
            S.load(a, x); tstore<M>(tab, tab_base(imm, nwl, lane, M), a);
        } else if (op == V2_INV) {
            S.load(a, x); S.load(b, V2_ACC);
            if (coop_inverse<M, L>(r, a, b, Pg, F.cm)) {
                S.store(V2_ACC, b);
                if (F.cm.part == 0) inv_fail[lane / L] = 1;
            }
            S.store(d, r);
        } else if (op == V2_ONE) {
#pragma unroll
            for (int k = 0; k < M; k++) a[k] = P.one[F.cm.part * M + k];
            S.store(d, a);
        }
    }
#pragma unroll 1
    for (uint32_t s = NG; s < NSLOT_S2; s++) {
        S.load(r, s);
#pragma unroll
        for (int k = 0; k < M; k++) S.gl[(s * M + k) * THREADS] = r[k];
    }
}

// the pair loop: accumulator, operand difference and the next step's operands in registers, no shared memory
template <int M, int L>
__global__ void __launch_bounds__(CoopS2Cfg<M, L>::PAIR_THREADS)
k_pair_coop(const ModParams<M * L> P, uint32_t *__restrict__ state2, uint32_t cap, const uint32_t *__restrict__ tab,
            const uint64_t *__restrict__ code, uint32_t npairs, uint32_t nlanes, uint32_t chunk_len, uint32_t groups, uint64_t item0)
{
    typedef CoopS2Cfg<M, L> C;
    constexpr int T2 = C::THREADS;                                   // lanes per group of the state2 layout
    const uint64_t item = item0 + blockIdx.x;
    const uint32_t g = (uint32_t)(item % groups);
    const uint32_t chunk = (uint32_t)(item / groups);
    const uint32_t lane = g * C::PAIR_THREADS + threadIdx.x;
    if (lane >= nlanes) return;                                      // whole warps: nlanes is a multiple of 32
    uint32_t i = chunk * chunk_len;
    const uint32_t end = (i + chunk_len < npairs) ? i + chunk_len : npairs;
    const uint32_t nwl = (cap * L) >> 5;
    const CoopField<M, L> F(P);
    uint32_t acc[M], t[M], pu[M], pv[M];
    uint32_t *accp = state2 + ((size_t)(lane / T2) * NSLOT_S2 + V2_ACC) * (M * T2) + (lane % T2);
#pragma unroll
    for (int k = 0; k < M; k++) acc[k] = accp[k * T2];
    auto fetch = [&](uint32_t j) {
        const uint32_t imm = (uint32_t)(__ldg(code + j) >> 32);
        tload<M>(pu, tab, tab_base(imm & 0xffffu, nwl, lane, M));
        tload<M>(pv, tab, tab_base(imm >> 16, nwl, lane, M));
    };
    if (i < end) fetch(i);
#pragma unroll 1
    for (; i < end; i++) {
        F.sub(t, pu, pv);
        if (i + 1 < end) fetch(i + 1);                               // in flight while the product runs
        F.mul(acc, t);
    }
#pragma unroll
    for (int k = 0; k < M; k++) accp[k * T2] = acc[k];
}

// wave set-up / collect on the striped layouts: thread = (wave curve, limb block)
template <int M, int L>
__global__ void k_s2_setup_coop(const uint32_t *state1, Geom G1, uint32_t xslot, uint32_t zslot, uint32_t spslot,
                                uint32_t first, uint32_t count, uint32_t *state2, uint32_t cap2, uint32_t *tab,
                                uint32_t e_qx, uint32_t e_qz, uint8_t *inv_fail)
{
    typedef CoopS2Cfg<M, L> C;
    constexpr int NL = M * L;
    const Geom G2{(uint32_t)C::CURVES, (uint32_t)C::THREADS, NSLOT_S2, (uint32_t)L};
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cap2) return;
    uint32_t src = first + c; if (src >= count) src = count - 1;
    const uint32_t nwl = (cap2 * L) >> 5;
    for (int k = 0; k < NL; k++) {
        const uint32_t lane = c * L + k / M, j = k % M;
        tab[tab_base(e_qx, nwl, lane, M) + (size_t)j * 32] = state1[G1.idx(src, xslot, k, NL)];
        tab[tab_base(e_qz, nwl, lane, M) + (size_t)j * 32] = state1[G1.idx(src, zslot, k, NL)];
        for (uint32_t s = 0; s < NSLOT_S2; s++)
            state2[G2.idx(c, s, k, NL)] = (s == V2_SP) ? state1[G1.idx(src, spslot, k, NL)] : 0;
    }
    inv_fail[c] = 0;
}
template <int M, int L>
__global__ void k_s2_collect_coop(const uint32_t *state2, uint32_t cap2, const uint8_t *inv_fail, uint32_t first, uint32_t n,
                                  uint32_t count, uint32_t *acc_out, uint8_t *fail_out)
{
    typedef CoopS2Cfg<M, L> C;
    constexpr int NL = M * L;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const Geom G2{(uint32_t)C::CURVES, (uint32_t)C::THREADS, NSLOT_S2, (uint32_t)L};
    for (int k = 0; k < NL; k++) acc_out[(size_t)k * count + first + c] = state2[G2.idx(c, V2_ACC, k, NL)];
    fail_out[first + c] = inv_fail[c];
}

}  // inline namespace ECM_VNS
}  // namespace ecmb200
